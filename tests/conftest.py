import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Order of the GPU suite under `pytest -x`: bit-exact / deterministic files first, files whose results
# depend on thread timing inside a kernel (concurrent POMCP simulations sharing a tree through atomics)
# last — a statistical test must never hide the deterministic evidence behind it.
_FILE_ORDER = [
    "test_capi_cpu", "test_oracle_vs_golden", "test_sharded_host_logic",
    "test_cuda_parity", "test_cuda_full_size", "test_cuda_edge_cases", "test_cuda_delta_storage", "test_cuda_journal",
    "test_cuda_bd_score", "test_cuda_native", "test_cuda_runs", "test_cuda_sampled_dirichlet", "test_cuda_mh",
    "test_cuda_multi_gpu", "test_dropin_adapter", "test_cuda_tree",
]


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        rank = _FILE_ORDER.index(name) if name in _FILE_ORDER else len(_FILE_ORDER) - 2
        timing = 1 if ("cuda-tree-po-uct" in item.name or "lockstep" in item.name) else 0
        # the two tests that judge a concurrent search by its statistics run after everything else: whatever they
        # do on a given box, `pytest -x` has by then executed every other test
        if item.name.startswith(("test_wave_width_trades", "test_flat_and_delta_beliefs")):
            timing = 2
        return (rank + 100 * timing,)
    items.sort(key=key)  # stable: keeps the order inside a file
