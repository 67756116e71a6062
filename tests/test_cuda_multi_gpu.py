"""GPU (>= 2 devices): the sharded belief, both exchange mechanisms, two domains, dense and journal storage
(tools/check_sharded.py). Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange,fixture,storage", [("p2p", "sysadmin", "dense"), ("allgather", "sysadmin", "dense"),
                                                      ("p2p", "ca", "dense"), ("p2p", "sysadmin", "journal")])
def test_sharded_invariants_two_gpus(exchange, fixture, storage):
    """tools/check_sharded.py on 2 GPUs: invariants, a x1000 weight skew, and equality of the sharded posterior
    with a single-GPU belief's (likelihoods and state histograms within 5 standard errors)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "check_sharded.py"), exchange, "20000", fixture, storage]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sharded check ok" in r.stdout
