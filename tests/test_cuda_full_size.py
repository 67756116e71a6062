"""GPU: REPLAY parity at the particle counts BASELINE.json's configurations name (VERDICT r1 "missing" #2).

The reference itself cannot run there — its resample is O(N^2) (WeightedFilter.cpp:163-191 inside
ImportanceSampler.hpp:71-94: 23 minutes per update at 10^6) — so the checker is the C oracle, which is
pinned to the reference at small N by tests/test_oracle_vs_golden.py, with two changes that keep its
arithmetic: the belief is processed block by block (the running weight total carried across blocks in
the reference's order: orc_is_propose), and the N weighted draws share one sequential pass over the
remainders (orc_weighted_sample_many, pinned draw by draw against the O(N) scan). Count blocks are
compared through position-sensitive 64-bit checksums (orc_block_checksums on the host, the same wrapping
integer arithmetic in torch on the device) so that neither side has to hold two copies of 10-25 GB.

Also here: the parallel evaluation of the reference's sequential weight chains (fba_kernels.cuh,
k_chain_*) against the one-thread kernel, bit for bit, on adversarial weight vectors."""
import ctypes as C
import time

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fba_pomdp_b200 as fba
    c = fba.Context(0)
    yield c
    c.close()


class _Raw:
    def __init__(self, ptr, n, typestr, itemsize):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_checksums(b, stride):
    """orc_block_checksums on the device: sum_k bits(c_k) * (2k + 1) * golden, wrapping 64-bit"""
    import torch
    n = b.size()
    flat = torch.as_tensor(_Raw(b.L.fba_belief_counts_ptr(b.h), n * stride, "<i4", 4), device="cuda")
    bits = flat.view(n, stride)
    k = torch.arange(stride, device="cuda", dtype=torch.int64)
    golden = torch.tensor(np.array([0x9E3779B97F4A7C15], np.uint64).view(np.int64)[0], device="cuda")
    mult = (2 * k + 1) * golden
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    step = 65536
    for i in range(0, n, step):
        blk = bits[i:i + step].to(torch.int64) & 0xFFFFFFFF
        out[i:i + step] = (blk * mult[None, :]).sum(1)
    return out.cpu().numpy().view(np.uint64)


def device_doubles(ptr, n):
    import torch
    return torch.as_tensor(_Raw(ptr, n, "<f8", 8), device="cuda").cpu().numpy()


def prototypes(g):
    sid, counts = g["is/init_struct_id"], g["is/init_counts"]
    seen, psid, pc = {}, [], []
    for i in range(len(sid)):
        k = (int(sid[i]), counts[i].tobytes())
        if k not in seen:
            seen[k] = len(psid)
            psid.append(int(sid[i]))
            pc.append(counts[i])
    return np.array(psid, np.int32), np.stack(pc)


FULL = [
    # fixture, particles, BASELINE.json configuration
    ("ftiger", 100_000, "configs[1]: factored tiger, 10^5 particles"),
    ("gridworld3", 1_000_000, "configs[2]: gridworld, 10^6 particles"),
    ("ca", 1_000_000, "configs[3]: collision avoidance, 10^6 particles"),
    ("sysadmin", 1_250_000, "configs[4]: sysadmin-10, one GPU's shard of 10^7"),
]


@pytest.mark.parametrize("name,n,what", FULL)
def test_replay_update_and_resample_at_the_configs_own_size(ctx, name, n, what):
    import fba_pomdp_b200 as fba
    import pyoracle as O
    g = G.load(name)
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par, max_structures=len(g.t_par))
    psid, protos = prototypes(g)
    stride = (protos.shape[1] + 3) & ~3         # the library pads blocks to 16 bytes
    protos = np.pad(protos, ((0, 0), (0, stride - protos.shape[1])))
    rs = np.random.RandomState(len(name) + n)
    pproto = rs.randint(0, len(psid), n).astype(np.int32)
    state = g["is/init_state"][rs.randint(0, len(g["is/init_state"]), n)].astype(np.int32)
    t = next(t for t in range(len(g.a)) if not (g.flags[t] & 1))
    a, o = int(g.a[t]), int(g.o[t])
    J = sim.FS + sim.FO
    words_u = rs.randint(0, 2**32, size=2 * J * n, dtype=np.uint64).astype(np.uint32)
    words_r = rs.randint(0, 2**32, size=2 * n, dtype=np.uint64).astype(np.uint32)

    # ---- CUDA, REPLAY mode ----
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, proto_struct_id=psid, proto_counts=protos, particle_proto=pproto, state=state, stride=stride)
    assert b.L.fba_belief_stride(b.h) == stride
    rng = fba.Rng.replay(words_u)
    ctx.synchronize()
    t0 = time.perf_counter()
    lik = b.update(a, o, rng)
    ms_update = (time.perf_counter() - t0) * 1e3
    assert rng.exhausted
    d = b.download(counts=False)
    sums_gpu = device_checksums(b, stride)

    # ---- oracle, block by block ----
    m = O.Model(g.desc)
    st = O.Structs(m, g.t_par, g.o_par)
    w = np.empty(n, np.float64)
    o_state = np.empty(n, np.int32)
    o_sums = np.empty(n, np.uint64)
    total = 0.0
    blk = 32768
    for i0 in range(0, n, blk):
        i1 = min(n, i0 + blk)
        ob = O.Belief(i1 - i0, stride)
        ob.counts[:] = protos[pproto[i0:i1]]
        ob.state[:] = state[i0:i1]
        ob.struct_id[:] = psid[pproto[i0:i1]]
        ob.w[:] = 1.0 / n
        total = O.is_propose(m, st, ob, a, o, O.Rng(words_u[2 * J * i0:2 * J * i1]), total)
        w[i0:i1], o_state[i0:i1] = ob.w, ob.state
        o_sums[i0:i1] = O.block_checksums(ob.counts)
    total_weight = O.normalize(w, total)

    assert lik == total, (lik, total)
    np.testing.assert_array_equal(d["state"], o_state)
    np.testing.assert_array_equal(sums_gpu, o_sums)
    np.testing.assert_array_equal(d["w"], w)                      # bit-identical (tolerance would be 1e-5)
    assert d["total_weight"] == total_weight
    redo = b.L.fba_belief_chain_recomputed(b.h)
    assert 0 < redo < 400, redo                                    # the parallel chains ran, and mostly shifted

    # ---- resample: N weighted draws ----
    rng = fba.Rng.replay(words_r)
    ctx.synchronize()
    t0 = time.perf_counter()
    b.resample(rng)
    ctx.synchronize()
    ms_resample = (time.perf_counter() - t0) * 1e3
    assert rng.exhausted
    anc = O.weighted_sample_many(w, total_weight, O.Rng(words_r), n)
    d2 = b.download(counts=False)
    np.testing.assert_array_equal(d2["state"], o_state[anc])
    np.testing.assert_array_equal(device_checksums(b, stride), o_sums[anc])
    np.testing.assert_array_equal(d2["w"], np.full(n, 1.0 / n))
    assert d2["total_weight"] == O.sequential_uniform_total(n)
    print("\nREPLAY %s: %d particles, update %.2f ms, resample %.2f ms (incl. word upload), %d segments recomputed"
          % (what, n, ms_update, ms_resample, redo))
    b.free()
    sim.close()


def _weights(kind, n, rs):
    if kind == "uniform":
        return rs.random_sample(n)
    if kind == "half-zero":
        w = rs.random_sample(n)
        w[rs.random_sample(n) < 0.5] = 0.0
        w[0] = w[-1] = 0.5
        return w
    if kind == "lognormal":
        return np.exp(rs.normal(0, 15, n))
    if kind == "equal":
        return np.full(n, 1.0 / n)
    if kind == "powers-of-two":                 # every addition is exact or an exact tie
        return np.ldexp(1.0, rs.randint(-60, 0, n))
    if kind == "collapsed":
        w = np.full(n, 1e-300)
        w[n // 2] = 1.0
        return w
    if kind == "ascending":
        return np.sort(np.exp(rs.normal(0, 8, n)))
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["uniform", "half-zero", "lognormal", "equal", "powers-of-two", "collapsed",
                                  "ascending"])
@pytest.mark.parametrize("n", [1025, 50_000, 1_000_003])
def test_parallel_weight_chains_are_bit_identical_to_the_sequential_kernel(ctx, kind, n):
    """Chains A, B, C of the REPLAY normalisation, evaluated (i) by one thread (k_seq_normalize: literally
    the reference's loops) and (ii) in parallel segments with the shift / recompute pass: the same total,
    the same normalised weights, the same _total_weight, every remainder R_k identical — on weight vectors
    chosen to hit exact ties, binade crossings, zeros and collapse."""
    import fba_pomdp_b200 as fba
    from fba_pomdp_b200.beliefs import _check
    from fba_pomdp_b200.capi import ptr
    g = G.load("tiger")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    rs = np.random.RandomState(n % 1000 + len(kind))
    w0 = np.ascontiguousarray(_weights(kind, n, rs), np.float64)
    t = next(t for t in range(len(g.a)) if not (g.flags[t] & 1))
    a, o = int(g.a[t]), int(g.o[t])
    words = rs.randint(0, 2**32, size=4 * n, dtype=np.uint64).astype(np.uint32)
    words_r = rs.randint(0, 2**32, size=2 * n, dtype=np.uint64).astype(np.uint32)
    res = []
    for threshold in (1 << 40, 0):          # never / always the parallel evaluation
        ctx.set_option("parallel_chains_min", threshold)
        b = fba.BAImportanceSampling(n)
        b.initiate(sim, proto_struct_id=[0], proto_counts=g["is/init_counts"][:1], particle_proto=None,
                   state=np.zeros(n, np.int32))
        _check(ctx.h, b.L.fba_belief_upload(b.h, 0, n, None, None, None, ptr(w0)))
        lik = b.update(a, o, fba.Rng.replay(words))
        d = b.download(counts=False)
        R = device_doubles(b.L.fba_belief_aux_ptr(b.h), n)[1:].copy()
        b.resample(fba.Rng.replay(words_r))
        anc_states = b.download(counts=False)["state"]
        # chain C alone (weights now uniform)
        b.resample(fba.Rng.replay(words_r))
        R2 = device_doubles(b.L.fba_belief_aux_ptr(b.h), n)[1:].copy()
        res.append((lik, d["w"], d["total_weight"], R, anc_states, R2, b.L.fba_belief_chain_recomputed(b.h)))
        b.free()
    ctx.set_option("parallel_chains_min", 8192)
    sim.close()
    seq, par = res
    assert seq[6] == -1 and par[6] >= 1
    assert seq[0] == par[0] and seq[2] == par[2]
    for k in (1, 3, 4, 5):
        np.testing.assert_array_equal(seq[k], par[k])
