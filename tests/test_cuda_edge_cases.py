"""GPU: edge cases and error behaviour of the C ABI — the situations the reference's tests poke at
(single particle, degenerate observation space, empty batches, bad arguments, short streams)."""
import ctypes as C

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


def tiny_desc(S=3, A=2, O=1):
    return dict(S=S, A=A, O=O, feat_s=[S], feat_o=[O], tabular=1, domain=0, action_draw=0,
                start_kind=0, start_ip=[0, 0, 0, 0], dom_ip=np.zeros(32, np.int32), dom_dp=np.zeros(8),
                rew_sa=np.arange(S * A, dtype=np.float64), rew_as2=np.zeros(A * S),
                term_sa=np.zeros(S * A, np.uint8), term_as2=np.zeros(A * S, np.uint8))


def test_single_particle_and_single_observation(ctx):
    """N = 1 and O = 1: the likelihood is exactly 1 (BAFlatModel.cpp:117-120), the lone particle is
    its own ancestor, counts grow by one transition and one observation count per update."""
    import fba_pomdp_b200 as fba
    import pyoracle as O
    d = tiny_desc()
    S, A = 3, 2
    sim = fba.BAPOMDP(ctx, d, np.ones((1, A), np.uint32), np.ones((1, A), np.uint32))
    stride = sim.structure_size(0)
    assert stride == A * (S * S + S * 1)
    counts = np.arange(1, stride + 1, dtype=np.float32)[None, :]
    b = fba.BAImportanceSampling(1)
    b.initiate(sim, struct_id=[0], counts=counts, state=[1])
    words = np.random.RandomState(0).randint(0, 2**32, 16, dtype=np.uint64).astype(np.uint32)
    rng = fba.Rng.replay(words)
    lik = b.updateEstimation(1, 0, rng)
    assert lik == 1.0 and rng.cursor == 4 + 2  # 2 uniforms for the step, 1 for the resample
    got = b.download()
    m = O.Model(d)
    st = O.Structs(m, np.ones((1, A), np.uint32), np.ones((1, A), np.uint32))
    ob = O.Belief(1, got["counts"].shape[1])
    ob.counts[0, :stride] = counts[0]
    ob.state[:] = 1
    ob.total_weight = 1.0
    orng = O.Rng(words)
    assert O.is_update(m, st, ob, 1, 0, orng) == 1.0
    ob2, _ = O.is_resample(ob, orng)
    np.testing.assert_array_equal(got["counts"], ob2.counts)
    np.testing.assert_array_equal(got["state"], ob2.state)
    assert got["w"][0] == 1.0 and got["counts"].sum() == counts.sum() + 2
    b.free()
    sim.close()


def test_stride_is_padded_to_16_bytes(ctx):
    import fba_pomdp_b200 as fba
    d = tiny_desc(S=3, A=1, O=2)  # 9 + 6 = 15 cells -> stride 16
    sim = fba.BAPOMDP(ctx, d, np.ones((1, 1), np.uint32), np.ones((1, 1), np.uint32))
    b = fba.BAImportanceSampling(5)
    b.initiate(sim, struct_id=np.zeros(5, np.int32), counts=np.ones((5, 15), np.float32), state=np.zeros(5, np.int32))
    assert b.L.fba_belief_stride(b.h) == 16
    got = b.download()
    assert got["counts"].shape == (5, 16) and not got["counts"][:, 15].any()
    b.free()
    sim.close()


def test_empty_rollout_batch_and_zero_depth(ctx):
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    b = fba.BAImportanceSampling(4)
    b.initiate(sim, struct_id=g["is/init_struct_id"][:4], counts=g["is/init_counts"][:4], state=g["is/init_state"][:4])
    assert len(fba.rollouts(b, [], [], [], 0.95, fba.Rng.philox(1))) == 0
    ret = fba.rollouts(b, [0, 1], [0, 1], [0, 0], 0.95, fba.Rng.philox(1))  # depth 0: no steps, return 0
    np.testing.assert_array_equal(ret, [0.0, 0.0])
    b.free()
    sim.close()


def test_error_statuses(ctx):
    """Bad arguments come back as FBA_ERR_INVALID with a message (the reference throws strings);
    a short replay stream is FBA_ERR_RNG_UNDERRUN and leaves the cursor alone."""
    import fba_pomdp_b200 as fba
    from fba_pomdp_b200 import capi
    g = G.load("tiger")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    b = fba.BAImportanceSampling(8)
    b.initiate(sim, struct_id=g["is/init_struct_id"][:8], counts=g["is/init_counts"][:8], state=g["is/init_state"][:8])
    with pytest.raises(fba.FbaError) as e:
        b.update(7, 0, fba.Rng.philox(1))
    assert e.value.status == capi.ERR_INVALID and "action" in str(e.value)
    with pytest.raises(fba.FbaError) as e:
        b.update(0, 9, fba.Rng.philox(1))
    assert e.value.status == capi.ERR_INVALID and "observation" in str(e.value)
    short = fba.Rng.replay(np.zeros(5, np.uint32))
    with pytest.raises(fba.FbaError) as e:
        b.update(2, 0, short)
    assert e.value.status == capi.ERR_RNG_UNDERRUN and short.cursor == 0
    with pytest.raises(fba.FbaError):
        fba.rollouts(b, [99], [0], [3], 0.95, fba.Rng.philox(1))  # particle index out of range
    with pytest.raises(fba.FbaError):
        fba.rollouts(b, [0], [0], [3], 1.5, fba.Rng.philox(1))  # discount outside (0, 1]
    # an unknown structure id is refused at upload
    rc = b.L.fba_belief_upload(b.h, 0, 1, None, fba.capi.ptr(np.array([5], np.int32)), None, None)
    assert rc == capi.ERR_INVALID
    # structure table capacity
    with pytest.raises(fba.FbaError) as e:
        sim.add_structures(np.zeros((1, 3), np.uint32), np.zeros((1, 3), np.uint32))
    assert e.value.status == capi.ERR_CAPACITY
    b.free()
    sim.close()


def test_model_validation(ctx):
    import fba_pomdp_b200 as fba
    d = tiny_desc()
    d["feat_s"] = [2]  # does not multiply to S
    with pytest.raises(fba.FbaError):
        fba.BAPOMDP(ctx, d, np.ones((1, 2), np.uint32), np.ones((1, 2), np.uint32))
    d = tiny_desc()
    with pytest.raises(fba.FbaError):  # parent mask names a feature that does not exist
        fba.BAPOMDP(ctx, d, np.full((1, 2), 2, np.uint32), np.ones((1, 2), np.uint32))


def test_zero_weight_particles_never_resampled_replay(ctx):
    """A particle whose weight is exactly 0 can never be drawn (WeightedFilter.cpp:163-191)."""
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    n = 64
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    b = fba.BAImportanceSampling(n)
    counts = np.tile(g["is/init_counts"][0], (n, 1))
    counts[:, 0] = np.arange(n)
    b.initiate(sim, struct_id=np.zeros(n, np.int32), counts=counts, state=np.zeros(n, np.int32))
    w = np.zeros(n)
    w[[3, 17, 40]] = [0.5, 0.25, 0.25]
    assert b.L.fba_belief_upload(b.h, 0, n, None, None, None, fba.capi.ptr(w)) == 0
    words = np.random.RandomState(4).randint(0, 2**32, 2 * n, dtype=np.uint64).astype(np.uint32)
    # the host mirror of _total_weight must match the uploaded weights for a replay resample
    b.L.fba_belief_download  # (weights were normalised by construction: total 1.0)
    rng = fba.Rng.replay(words)
    b.resample(rng)
    ids = b.download()["counts"][:, 0].astype(int)
    assert set(ids) <= {3, 17, 40} and rng.exhausted
    b.free()
    sim.close()


def test_impossible_observation_keeps_the_belief_intact(ctx):
    """Every particle gives the real observation probability 0 (its count row is all zeros): the
    step likelihood is 0 and the normalised weights are 0/0. The reference's behaviour is undefined
    there (NaN weights, WeightedFilter.cpp:130-143); the CUDA path must stay memory-safe and
    deterministic: no offspring are assigned, so every particle stays where it is, the weights go
    back to uniform and the next update works."""
    import fba_pomdp_b200 as fba
    S, A, O = 2, 1, 2
    d = tiny_desc(S=S, A=A, O=O)
    sim = fba.BAPOMDP(ctx, d, np.ones((1, A), np.uint32), np.ones((1, A), np.uint32))
    n = 5000
    stride = sim.structure_size(0)          # T[a][s][s'] (4 cells) then O[a][s'][o] (4 cells)
    counts = np.ones((n, stride), np.float32)
    counts[:, 4 + 1] = 0.0                  # P(o = 1 | s' = 0) = 0
    counts[:, 4 + 3] = 0.0                  # P(o = 1 | s' = 1) = 0
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, struct_id=np.zeros(n, np.int32), counts=counts, state=np.zeros(n, np.int32))
    rng = fba.Rng.philox(3)
    lik = b.update(0, 1, rng)
    assert lik == 0.0
    b.resample(rng)
    got = b.download()
    np.testing.assert_array_equal(got["w"], np.full(n, 1.0 / n))
    assert set(np.unique(got["state"])) <= {0, 1}
    # the update itself ran (one transition count and one observation count per particle) ...
    np.testing.assert_array_equal(got["counts"].astype(np.float64).sum(1), np.full(n, 6.0 + 2.0))
    # ... and the belief is usable afterwards
    lik2 = b.updateEstimation(0, 0, rng)
    assert 0.99 < lik2 <= 1.0
    b.free()
    sim.close()


def test_states_out_of_range_are_rejected(ctx):
    import fba_pomdp_b200 as fba
    d = tiny_desc()
    sim = fba.BAPOMDP(ctx, d, np.ones((1, 2), np.uint32), np.ones((1, 2), np.uint32))
    b = fba.BAImportanceSampling(4)
    counts = np.ones((4, sim.structure_size(0)), np.float32)
    with pytest.raises(fba.capi.FbaError):
        b.initiate(sim, struct_id=np.zeros(4, np.int32), counts=counts, state=np.array([0, 1, 3, 0], np.int32))
    b.free()
    b = fba.BAImportanceSampling(4)
    with pytest.raises(fba.capi.FbaError):
        b.initiate(sim, proto_struct_id=[0], proto_counts=counts[:1], particle_proto=None,
                   state=np.array([0, -1, 0, 0], np.int32))
    b.free()
    sim.close()


def test_back_buffer_is_allocated_on_first_use(ctx):
    """In-place (PHILOX) importance sampling never touches the second particle buffer; a full-copy
    resample (REPLAY) allocates it on the spot and gives the same belief as before the change."""
    import torch
    import fba_pomdp_b200 as fba
    g = G.load("sysadmin")
    n = 20000
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    free0 = torch.cuda.mem_get_info(0)[0]
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, proto_struct_id=[0], proto_counts=g["is/init_counts"][:1], particle_proto=None,
               state=np.zeros(n, np.int32))
    rng = fba.Rng.philox(1)
    b.updateEstimation(int(g.a[0]), int(g.o[0]), rng)
    ctx.synchronize()
    used_inplace = free0 - torch.cuda.mem_get_info(0)[0]
    block = n * sim.max_structure_size() * 4
    assert used_inplace < 1.6 * block, (used_inplace, block)     # one count buffer, not two
    words = np.random.RandomState(0).randint(0, 2**32, 64 * n, dtype=np.uint64).astype(np.uint32)
    b.updateEstimation(int(g.a[1]), int(g.o[1]), fba.Rng.replay(words))   # full copy: needs the back buffer
    ctx.synchronize()
    assert free0 - torch.cuda.mem_get_info(0)[0] >= 2 * block
    d = b.download(counts=False)
    assert abs(d["w"].sum() - 1.0) < 1e-9
    b.free()
    sim.close()
