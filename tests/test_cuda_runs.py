"""GPU: many independent runs batched on one GPU (fba_runs_*). One kernel launch
advances the belief of every run (one CTA per run). The parity contract is exact: run r of the batch
is BIT-IDENTICAL — states, structure ids, counts, weights, step likelihood, sampled index — to a
stand-alone BAImportanceSampling of the same size driven by Rng.philox(seed + r) through the same
sequence of calls (that stand-alone path is the one checked against the oracle and the reference
elsewhere in this suite)."""
import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fba_pomdp_b200 as fba
    c = fba.Context(0)
    # the stand-alone beliefs the batches are compared with take the launch-per-phase path, whatever
    # their size (small ones would otherwise use the same fused kernel as the batch)
    c.set_option("fused_update", 0)
    yield c
    c.close()


def test_fused_small_belief_update_equals_launch_per_phase(ctx):
    """fba_belief_update_estimation on a belief of <= 2048 particles runs update + resample in ONE
    launch; the result is bit-identical to the nine-launch path."""
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    out = {}
    for fused in (0, 1):
        ctx.set_option("fused_update", fused)
        b = fba.BAImportanceSampling(1500)
        rng = fba.Rng.philox(11)
        b.initiate_sampled(sim, [0], g["is/init_counts"][:1], None, rng)
        l0 = ctx.launches
        liks = [b.updateEstimation(2, t % 2, rng) for t in range(6)]
        launches = ctx.launches - l0
        d = b.download()
        out[fused] = (liks, d, launches, b.resample_stats())
        b.free()
    ctx.set_option("fused_update", 0)
    assert out[0][0] == out[1][0] and out[0][3] == out[1][3]
    for k in ("state", "struct_id", "w", "counts"):
        np.testing.assert_array_equal(out[0][1][k], out[1][1][k])
    assert out[1][2] == 6 and out[0][2] > 6 * 5
    sim.close()


def _protos(g):
    sid0, counts0 = g["is/init_struct_id"], g["is/init_counts"]
    used, sid0 = np.unique(sid0, return_inverse=True)
    keys, psid, pc = {}, [], []
    for i in range(len(sid0)):
        k = (int(sid0[i]), counts0[i].tobytes())
        if k not in keys:
            keys[k] = len(psid)
            psid.append(int(sid0[i]))
            pc.append(counts0[i])
    freq = np.bincount([keys[(int(sid0[i]), counts0[i].tobytes())] for i in range(len(sid0))],
                       minlength=len(psid)).astype(np.float64)
    return used, np.array(psid, np.int32), np.stack(pc), freq


# name, runs, particles per run (1024 = one scan tile; 300 / 2500: ragged and multi-tile runs)
CASES = [("tiger", 7, 1024), ("tiger", 3, 2500), ("ftiger", 4, 300), ("sysadmin3", 5, 700), ("ca", 3, 1500),
         ("gridworld3", 3, 640)]


@pytest.mark.parametrize("name,R,n", CASES)
def test_batched_runs_equal_stand_alone_beliefs(ctx, name, R, n):
    import fba_pomdp_b200 as fba
    g = G.load(name)
    used, psid, pc, freq = _protos(g)
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par[used], g.o_par[used])
    seed = 1234
    A, O = int(g.desc["A"]), int(g.desc["O"])
    rs = np.random.RandomState(5)
    T = 5
    acts = rs.randint(0, A, (T, R)).astype(np.int32)
    # observations the particles can actually produce: the golden script's (a, o) pairs
    pairs = {}
    for a, o in zip(g.a, g.o):
        pairs.setdefault(int(a), []).append(int(o))
    for t in range(T):
        for r in range(R):
            if int(acts[t, r]) not in pairs:
                acts[t, r] = int(g.a[0])
    obs = np.array([[pairs[int(acts[t, r])][rs.randint(len(pairs[int(acts[t, r])]))] for r in range(R)]
                    for t in range(T)], np.int32)
    active = np.ones((T, R), np.uint8)
    active[2, 1] = 0      # run 1 sits out step 2
    active[3, 0] = 0

    # --- the batch
    batch = fba.BatchedBAImportanceSampling(R, n)
    rng = fba.Rng.philox(seed)
    batch.initiate_sampled(sim, psid, pc, freq, rng, stride=pc.shape[1])
    b_lik, b_idx = [], []
    for t in range(T):
        if t == 3:
            batch.resetDomainStateDistribution(rng, active=active[t])
        b_lik.append(batch.updateEstimation(acts[t], obs[t], rng, active=active[t]))
        b_idx.append(batch.sample(rng, active=active[t]))
    got = [batch.download(r) for r in range(R)]
    copies = batch.copies()
    batch.free()

    # --- R stand-alone beliefs
    total_copies = 0
    for r in range(R):
        b = fba.BAImportanceSampling(n)
        rr = fba.Rng.philox(seed + r)
        b.initiate_sampled(sim, psid, pc, freq, rr, stride=pc.shape[1])
        for t in range(T):
            if not active[t, r]:
                rr.offset += (2 if t == 3 else 0) + 2 + 1    # the batch's calls advanced the offset anyway
                continue
            if t == 3:
                b.resetDomainStateDistribution(rr)
            lik = b.updateEstimation(int(acts[t, r]), int(obs[t, r]), rr)
            assert lik == b_lik[t][r], (name, r, t)
            assert b.sample(rr) == b_idx[t][r], (name, r, t)
        want = b.download()
        for k in ("state", "struct_id", "w"):
            np.testing.assert_array_equal(got[r][k], want[k], err_msg="%s run %d %s" % (name, r, k))
        np.testing.assert_array_equal(got[r]["counts"], want["counts"])
        total_copies += b.resample_stats()[0]
        b.free()
    assert copies == total_copies
    # inactive steps really were skipped
    assert np.isnan(b_lik[2][1]) and b_idx[2][1] == -1 and np.isnan(b_lik[3][0])
    sim.close()


def test_runs_do_not_mix(ctx):
    """Particles never cross a run boundary: every run is tagged through an otherwise unused count
    cell and keeps its tag through ten updates with different actions and observations per run."""
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    R, n = 64, 1024
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    batch = fba.BatchedBAImportanceSampling(R, n)
    rng = fba.Rng.philox(9)
    proto = g["is/init_counts"][:1]
    batch.initiate_sampled(sim, [0], proto, None, rng)
    L, h = batch.L, batch.storage
    # tag: T[open-left][s=0][s'] is never touched by listen-only scripts
    full = fba.beliefs._ParticleBelief._download(L, ctx, h, True)
    full["counts"][:, 0] = 1000.0 + np.repeat(np.arange(R), n)
    assert L.fba_belief_upload(h, 0, R * n, None, None, fba.capi.ptr(full["counts"]), None) == 0
    rs = np.random.RandomState(0)
    for t in range(10):
        lik = batch.updateEstimation(np.full(R, 2), rs.randint(0, 2, R), rng)
        assert np.all((lik > 0) & (lik <= 1))
    after = fba.beliefs._ParticleBelief._download(L, ctx, h, True)
    np.testing.assert_array_equal(after["counts"][:, 0], 1000.0 + np.repeat(np.arange(R), n))
    np.testing.assert_array_equal(after["w"], np.full(R * n, 1.0 / n))
    base = proto[0].astype(np.float64).sum() + 1000.0 - proto[0][0]
    sums = after["counts"].astype(np.float64).sum(1) - np.repeat(np.arange(R), n)
    np.testing.assert_array_equal(sums, np.full(R * n, base + 2 * 10))
    batch.free()
    sim.close()


def test_runs_argument_checks(ctx):
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    batch = fba.BatchedBAImportanceSampling(2, 64)
    rng = fba.Rng.philox(1)
    batch.initiate_sampled(sim, [0], g["is/init_counts"][:1], None, rng)
    with pytest.raises(fba.FbaError):
        batch.updateEstimation([2, 9], [0, 0], rng)
    with pytest.raises(fba.FbaError):
        batch.updateEstimation([2, 2], [0, 0], fba.Rng.replay(np.zeros(64, np.uint32)))
    # an out-of-range action of an INACTIVE run is ignored
    batch.updateEstimation([2, 9], [0, 0], rng, active=[1, 0])
    batch.free()
    with pytest.raises(fba.FbaError):
        fba.BatchedBAImportanceSampling(0, 64)
    sim.close()


@pytest.mark.parametrize("name,R,n", [("tiger", 6, 512), ("sysadmin3", 4, 300), ("gridworld3", 3, 256)])
def test_batched_planning_equals_stand_alone_search(ctx, name, R, n):
    """fba_runs_plan with one simulation per run per wave: every run's POMCP search is the sequential
    algorithm, and bit-identical — chosen action, root values, visit counts — to fba_tree_search
    (wave = 1) on a stand-alone belief seeded seed + r that went through the same calls."""
    import fba_pomdp_b200 as fba
    g = G.load(name)
    used, psid, pc, freq = _protos(g)
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par[used], g.o_par[used])
    seed, sims, u, disc = 77, 200, 5.0, 0.95
    a0, o0 = int(g.a[0]), int(g.o[0])
    depth = np.array([2 + (r % 3) for r in range(R)], np.int32)
    batch = fba.BatchedBAImportanceSampling(R, n)
    rng = fba.Rng.philox(seed)
    batch.initiate_sampled(sim, psid, pc, freq, rng, stride=pc.shape[1])
    batch.updateEstimation(np.full(R, a0), np.full(R, o0), rng)
    act, q, visits = batch.selectAction(sims, depth, u, disc, rng, sims_per_wave=1)
    act2, q2, visits2 = batch.selectAction(sims, depth, u, disc, rng, sims_per_wave=1)   # a second search: fresh trees
    batch.free()
    assert np.all(visits.sum(1) == sims) and np.all(visits2.sum(1) == sims)
    for r in range(R):
        b = fba.BAImportanceSampling(n)
        rr = fba.Rng.philox(seed + r)
        b.initiate_sampled(sim, psid, pc, freq, rr, stride=pc.shape[1])
        b.updateEstimation(a0, o0, rr)
        tree = fba.SearchTree(sim, sims, 8)
        for want_a, want_q, want_v in ((act, q, visits), (act2, q2, visits2)):
            a, qq, vv = tree.selectAction(b, sims, int(depth[r]), u, disc, 1, rr)
            np.testing.assert_array_equal(vv, want_v[r], err_msg="%s run %d visits" % (name, r))
            np.testing.assert_array_equal(qq, want_q[r], err_msg="%s run %d q" % (name, r))
            assert a == want_a[r]
        tree.free()
        b.free()
    sim.close()


def test_batched_planning_wide_waves_and_masks(ctx):
    """Several simulations per run per wave (atomics inside a run's tree) and an active mask: visit
    counts add up, parked runs are untouched, informed runs pick the right door."""
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    R, n = 32, 256
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    batch = fba.BatchedBAImportanceSampling(R, n)
    rng = fba.Rng.philox(3)
    batch.initiate_sampled(sim, [0], g["is/init_counts"][:1], None, rng)
    # put the tiger behind door r % 2 in every particle of run r
    st = np.repeat(np.arange(R) % 2, n).astype(np.int32)
    assert batch.L.fba_belief_upload(batch.storage, 0, R * n, fba.capi.ptr(st), None, None, None) == 0
    active = np.ones(R, np.uint8)
    active[5] = 0
    act, q, visits = batch.selectAction(512, 3, 20.0, 0.95, rng, sims_per_wave=16, active=active)
    assert act[5] == -1 and visits[5].sum() == 0
    for r in range(R):
        if r == 5:
            continue
        assert visits[r].sum() == 512
        assert act[r] == r % 2 and q[r, r % 2] == 10.0 and q[r, 1 - r % 2] == -100.0
    batch.free()
    sim.close()
