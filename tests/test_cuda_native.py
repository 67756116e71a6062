"""GPU: the production (PHILOX) path. No replay stream here, so parity is (a) exact where the
operation is deterministic given its inputs (offspring counts of systematic resampling, count
conservation, survivors never move) and (b) statistical against the CPU oracle driven by an
independent stream (posterior state marginals, step likelihood), with the tolerance written out."""
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


def _tiger_belief(ctx, n, fba):
    g = G.load("tiger")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    b = fba.BAImportanceSampling(n)
    proto = g["is/init_counts"][0]
    rs = np.random.RandomState(3)
    b.initiate(sim, proto_struct_id=[0], proto_counts=proto[None, :], particle_proto=None,
               state=rs.randint(0, 2, n).astype(np.int32))
    return g, sim, b


@pytest.mark.parametrize("inplace", [1, 0])
def test_systematic_resample_offspring_counts(ctx, inplace):
    """Each particle gets floor(N w) or ceil(N w) offspring; in place, survivors keep their slot."""
    import ctypes as C
    import fba_pomdp_b200 as fba
    ctx.set_option("inplace_resample", inplace)
    try:
        n = 4096
        g, sim, b = _tiger_belief(ctx, n, fba)
        rs = np.random.RandomState(5)
        # tag every particle: cell 0 carries its index (exact in float32)
        d = b.download()
        d["counts"][:, 0] = np.arange(n, dtype=np.float32)
        w = rs.gamma(0.3, size=n)
        w[rs.randint(0, n, 50)] = 0.0  # some exactly dead particles
        w /= w.sum()
        st = rs.randint(0, 2, n).astype(np.int32)
        rc = b.L.fba_belief_upload(b.h, 0, n, fba.capi.ptr(st), None, fba.capi.ptr(d["counts"]),
                                   fba.capi.ptr(w))
        assert rc == 0
        b.resample(fba.Rng.philox(11))
        out = b.download()
        ids = out["counts"][:, 0].astype(np.int64)
        cnt = np.bincount(ids, minlength=n)
        assert cnt.sum() == n
        expect = n * w
        assert np.all(cnt >= np.floor(expect - 1e-9)) and np.all(cnt <= np.ceil(expect + 1e-9))
        np.testing.assert_array_equal(out["state"], st[ids])  # the domain state travels with the block
        np.testing.assert_array_equal(out["w"], np.full(n, 1.0 / n))
        if inplace:
            alive = cnt > 0
            np.testing.assert_array_equal(ids[alive], np.arange(n)[alive])
        b.free()
        sim.close()
    finally:
        ctx.set_option("inplace_resample", 1)


def test_native_update_statistics_vs_oracle(ctx):
    """Tiger, listen: posterior over the tiger's location and the step likelihood from the CUDA
    PHILOX path agree with the CPU oracle run on an independent stream. N = 200k particles, so the
    standard error of a marginal is ~1e-3; tolerance 5e-3 (5 sigma)."""
    import fba_pomdp_b200 as fba
    import pyoracle as O
    n = 200_000
    g, sim, b = _tiger_belief(ctx, n, fba)
    start = b.download(counts=False)["state"].copy()
    rng = fba.Rng.philox(123)
    liks = [b.update(2, 0, rng)]
    post = b.download(counts=False)
    marg = np.array([post["w"][post["state"] == s].sum() for s in (0, 1)])
    b.resample(rng)
    liks.append(b.update(2, 0, rng))
    post2 = b.download(counts=False)
    marg2 = np.array([post2["w"][post2["state"] == s].sum() for s in (0, 1)])

    m = O.Model(g.desc)
    st = O.Structs(m, g.t_par, g.o_par)
    ob = O.Belief(n, 24)
    ob.counts[:] = g["is/init_counts"][0]
    ob.state[:] = start
    ob.total_weight = 1.0
    words = np.random.RandomState(99).randint(0, 2**32, size=20 * n, dtype=np.uint64).astype(np.uint32)
    orng = O.Rng(words)
    oliks = [O.is_update(m, st, ob, 2, 0, orng)]
    omarg = np.array([ob.w[ob.state == s].sum() for s in (0, 1)])
    np.testing.assert_allclose(marg, omarg, atol=5e-3)
    assert abs(liks[0] - oliks[0]) < 5e-3
    # second step: only the CUDA side continues (the oracle's multinomial resample is O(N^2));
    # listening twice to the same observation sharpens the posterior
    assert marg2.max() > marg.max() and abs(marg2.sum() - 1.0) < 1e-9
    assert 0.0 < liks[1] <= 1.0
    b.free()
    sim.close()


def test_count_conservation_and_weights_full_size(ctx):
    """BASELINE size for one GPU (sysadmin, 1.25e6 particles): size-independent properties.
    Every update adds exactly FS+FO = 11 to each particle's count block; weights are a distribution;
    resampling leaves uniform weights; domain states stay in range."""
    import fba_pomdp_b200 as fba
    g = G.load("sysadmin")
    n = 1_250_000
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    b = fba.BAImportanceSampling(n)
    proto = g["is/init_counts"][0]
    rng = fba.Rng.philox(7)
    b.initiate_sampled(sim, [0], proto[None, :], None, rng)
    base = float(proto.astype(np.float64).sum())
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    probe = np.array([0, 1, n // 3, n // 2, n - 2, n - 1])
    for t in range(4):
        a, o = script[t]
        lik = b.update(a, o, rng)
        assert 0.0 < lik <= 1.0
        d = b.download(counts=False)
        assert abs(d["w"].sum() - 1.0) < 1e-9 and d["w"].min() >= 0.0
        assert d["state"].min() >= 0 and d["state"].max() < sim.S
        b.resample(rng)
        d = b.download(counts=False)
        np.testing.assert_array_equal(d["w"], np.full(n, 1.0 / n))
        for i in probe:
            blk = b.download(int(i), 1)["counts"][0].astype(np.float64)
            assert blk.sum() == base + 11.0 * (t + 1), (t, i)
    b.free()
    sim.close()


def test_rejection_sampling_native_acceptance(ctx):
    """PHILOX rejection sampling: every accepted particle's count block grew by J, and the posterior
    over the tiger location after listening matches Bayes' rule (0.85 / 0.15 prior counts)."""
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    n = 50_000
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    b = fba.BARejectionSampling(n)
    proto = g["is/init_counts"][0]
    rs = np.random.RandomState(3)
    b.initiate(sim, proto_struct_id=[0], proto_counts=proto[None, :], particle_proto=None,
               state=rs.randint(0, 2, n).astype(np.int32))
    attempts = b.updateEstimation(2, 0, fba.Rng.philox(5))
    d = b.download()
    base = float(proto.astype(np.float64).sum())
    np.testing.assert_array_equal(d["counts"].astype(np.float64).sum(1), np.full(n, base + 2.0))
    frac0 = (d["state"] == 0).mean()
    assert abs(frac0 - 0.85) < 0.01  # s.e. ~ 0.0016
    assert 1.8 * n < attempts < 2.2 * n  # P(o) = 0.5
    b.free()
    sim.close()


def test_rollouts_native_mean_return(ctx):
    """PHILOX rollouts vs the CPU oracle on an independent stream: mean discounted return of 20k
    random-policy rollouts on the tiger prior agrees within 4 standard errors."""
    import fba_pomdp_b200 as fba
    import pyoracle as O
    g, sim, b = _tiger_belief(ctx, 1024, fba)
    n = 20_000
    rs = np.random.RandomState(1)
    pid = rs.randint(0, 1024, n)
    start = rs.randint(0, 2, n).astype(np.int32)
    depth = np.full(n, 10, np.int32)
    ours = fba.rollouts(b, pid, start, depth, 0.95, fba.Rng.philox(77))
    m = O.Model(g.desc)
    st = O.Structs(m, g.t_par, g.o_par)
    words = rs.randint(0, 2**32, size=4_000_000, dtype=np.uint64).astype(np.uint32)
    orng = O.Rng(words)
    proto = g["is/init_counts"][0].copy()
    ref = np.array([O.rollout(m, st.t_par[0], st.o_par[0], proto, int(s), 10, 0.95, orng)
                    for s in start[:5000]])
    se = np.sqrt(ours.var() / n + ref.var() / len(ref))
    assert abs(ours.mean() - ref.mean()) < 4 * se, (ours.mean(), ref.mean(), se)
    b.free()
    sim.close()


def test_shard_plan_matches_python(ctx):
    """fba_belief_shard_resample's quota allocation and exchange plan (C) equal the numpy
    restatement the gloo tests exercise; the shard resamples to exactly its quota."""
    import ctypes as C
    import fba_pomdp_b200 as fba
    n = 2048
    g, sim, b = _tiger_belief(ctx, n, fba)
    rs = np.random.RandomState(8)
    for trial in range(6):
        G_ = int(rs.randint(1, 9))
        rank = int(rs.randint(0, G_))
        rng = fba.Rng.philox(100 + trial)
        local = C.c_double(0)
        assert b.L.fba_belief_propose(b.h, 2, 0, C.byref(rng), C.byref(local)) == 0
        totals = rs.gamma(2.0, size=G_) * local.value
        totals[rank] = local.value
        u = float(rs.random_sample())
        plan = np.zeros((G_, G_), np.int64)
        tot = C.c_double(0)
        rc = b.L.fba_belief_shard_resample(b.h, fba.capi.ptr(np.ascontiguousarray(totals)), G_, rank, u,
                                           C.byref(rng), fba.capi.ptr(plan), C.byref(tot))
        assert rc == 0, b.L.fba_last_error(ctx.h)
        q = fba.offspring_quotas(totals, n * G_, u)
        np.testing.assert_array_equal(plan, fba.exchange_plan(q, n))
        assert abs(tot.value - totals.sum()) <= 1e-12 * totals.sum()
        assert b.L.fba_belief_export_count(b.h) == max(0, q[rank] - n)
        ctx.synchronize()
        # refill the dead slots locally so the next trial starts from a full shard
        d = b.download()
        rc = b.L.fba_belief_upload(b.h, 0, n, fba.capi.ptr(np.zeros(n, np.int32)), None,
                                   fba.capi.ptr(np.tile(g["is/init_counts"][0], (n, 1))),
                                   fba.capi.ptr(np.full(n, 1.0 / n)))
        assert rc == 0
    b.free()
    sim.close()


@pytest.mark.parametrize("name,n", [("ftiger_mu", 100_000), ("gridworld3", 1_000_000), ("ca", 1_000_000)])
def test_full_size_properties_other_configs(ctx, name, n):
    """BASELINE sizes of configs 2-4 (heterogeneous structures included): after every update the
    weights are a distribution, states are in range, probed count blocks grew by exactly FS+FO per
    update; after the in-place resample the weights are uniform and each structure's share of the
    belief only moved by what resampling can do (no particle lost or invented: structure ids stay
    within the prior's set)."""
    import fba_pomdp_b200 as fba
    g = G.load(name)
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    sid, counts = g["is/init_struct_id"], g["is/init_counts"]
    keys, psid, pc = {}, [], []
    for i in range(len(sid)):
        k = (int(sid[i]), counts[i].tobytes())
        if k not in keys:
            keys[k] = len(psid)
            psid.append(int(sid[i]))
            pc.append(counts[i])
    pc = np.stack(pc)
    base = {s: float(c.astype(np.float64).sum()) for s, c in zip(psid, pc)}
    assert len(set(psid)) == len(psid)  # one prototype per structure in these priors
    b = fba.BAImportanceSampling(n)
    rng = fba.Rng.philox(9)
    b.initiate_sampled(sim, np.array(psid, np.int32), pc, np.ones(len(psid)), rng, stride=pc.shape[1])
    J = sim.FS + sim.FO
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    probe = np.unique(np.random.RandomState(1).randint(0, n, 24))
    for t in range(3):
        a, o = script[t % len(script)]
        lik = b.update(a, o, rng)
        assert 0.0 < lik <= 1.0
        d = b.download(counts=False)
        assert abs(d["w"].sum() - 1.0) < 1e-9 and d["w"].min() >= 0.0
        assert d["state"].min() >= 0 and d["state"].max() < sim.S
        b.resample(rng)
        d = b.download(counts=False)
        np.testing.assert_array_equal(d["w"], np.full(n, 1.0 / n))
        assert set(np.unique(d["struct_id"])) <= set(psid)
        for i in probe:
            p = b.download(int(i), 1)
            assert p["counts"][0].astype(np.float64).sum() == base[int(p["struct_id"][0])] + J * (t + 1)
    b.free()
    sim.close()


@pytest.mark.parametrize("name", ["sysadmin", "ca", "gridworld3", "ftiger"])
def test_native_multi_step_statistics_vs_oracle(ctx, name):
    """The production path (PHILOX draws, in-place SYSTEMATIC resampling) against the CPU oracle
    (mt19937-style word stream, the reference's MULTINOMIAL resampling) over six consecutive belief
    updates of the domain's script: per-step likelihood and the posterior marginal of every state
    feature. GPU: 200 000 particles; oracle: 4 independent replicas of 3000 particles averaged
    (standard error of a marginal ~ 0.005 per replica set). Tolerances: 0.025 absolute on marginals,
    5 % relative on the step likelihood — systematic vs multinomial resampling are both unbiased, so
    the posteriors agree up to Monte-Carlo noise."""
    import fba_pomdp_b200 as fba
    import pyoracle as O
    g = G.load(name)
    m = O.Model(g.desc)
    st = O.Structs(m, g.t_par, g.o_par)
    sid0, counts0 = g["is/init_struct_id"], g["is/init_counts"]
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)][:6]
    fs = np.asarray(g.desc["feat_s"]).reshape(-1)
    steps_ = np.concatenate([np.cumprod(fs[::-1])[::-1][1:], [1]])

    def marginals(state, w):
        out = []
        for f in range(len(fs)):
            v = (state // steps_[f]) % fs[f]
            out.append(np.bincount(v, weights=w, minlength=fs[f]))
        return np.concatenate(out)

    # --- oracle replicas
    n_o, reps = 3000, 4
    o_lik = np.zeros((reps, len(script)))
    o_marg = [[] for _ in script]
    for r in range(reps):
        rs = np.random.RandomState(100 + r)
        idx = rs.randint(0, len(sid0), n_o)
        ob = O.Belief(n_o, counts0.shape[1])
        ob.counts[:], ob.struct_id[:] = counts0[idx], sid0[idx]
        words = rs.randint(0, 2**32, size=2_000_000, dtype=np.uint64).astype(np.uint32)
        orng = O.Rng(words)
        ob.state[:] = [m.sample_start_state(orng) for _ in range(n_o)]
        ob.total_weight = 1.0
        for t, (a, o) in enumerate(script):
            words = rs.randint(0, 2**32, size=2 * (m.FS + m.FO) * n_o + 2 * n_o + 8,
                               dtype=np.uint64).astype(np.uint32)
            orng = O.Rng(words)
            o_lik[r, t] = O.is_update(m, st, ob, a, o, orng)
            o_marg[t].append(marginals(ob.state, ob.w))
            ob, _ = O.is_resample(ob, orng)
    # --- CUDA
    n = 200_000
    keys, psid, pc = {}, [], []
    for i in range(len(sid0)):
        k = (int(sid0[i]), counts0[i].tobytes())
        if k not in keys:
            keys[k] = len(psid)
            psid.append(int(sid0[i]))
            pc.append(counts0[i])
    # prototype probabilities = their frequency in the reference prior's sample
    freq = np.bincount([keys[(int(sid0[i]), counts0[i].tobytes())] for i in range(len(sid0))],
                       minlength=len(psid)).astype(np.float64)
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    b = fba.BAImportanceSampling(n)
    rng = fba.Rng.philox(2024)
    b.initiate_sampled(sim, np.array(psid, np.int32), np.stack(pc), freq, rng, stride=counts0.shape[1])
    for t, (a, o) in enumerate(script):
        lik = b.update(a, o, rng)
        d = b.download(counts=False)
        want_lik = o_lik[:, t].mean()
        assert abs(lik - want_lik) <= 0.05 * want_lik + 1e-12, (t, lik, want_lik)
        np.testing.assert_allclose(marginals(d["state"], d["w"]), np.mean(o_marg[t], axis=0), atol=0.025)
        b.resample(rng)
    b.free()
    sim.close()
