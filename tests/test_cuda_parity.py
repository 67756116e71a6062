"""GPU: the CUDA path (through the C ABI) against the golden fixtures generated from the unmodified
reference, in replay mode: same mt19937 words in; bit-identical domain states, count blocks and
ancestor choices out; weights within 1e-5 relative (in fact bit-identical here); and the same
number of words consumed."""
import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu

REL_TOL_W = 1e-5  # BASELINE.json north_star: "log-weights agree to 1e-5 relative"


@pytest.fixture(scope="module")
def ctx():
    import fba_pomdp_b200 as fba
    c = fba.Context(0)
    yield c
    c.close()


def make_sim(ctx, g, extra_structs=0):
    import fba_pomdp_b200 as fba
    return fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par, max_structures=len(g.t_par) + extra_structs)


def count_sums(c):
    return c.astype(np.float64).sum(1)


def assert_counts_equal(got, want):
    """The belief's stride is the largest structure registered in the model, which may exceed the
    fixture's (e.g. when the structure table also holds the fully connected structures)."""
    np.testing.assert_array_equal(got[:, :want.shape[1]], want)
    assert not got[:, want.shape[1]:].any()


@pytest.mark.parametrize("name", G.NAMES)
def test_importance_sampling_replay(ctx, name):
    import fba_pomdp_b200 as fba
    g = G.load(name)
    sim = make_sim(ctx, g)
    b = fba.BAImportanceSampling(len(g["is/init_state"]))
    b.initiate(sim, struct_id=g["is/init_struct_id"], counts=g["is/init_counts"], state=g["is/init_state"])
    n_upd = 0
    for t in g.steps("is"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = fba.Rng.replay(g["is/%d/reset_words" % t])
            b.resetDomainStateDistribution(rng)
            assert rng.exhausted
            d = b.download()
            np.testing.assert_array_equal(d["state"], g["is/%d/reset_state" % t])
            np.testing.assert_array_equal(count_sums(d["counts"]), g["is/%d/reset_count_sums" % t])
        if fl & 1:
            continue
        rng = fba.Rng.replay(g["is/%d/update_words" % t])
        lik = b.update(a, o, rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["is/%d/state" % t])
        np.testing.assert_allclose(d["w"], g["is/%d/w" % t], rtol=REL_TOL_W, atol=0)
        np.testing.assert_array_equal(d["w"], g["is/%d/w" % t])  # and in fact bit-identical
        assert lik == float(g["is/%d/likelihood" % t])
        assert d["total_weight"] == float(g["is/%d/total_weight" % t])
        np.testing.assert_array_equal(count_sums(d["counts"]), g["is/%d/count_sums" % t])
        if g.has("is/%d/counts" % t):
            assert_counts_equal(d["counts"], g["is/%d/counts" % t])
        rng = fba.Rng.replay(g["is/%d/resample_words" % t])
        b.resample(rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["is/%d/rs_state" % t])
        np.testing.assert_array_equal(d["struct_id"], g["is/%d/rs_struct_id" % t])
        np.testing.assert_array_equal(count_sums(d["counts"]), g["is/%d/rs_count_sums" % t])
        assert d["total_weight"] == float(g["is/%d/rs_total_weight" % t])
        n_upd += 1
    assert n_upd >= 2
    d = b.download()
    assert_counts_equal(d["counts"], g["is/final_counts"])
    np.testing.assert_array_equal(d["state"], g["is/final_state"])
    b.free()
    sim.close()


@pytest.mark.parametrize("coop", [0, 1])
@pytest.mark.parametrize("name", G.NAMES)
def test_rollouts_replay(ctx, name, coop):
    """Both rollout kernels (thread per rollout / warp per rollout with cooperative row loads)."""
    import fba_pomdp_b200 as fba
    g = G.load(name)
    ctx.set_option("rollout_coop", coop)
    sim = make_sim(ctx, g)
    b = fba.BAImportanceSampling(len(g["is/final_state"]))
    b.initiate(sim, struct_id=g["is/final_struct_id"], counts=g["is/final_counts"], state=g["is/final_state"])
    rng = fba.Rng.replay(g["roll/words"])
    ret = fba.rollouts(b, g["roll/particle"], g["roll/start"], g["roll/depth"], g.discount, rng,
                       g["roll/offsets"][:-1])
    np.testing.assert_array_equal(ret, g["roll/ret"])
    # rollouts are KeepCounts: the belief is untouched
    assert_counts_equal(b.download()["counts"], g["is/final_counts"])
    ctx.set_option("rollout_coop", -1)
    b.free()
    sim.close()


@pytest.mark.parametrize("name", [n for n in G.NAMES if G.load(n).has("rs/init_counts")])
def test_rejection_sampling_replay(ctx, name):
    import fba_pomdp_b200 as fba
    g = G.load(name)
    sim = make_sim(ctx, g)
    b = fba.BARejectionSampling(len(g["rs/init_state"]))
    b.initiate(sim, struct_id=g["rs/init_struct_id"], counts=g["rs/init_counts"], state=g["rs/init_state"])
    done = 0
    for t in g.steps("rs"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = fba.Rng.replay(g["rs/%d/reset_words" % t])
            b.resetDomainStateDistribution(rng)
            assert rng.exhausted
            np.testing.assert_array_equal(b.download(counts=False)["state"], g["rs/%d/reset_state" % t])
        if fl & 1 or not g.has("rs/%d/words" % t):
            continue
        rng = fba.Rng.replay(g["rs/%d/words" % t])
        b.updateEstimation(a, o, rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["rs/%d/state" % t])
        np.testing.assert_array_equal(d["struct_id"], g["rs/%d/struct_id" % t])
        np.testing.assert_array_equal(count_sums(d["counts"]), g["rs/%d/count_sums" % t])
        done += 1
    assert done >= 1
    assert_counts_equal(b.download()["counts"], g["rs/final_counts"])
    b.free()
    sim.close()


def _struct_key_map(sim, g):
    key = {}
    for i in range(sim.num_structures):
        t, o = sim.structure(i)
        key[(t.tobytes(), o.tobytes())] = i
    return key


@pytest.mark.parametrize("name", [n for n in G.NAMES if G.load(n).has("reinv/K")])
def test_reinvigoration_replay(ctx, name):
    import fba_pomdp_b200 as fba
    g = G.load(name)
    sim = make_sim(ctx, g, extra_structs=4096)
    stride = int(g["reinv/stride"])
    K = int(g["reinv/K"])
    b = fba.ReinvigoratingRejectionSampling(len(g["reinv/init_b_state"]), K, G.MUTATE_KIND[name])
    b.initiate(sim, stride=stride,
               belief=dict(struct_id=g["reinv/init_b_struct_id"], counts=g["reinv/init_b_counts"],
                           state=g["reinv/init_b_state"]),
               fully_connected=dict(struct_id=g["reinv/init_fc_struct_id"],
                                    counts=g["reinv/init_fc_counts"], state=g["reinv/init_fc_state"]))
    done = 0
    for t in g.steps("reinv"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = fba.Rng.replay(g["reinv/%d/reset_words" % t])
            b.resetDomainStateDistribution(rng)
            assert rng.exhausted
            np.testing.assert_array_equal(b.download(counts=False)["state"], g["reinv/%d/reset_b_state" % t])
            np.testing.assert_array_equal(b.download_fully_connected(False)["state"],
                                          g["reinv/%d/reset_fc_state" % t])
        if fl & 1 or not g.has("reinv/%d/breed_words" % t):
            continue
        rng = fba.Rng.replay(g["reinv/%d/breed_words" % t])
        b.reinvigorateParticles(rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["reinv/%d/breed_b_state" % t])
        np.testing.assert_array_equal(d["counts"], g["reinv/%d/breed_b_counts" % t])
        key = _struct_key_map(sim, g)
        want = np.array([key[(g.t_par[j].tobytes(), g.o_par[j].tobytes())]
                         for j in g["reinv/%d/breed_b_struct_id" % t]])
        np.testing.assert_array_equal(d["struct_id"], want)
        # the two rejectSample calls share one stream
        rng = fba.Rng.replay(g["reinv/%d/reject_words" % t])
        import ctypes as C
        n = C.c_int64(0)
        for h in (b.h, b.fc):
            rc = b.L.fba_belief_reject_sample(h, a, o, C.byref(rng), C.byref(n))
            assert rc == 0, b.L.fba_last_error(ctx.h)
        assert rng.exhausted
        for tag, d in (("b", b.download()), ("fc", b.download_fully_connected())):
            np.testing.assert_array_equal(d["state"], g["reinv/%d/%s_state" % (t, tag)])
            np.testing.assert_array_equal(count_sums(d["counts"]), g["reinv/%d/%s_count_sums" % (t, tag)])
        done += 1
    assert done >= 1
    np.testing.assert_array_equal(b.download()["counts"], g["reinv/final_b_counts"])
    np.testing.assert_array_equal(b.download_fully_connected()["counts"], g["reinv/final_fc_counts"])
    b.free()
    sim.close()


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


@pytest.mark.parametrize("name,n", [("tiger", 5000), ("tiger", 20000), ("sysadmin", 6000), ("ca", 4100),
                                    ("gridworld3", 3000)])
def test_replay_vs_oracle_beyond_fixture_sizes(ctx, name, n):
    """CUDA vs the CPU oracle on seeded streams at particle counts the fixtures do not reach — in
    particular several chunks of the sequential weight chains (k_seq_normalize stages 2048 weights at
    a time) and heterogeneous structures. Three update+resample rounds: bit-identical states,
    weights, likelihood, _total_weight and count blocks."""
    import fba_pomdp_b200 as fba
    import pyoracle as O
    g = G.load(name)
    m = O.Model(g.desc)
    st = O.Structs(m, g.t_par, g.o_par)
    rs = np.random.RandomState(n)
    idx = rs.randint(0, len(g["is/init_state"]), n)
    counts0 = g["is/init_counts"][idx]
    sid0 = g["is/init_struct_id"][idx]
    state0 = rs.randint(0, m.S, n).astype(np.int32)

    ob = O.Belief(n, counts0.shape[1])
    ob.counts[:], ob.state[:], ob.struct_id[:] = counts0, state0, sid0
    ob.total_weight = O.sequential_uniform_total(n)

    sim = make_sim(ctx, g)
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, struct_id=sid0, counts=counts0, state=state0, stride=counts0.shape[1])
    J = m.FS + m.FO
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    for t in range(3):
        a, o = script[t % len(script)]
        words = rs.randint(0, 2**32, size=2 * J * n + 2 * n, dtype=np.uint64).astype(np.uint32)
        orng = O.Rng(words)
        lik_ref = O.is_update(m, st, ob, a, o, orng)
        rng = fba.Rng.replay(words)
        lik = b.update(a, o, rng)
        assert rng.cursor == orng.cur
        d = b.download()
        assert lik == lik_ref and d["total_weight"] == ob.total_weight
        np.testing.assert_array_equal(d["state"], ob.state)
        np.testing.assert_array_equal(d["w"], ob.w)
        assert_counts_equal(d["counts"], ob.counts)
        ob, anc = O.is_resample(ob, orng)
        b.resample(rng)
        assert rng.cursor == orng.cur and rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], ob.state)
        np.testing.assert_array_equal(d["struct_id"], ob.struct_id)
        assert_counts_equal(d["counts"], ob.counts)
        assert d["total_weight"] == ob.total_weight
    b.free()
    sim.close()
