"""GPU: base+delta storage for tabular models (fba_model_desc.delta_capacity > 0) — particles share
the prior's dense tables and own only their list of increments. Same fixtures, same replay streams,
same bit-exact expectations as the dense storage; gridworld size 5 (720 KB dense per particle) is the
case it exists for."""
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


def delta_sim(ctx, g, cap=256):
    import fba_pomdp_b200 as fba
    desc = dict(g.desc, delta_capacity=cap)
    return fba.BAPOMDP(ctx, desc, g.t_par[:1], g.o_par[:1])


def delta_belief(cls, sim, g, prefix, **kw):
    counts = g[prefix + "_counts"]
    assert (counts == counts[0]).all(), "the tabular priors hand every particle the same tables"
    b = cls(len(counts), **kw)
    b.initiate(sim, proto_struct_id=[0], proto_counts=counts[:1], particle_proto=None, state=g[prefix + "_state"])
    return b


def sums(c):
    return c.astype(np.float64).sum(1)


@pytest.mark.parametrize("name", G.TABULAR)
def test_importance_sampling_replay_delta(ctx, name):
    import fba_pomdp_b200 as fba
    g = G.load(name)
    sim = delta_sim(ctx, g)
    b = delta_belief(fba.BAImportanceSampling, sim, g, "is/init")
    np.testing.assert_array_equal(b.download()["counts"], g["is/init_counts"])
    n_upd = 0
    for t in g.steps("is"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = fba.Rng.replay(g["is/%d/reset_words" % t])
            b.resetDomainStateDistribution(rng)
            assert rng.exhausted
            d = b.download()
            np.testing.assert_array_equal(d["state"], g["is/%d/reset_state" % t])
            np.testing.assert_array_equal(sums(d["counts"]), g["is/%d/reset_count_sums" % t])
        if fl & 1:
            continue
        rng = fba.Rng.replay(g["is/%d/update_words" % t])
        lik = b.update(a, o, rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["is/%d/state" % t])
        np.testing.assert_array_equal(d["w"], g["is/%d/w" % t])
        assert lik == float(g["is/%d/likelihood" % t])
        assert d["total_weight"] == float(g["is/%d/total_weight" % t])
        np.testing.assert_array_equal(sums(d["counts"]), g["is/%d/count_sums" % t])
        if g.has("is/%d/counts" % t):
            np.testing.assert_array_equal(d["counts"], g["is/%d/counts" % t])
        rng = fba.Rng.replay(g["is/%d/resample_words" % t])
        b.resample(rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["is/%d/rs_state" % t])
        np.testing.assert_array_equal(sums(d["counts"]), g["is/%d/rs_count_sums" % t])
        n_upd += 1
    assert n_upd >= 2
    d = b.download()
    np.testing.assert_array_equal(d["counts"], g["is/final_counts"])
    np.testing.assert_array_equal(d["state"], g["is/final_state"])

    # rollouts on the learned belief (KeepCounts), same stream as the dense fixture
    rng = fba.Rng.replay(g["roll/words"])
    ret = fba.rollouts(b, g["roll/particle"], g["roll/start"], g["roll/depth"], g.discount, rng,
                       g["roll/offsets"][:-1])
    np.testing.assert_array_equal(ret, g["roll/ret"])
    b.free()
    sim.close()


@pytest.mark.parametrize("name", [n for n in G.TABULAR if G.load(n).has("rs/init_counts")])
def test_rejection_sampling_replay_delta(ctx, name):
    import fba_pomdp_b200 as fba
    g = G.load(name)
    sim = delta_sim(ctx, g)
    b = delta_belief(fba.BARejectionSampling, sim, g, "rs/init")
    done = 0
    for t in g.steps("rs"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = fba.Rng.replay(g["rs/%d/reset_words" % t])
            b.resetDomainStateDistribution(rng)
            assert rng.exhausted
        if fl & 1 or not g.has("rs/%d/words" % t):
            continue
        rng = fba.Rng.replay(g["rs/%d/words" % t])
        b.updateEstimation(a, o, rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["rs/%d/state" % t])
        np.testing.assert_array_equal(sums(d["counts"]), g["rs/%d/count_sums" % t])
        done += 1
    assert done >= 1
    np.testing.assert_array_equal(b.download()["counts"], g["rs/final_counts"])
    b.free()
    sim.close()


def test_gridworld5_at_scale_native(ctx):
    """10^5 gridworld-5 particles: 72 GB dense, 100 MB as base+delta (capacity 256). PHILOX mode,
    in-place resampling: every particle's increments add up (2 per update), weights stay a
    distribution, rollouts run."""
    import fba_pomdp_b200 as fba
    g = G.load("gridworld5")
    n = 100_000
    sim = delta_sim(ctx, g)
    b = fba.BAImportanceSampling(n)
    rng = fba.Rng.philox(3)
    b.initiate_sampled(sim, [0], g["is/init_counts"][:1], None, rng)
    base = float(g["is/init_counts"][0].astype(np.float64).sum())
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    for t in range(5):
        lik = b.updateEstimation(*script[t], rng)
        assert 0.0 < lik <= 1.0
    d = b.download(counts=False)
    np.testing.assert_array_equal(d["w"], np.full(n, 1.0 / n))
    for i in (0, 17, n // 2, n - 1):
        blk = b.download(i, 1)["counts"][0].astype(np.float64)
        assert blk.sum() == base + 2.0 * 5
    ret = fba.rollouts(b, np.arange(4096) % n, np.zeros(4096, np.int32), np.full(4096, 20, np.int32), 0.95, rng)
    assert np.all(np.isfinite(ret)) and ret.min() >= 0.0
    b.free()
    sim.close()


def test_delta_capacity_overflow_is_reported(ctx):
    import fba_pomdp_b200 as fba
    from fba_pomdp_b200 import capi
    g = G.load("tiger")
    sim = delta_sim(ctx, g, cap=4)  # room for two updates
    ctx.set_option("auto_compact", 0)   # (by default a full list turns the belief into a dense one: test below)
    b = delta_belief(fba.BAImportanceSampling, sim, g, "is/init")
    words = np.random.RandomState(0).randint(0, 2**32, 6 * 1024, dtype=np.uint64).astype(np.uint32)
    for _ in range(2):
        b.updateEstimation(2, 0, fba.Rng.replay(words))
    with pytest.raises(fba.FbaError) as e:
        b.updateEstimation(2, 0, fba.Rng.replay(words))
    assert e.value.status == capi.ERR_CAPACITY
    ctx.set_option("auto_compact", 1)
    b.free()
    sim.close()


def test_full_delta_lists_continue_in_dense_storage(ctx):
    """auto_compact (default): when the increment lists are full the belief becomes a dense one in place
    (fba_belief_compact) and goes on — through the compaction it stays IDENTICAL, update after update, to a belief
    that was dense from the start (same Philox seed): likelihoods, states, weights, count blocks."""
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)][:7]
    out = []
    for cap in (0, 6):                                   # dense / room for three updates
        sim = fba.BAPOMDP(ctx, dict(g.desc, delta_capacity=cap), g.t_par, g.o_par)
        b = fba.BAImportanceSampling(4096)
        rng = fba.Rng.philox(11)
        b.initiate_sampled(sim, [0], g["is/init_counts"][:1], None, rng)
        caps, liks = [], []
        for a, o in script:
            liks.append(b.updateEstimation(a, o, rng))
            caps.append(b.L.fba_belief_delta_capacity(b.h))
        d = b.download()
        out.append((liks, d["state"], d["w"], d["counts"], caps))
        b.free()
        sim.close()
    assert out[1][4] == [6, 6, 6, 0, 0, 0, 0]            # the fourth update compacted first
    assert out[0][0] == out[1][0]
    for k in (1, 2, 3):
        np.testing.assert_array_equal(out[0][k], out[1][k])
