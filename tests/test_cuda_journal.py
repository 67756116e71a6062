"""GPU: base+journal storage for FACTORED models (fba_model_desc.delta_capacity > 0 on a DBN model, VERDICT
r1 #5): particles share their prior prototype's count tables and own only the list of cells they
incremented, J = FS + FO per update. Same fixtures (generated from the unmodified reference), same replay
streams, same bit-exact expectations as dense storage — states, weights, likelihood, dense count blocks,
ancestor choices, rollout returns — including priors with many structures (collision avoidance,
factored tiger with a structure prior: one prototype per (structure, counts) pair)."""
import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu

FACTORED = [n for n in G.NAMES if n not in G.TABULAR]


@pytest.fixture(scope="module")
def ctx():
    import fba_pomdp_b200 as fba
    c = fba.Context(0)
    yield c
    c.close()


def sums(c):
    return c.astype(np.float64).sum(1)


def prototypes(g, prefix):
    sid, counts = g[prefix + "_struct_id"], g[prefix + "_counts"]
    seen, psid, pc, pp = {}, [], [], np.zeros(len(sid), np.int32)
    for i in range(len(sid)):
        k = (int(sid[i]), counts[i].tobytes())
        if k not in seen:
            seen[k] = len(psid)
            psid.append(int(sid[i]))
            pc.append(counts[i])
        pp[i] = seen[k]
    return np.array(psid, np.int32), np.stack(pc), pp


@pytest.mark.parametrize("name", FACTORED)
def test_importance_sampling_replay_journal(ctx, name):
    import fba_pomdp_b200 as fba
    g = G.load(name)
    J = len(g.desc["feat_s"]) + len(g.desc["feat_o"])
    n_upd_max = sum(1 for f in g.flags if not (f & 1))
    sim = fba.BAPOMDP(ctx, dict(g.desc, delta_capacity=J * (n_upd_max + 1)), g.t_par, g.o_par,
                      max_structures=len(g.t_par))
    psid, protos, pp = prototypes(g, "is/init")
    b = fba.BAImportanceSampling(len(pp))
    b.initiate(sim, proto_struct_id=psid, proto_counts=protos, particle_proto=pp, state=g["is/init_state"])
    d = b.download()
    np.testing.assert_array_equal(d["counts"][:, :g["is/init_counts"].shape[1]], g["is/init_counts"])
    np.testing.assert_array_equal(d["struct_id"], g["is/init_struct_id"])
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
            want = g["is/%d/counts" % t]
            np.testing.assert_array_equal(d["counts"][:, :want.shape[1]], want)
        rng = fba.Rng.replay(g["is/%d/resample_words" % t])
        b.resample(rng)
        assert rng.exhausted
        d = b.download()
        np.testing.assert_array_equal(d["state"], g["is/%d/rs_state" % t])
        np.testing.assert_array_equal(sums(d["counts"]), g["is/%d/rs_count_sums" % t])
        n_upd += 1
    assert n_upd >= 2
    d = b.download()
    want = g["is/final_counts"]
    np.testing.assert_array_equal(d["counts"][:, :want.shape[1]], want)
    np.testing.assert_array_equal(d["state"], g["is/final_state"])
    if g.has("roll/words"):
        rng = fba.Rng.replay(g["roll/words"])
        ret = fba.rollouts(b, g["roll/particle"], g["roll/start"], g["roll/depth"], g.discount, rng,
                           g["roll/offsets"][:-1])
        np.testing.assert_array_equal(ret, g["roll/ret"])
    # one more update than the journal holds: reported, not dropped silently
    b.free()
    sim.close()


def test_journal_overflow_is_reported(ctx):
    import fba_pomdp_b200 as fba
    g = G.load("sysadmin3")
    J = len(g.desc["feat_s"]) + len(g.desc["feat_o"])
    sim = fba.BAPOMDP(ctx, dict(g.desc, delta_capacity=2 * J), g.t_par, g.o_par, max_structures=len(g.t_par))
    psid, protos, pp = prototypes(g, "is/init")
    b = fba.BAImportanceSampling(len(pp))
    b.initiate(sim, proto_struct_id=psid, proto_counts=protos, particle_proto=pp, state=g["is/init_state"])
    rng = fba.Rng.philox(1)
    ctx.set_option("auto_compact", 0)           # (by default a full journal turns the belief into a dense one)
    b.updateEstimation(0, 0, rng)
    b.updateEstimation(0, 0, rng)
    with pytest.raises(fba.FbaError):
        b.updateEstimation(0, 0, rng)
        b.download()
    ctx.set_option("auto_compact", 1)
    b.free()
    sim.close()
    with pytest.raises(fba.FbaError):           # capacity must hold whole updates
        fba.BAPOMDP(ctx, dict(g.desc, delta_capacity=2 * J + 1), g.t_par, g.o_par)


def test_sysadmin_at_scale_journal_vs_dense_native(ctx):
    """10^5 sysadmin-10 particles, PHILOX mode, the same seed in journal and dense storage: the two
    beliefs stay IDENTICAL update after update (same Philox streams, same arithmetic) — states, weights,
    dense count views — through in-place resampling."""
    import fba_pomdp_b200 as fba
    g = G.load("sysadmin")
    J = len(g.desc["feat_s"]) + len(g.desc["feat_o"])
    n = 100_000
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    out = []
    # dense from the start / a journal that holds all 8 updates / a journal that is full after 5: the belief is
    # compacted into dense storage in place (fba_belief_compact, automatic) and goes on as a dense one
    for cap in (0, J * 16, J * 5):
        sim = fba.BAPOMDP(ctx, dict(g.desc, delta_capacity=cap), g.t_par, g.o_par)
        b = fba.BAImportanceSampling(n)
        rng = fba.Rng.philox(5)
        b.initiate_sampled(sim, [0], g["is/init_counts"][:1], None, rng)
        liks = [b.updateEstimation(a, o, rng) for a, o in script[:8]]
        d = b.download()
        out.append((liks, d["state"], d["w"], d["counts"][:, :g["is/init_counts"].shape[1]],
                    b.L.fba_belief_delta_capacity(b.h)))
        b.free()
        sim.close()
    assert [o[4] for o in out] == [0, J * 16, 0]
    for other in (1, 2):
        assert out[0][0] == out[other][0]
        for k in (1, 2, 3):
            np.testing.assert_array_equal(out[0][k], out[other][k])
