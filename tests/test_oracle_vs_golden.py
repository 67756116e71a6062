"""CPU: the plain-C oracle (oracle/fba_oracle.c) against fixtures generated from the unmodified
reference (oracle/gen_golden.py). This is what pins the oracle: same mt19937 words in, bit-identical
domain states / counts / weights out, and the same NUMBER of words consumed."""
import numpy as np
import pytest

import golden_util as G
import pyoracle as O


def make_model(g):
    m = O.Model(g.desc)
    st = O.Structs(m, g.t_par, g.o_par, cap=len(g.t_par) + 4096)
    return m, st


def belief_from(g, prefix, stride=None, weighted=True):
    counts = g[prefix + "_counts"]
    b = O.Belief(counts.shape[0], stride or counts.shape[1], weighted)
    b.counts[:, :counts.shape[1]] = counts
    b.state[:] = g[prefix + "_state"]
    b.struct_id[:] = g[prefix + "_struct_id"]
    if weighted:
        b.total_weight = O.sequential_uniform_total(b.N)
    return b


@pytest.mark.parametrize("name", G.NAMES)
def test_struct_sizes_cover_counts(name):
    g = G.load(name)
    m, st = make_model(g)
    sizes = st.sizes()
    c = g["is/init_counts"]
    sid = g["is/init_struct_id"]
    assert c.shape[1] == sizes[np.unique(sid)].max()
    # cells beyond a particle's own structure size are padding and stay zero
    for i in range(0, len(sid), max(1, len(sid) // 16)):
        assert not c[i, sizes[sid[i]]:].any()


@pytest.mark.parametrize("name", G.NAMES)
def test_domain_functor(name):
    g = G.load(name)
    m, _ = make_model(g)
    trip = g["functor/triples"]
    for (s, a, s2), r, t in zip(trip, g["functor/reward"], g["functor/terminal"]):
        rr, tt = m.reward(int(s), int(a), int(s2))
        assert rr == r and tt == bool(t), (s, a, s2)


@pytest.mark.parametrize("name", G.NAMES)
def test_start_state_draws(name):
    """BAImportanceSampling::initiate = N x sampleStartState: same words, same domain states."""
    g = G.load(name)
    m, _ = make_model(g)
    words = g["is/init_words"]
    want = g["is/init_state"]
    if name in ("ftiger_mu", "ca"):
        pytest.skip("prior draws a random structure per particle (host-side prior, out of scope)")
    rng = O.Rng(words)
    got = np.array([m.sample_start_state(rng) for _ in range(len(want))])
    assert not rng.overrun and rng.cur == len(words)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("name", G.NAMES)
def test_importance_sampling_replay(name):
    g = G.load(name)
    m, st = make_model(g)
    b = belief_from(g, "is/init")
    n_upd = 0
    for t in g.steps("is"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = O.Rng(g["is/%d/reset_words" % t])
            b, _ = O.is_reset_domain_states(m, b, rng)
            assert not rng.overrun and rng.cur == len(rng.words)
            np.testing.assert_array_equal(b.state, g["is/%d/reset_state" % t])
            np.testing.assert_array_equal(b.counts.astype(np.float64).sum(1),
                                          g["is/%d/reset_count_sums" % t])
        if fl & 1:
            continue
        rng = O.Rng(g["is/%d/update_words" % t])
        lik = O.is_update(m, st, b, a, o, rng)
        assert not rng.overrun and rng.cur == len(rng.words)
        assert lik == float(g["is/%d/likelihood" % t])
        np.testing.assert_array_equal(b.state, g["is/%d/state" % t])
        np.testing.assert_array_equal(b.w, g["is/%d/w" % t])
        assert b.total_weight == float(g["is/%d/total_weight" % t])
        np.testing.assert_array_equal(b.counts.astype(np.float64).sum(1), g["is/%d/count_sums" % t])
        if g.has("is/%d/counts" % t):
            np.testing.assert_array_equal(b.counts, g["is/%d/counts" % t])
        rng = O.Rng(g["is/%d/resample_words" % t])
        b, _ = O.is_resample(b, rng)
        assert not rng.overrun and rng.cur == len(rng.words)
        np.testing.assert_array_equal(b.state, g["is/%d/rs_state" % t])
        np.testing.assert_array_equal(b.struct_id, g["is/%d/rs_struct_id" % t])
        np.testing.assert_array_equal(b.counts.astype(np.float64).sum(1),
                                      g["is/%d/rs_count_sums" % t])
        assert b.total_weight == float(g["is/%d/rs_total_weight" % t])
        n_upd += 1
    assert n_upd >= 2
    np.testing.assert_array_equal(b.counts, g["is/final_counts"])
    np.testing.assert_array_equal(b.state, g["is/final_state"])


@pytest.mark.parametrize("name", G.NAMES)
def test_rollouts_replay(name):
    g = G.load(name)
    m, st = make_model(g)
    counts, sid = g["is/final_counts"], g["is/final_struct_id"]
    offs, words = g["roll/offsets"], g["roll/words"]
    for i, (p, s0, d) in enumerate(zip(g["roll/particle"], g["roll/start"], g["roll/depth"])):
        rng = O.Rng(words[offs[i]:offs[i + 1]])
        ret = O.rollout(m, st.t_par[sid[p]], st.o_par[sid[p]], counts[p].copy(), int(s0), int(d),
                        g.discount, rng)
        assert not rng.overrun and rng.cur == len(rng.words), i
        assert ret == g["roll/ret"][i], i


@pytest.mark.parametrize("name", [n for n in G.NAMES if G.load(n).has("rs/init_counts")])
def test_rejection_sampling_replay(name):
    g = G.load(name)
    m, st = make_model(g)
    b = belief_from(g, "rs/init", weighted=False)
    done = 0
    for t in g.steps("rs"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = O.Rng(g["rs/%d/reset_words" % t])
            O.flat_reset_domain_states(m, b, rng)
            assert not rng.overrun and rng.cur == len(rng.words)
            np.testing.assert_array_equal(b.state, g["rs/%d/reset_state" % t])
        if fl & 1 or not g.has("rs/%d/words" % t):
            continue
        rng = O.Rng(g["rs/%d/words" % t])
        b, _, attempts = O.reject_sample(m, st, b, a, o, rng)
        assert not rng.overrun and rng.cur == len(rng.words)
        np.testing.assert_array_equal(b.state, g["rs/%d/state" % t])
        np.testing.assert_array_equal(b.struct_id, g["rs/%d/struct_id" % t])
        np.testing.assert_array_equal(b.counts.astype(np.float64).sum(1), g["rs/%d/count_sums" % t])
        done += 1
    assert done >= 1
    np.testing.assert_array_equal(b.counts, g["rs/final_counts"])


def _remap(st, g, sid_ref):
    """Structure ids are table positions; map fixture ids -> oracle ids through the masks."""
    key = {(st.t_par[i].tobytes(), st.o_par[i].tobytes()): i for i in range(st.n)}
    return np.array([key[(g.t_par[j].tobytes(), g.o_par[j].tobytes())] for j in sid_ref], np.int32)


@pytest.mark.parametrize("name", [n for n in G.NAMES if G.load(n).has("reinv/K")])
def test_reinvigoration_replay(name):
    g = G.load(name)
    m, st = make_model(g)
    stride = int(g["reinv/stride"])
    K = int(g["reinv/K"])
    b = belief_from(g, "reinv/init_b", stride, weighted=False)
    fc = belief_from(g, "reinv/init_fc", stride, weighted=False)
    done = 0
    for t in g.steps("reinv"):
        a, o, fl = int(g.a[t]), int(g.o[t]), int(g.flags[t])
        if fl & 2 and t > 0:
            rng = O.Rng(g["reinv/%d/reset_words" % t])
            O.flat_reset_domain_states(m, b, rng)
            O.flat_reset_domain_states(m, fc, rng)
            assert not rng.overrun and rng.cur == len(rng.words)
            np.testing.assert_array_equal(b.state, g["reinv/%d/reset_b_state" % t])
            np.testing.assert_array_equal(fc.state, g["reinv/%d/reset_fc_state" % t])
        if fl & 1 or not g.has("reinv/%d/breed_words" % t):
            continue
        rng = O.Rng(g["reinv/%d/breed_words" % t])
        O.reinvigorate(m, st, b, fc, K, G.MUTATE_KIND[name], rng)
        assert not rng.overrun and rng.cur == len(rng.words)
        np.testing.assert_array_equal(b.state, g["reinv/%d/breed_b_state" % t])
        np.testing.assert_array_equal(b.struct_id, _remap(st, g, g["reinv/%d/breed_b_struct_id" % t]))
        np.testing.assert_array_equal(b.counts, g["reinv/%d/breed_b_counts" % t])
        rng = O.Rng(g["reinv/%d/reject_words" % t])
        b, _, _ = O.reject_sample(m, st, b, a, o, rng)
        fc, _, _ = O.reject_sample(m, st, fc, a, o, rng)
        assert not rng.overrun and rng.cur == len(rng.words)
        for tag, bel in (("b", b), ("fc", fc)):
            np.testing.assert_array_equal(bel.state, g["reinv/%d/%s_state" % (t, tag)])
            np.testing.assert_array_equal(bel.counts.astype(np.float64).sum(1),
                                          g["reinv/%d/%s_count_sums" % (t, tag)])
        done += 1
    assert done >= 1
    np.testing.assert_array_equal(b.counts, g["reinv/final_b_counts"])
    np.testing.assert_array_equal(fc.counts, g["reinv/final_fc_counts"])


@pytest.mark.parametrize("name", ["ftiger", "sysadmin3", "sysadmin", "gridworld3_fba"])
def test_log_bd_score_bit_exact(name):
    """orc_log_bd_score against BABNModel::LogBDScore of the unmodified reference
    (oracle/gen_bd_score.py -> tests/golden/bd_score.npz): same libm, same accumulation order,
    identical doubles."""
    import os
    import pyoracle as O
    z = np.load(os.path.join(G.GOLDEN_DIR, "bd_score.npz"))
    g = G.load(name)
    m = O.Model(g.desc)
    for i in range(len(z[name + "/score"])):
        got = O.log_bd_score(m, z[name + "/t_par"], z[name + "/o_par"], z[name + "/counts"][i], z[name + "/prior"])
        assert got == z[name + "/score"][i], (name, i, got, z[name + "/score"][i])


@pytest.mark.parametrize("n", [1, 2, 3, 1000, 16384])
def test_fast_weighted_sampling_equals_the_reference_scan(n):
    """orc_weighted_sample_many (one sequential pass for the remainders + binary search per draw) picks
    exactly the indices WeightedFilter::sample's O(n) scan picks (orc_weighted_sample, pinned against the
    reference by the golden fixtures), draw by draw, on the same word stream — including weights with
    zeros, a huge dynamic range, and thresholds that fall exactly on a remainder."""
    import pyoracle as O
    rs = np.random.RandomState(n)
    for kind in range(4):
        w = rs.random_sample(n)
        if kind == 1:
            w[rs.random_sample(n) < 0.5] = 0.0
            w[0] = max(w[0], 1e-3)
        elif kind == 2:
            w = np.exp(rs.normal(0, 12, n))
        elif kind == 3:
            w = np.full(n, 1.0 / n)
        total = O.normalize(w, float(np.cumsum(w)[-1]))
        b = O.Belief(n, 4)
        b.w[:] = w
        b.total_weight = total
        words = rs.randint(0, 2**32, size=2 * 600 + 8, dtype=np.uint64).astype(np.uint32)
        if kind == 3 and n > 2:
            words[:4] = [0, 0, 0, 1 << 31]          # u = 0 and u = 0.5: thresholds on exact remainders
        r1, r2 = O.Rng(words), O.Rng(words)
        slow = np.array([O.weighted_sample(b, r1) for _ in range(600)])
        fast = O.weighted_sample_many(b.w, total, r2, 600)
        np.testing.assert_array_equal(slow, fast)
        assert r1.cur == r2.cur
