"""GPU: the Metropolis-Hastings structure belief's device pieces against the oracle,
which tests/test_oracle_mh.py pins to the reference's MHNIPS2018 through tests/golden/mh.npz.

fba_belief_replay_history = computePosterior (MHNIPS2018.cpp:41-109) on many proposal particles at once, in
REPLAY mode (proposal i draws from the i-th slice of the word stream): count blocks and final domain
states bit-identical to orc_mh_replay_history proposal by proposal, including the episode retries and
their -1 undo. fba_belief_assign_from moves accepted proposals into the new belief; LogBDScore of the
replayed proposals against their priors (the MH acceptance ratio's ingredients) agrees with the oracle."""
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


def test_history_replay_on_proposals_is_bit_exact(ctx):
    import fba_pomdp_b200 as fba
    import pyoracle as O
    g = G.load("mh")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par, max_structures=len(g.t_par))
    m = O.Model(g.desc)
    P, W = 96, 16384
    rs = np.random.RandomState(9)
    k = rs.randint(0, len(g.t_par), P).astype(np.int32)
    priors = g["priors/counts"]
    words = rs.randint(0, 2**32, size=P * W, dtype=np.uint64).astype(np.uint32)
    ln, ac, ob = g["history/len"], g["history/a"], g["history/o"]

    prop = fba.BARejectionSampling(P)                     # a flat (unweighted) set of proposal particles
    prop.initiate(sim, struct_id=k, counts=priors[k], state=np.zeros(P, np.int32))
    rng = fba.Rng.replay(words)
    prop.replay_history(ln, ac, ob, rng)
    d = prop.download()

    want_c, want_s, attempts = np.zeros_like(priors[k]), np.zeros(P, np.int32), []
    for i in range(P):
        c = priors[k[i]].copy()
        r = O.Rng(words[i * W:(i + 1) * W])
        n, last = O.mh_replay_history(m, g.t_par[k[i]], g.o_par[k[i]], c, ln, ac, ob, r)
        assert 0 < n and r.cur <= W
        want_c[i], want_s[i] = c, last
        attempts.append(n)
    np.testing.assert_array_equal(d["state"], want_s)
    np.testing.assert_array_equal(d["counts"][:, :want_c.shape[1]], want_c)
    assert max(attempts) > len(ln)                        # episodes were retried

    # the acceptance ratio's ingredients: LogBDScore(replayed proposal, its prior) (MHNIPS2018.cpp:238)
    prior = fba.BARejectionSampling(P)
    prior.initiate(sim, struct_id=k, counts=priors[k], state=np.zeros(P, np.int32))
    got = fba.log_bd_score(prop, prior)
    for i in range(0, P, 7):
        sz = m.struct_size(g.t_par[k[i]], g.o_par[k[i]])
        ref = O.log_bd_score(m, g.t_par[k[i]], g.o_par[k[i]], want_c[i][:sz].copy(), priors[k[i]][:sz].copy())
        assert abs(got[i] - ref) <= 1e-10 * abs(ref) + 1e-9

    # accepted proposals -> the new belief, in order (MHNIPS2018.cpp:241-246)
    take = np.array([5, 3, 3, 90, 0], np.int64)
    fresh = fba.BARejectionSampling(8)
    fresh.initiate(sim, struct_id=k[:8], counts=priors[k[:8]], state=np.zeros(8, np.int32))
    fresh.assign_from(2, prop, take)
    f = fresh.download()
    np.testing.assert_array_equal(f["counts"][2:7], d["counts"][take])
    np.testing.assert_array_equal(f["state"][2:7], d["state"][take])
    np.testing.assert_array_equal(f["struct_id"][2:7], d["struct_id"][take])
    np.testing.assert_array_equal(f["counts"][:2], priors[k[:2]][:, :f["counts"].shape[1]])
    with pytest.raises(fba.FbaError):
        fresh.assign_from(6, prop, take)                 # runs past the end
    with pytest.raises(fba.FbaError):
        prop.replay_history(ln, ac, ob, fba.Rng.replay(words[:P * 8]))   # slices too short: underrun reported
    # a history no model explains within the allowed attempts is reported, not looped on
    with pytest.raises(fba.FbaError):
        prop.replay_history([40], [2] * 40, [0, 1] * 20, fba.Rng.philox(3), max_attempts=3)
    for b in (prop, prior, fresh):
        b.free()
    sim.close()


def test_history_replay_philox_matches_the_oracle_statistically(ctx):
    """PHILOX mode (what the adapter uses): the distribution of the final tiger feature and the mean number of
    increments agree with the oracle's over many proposals (both sample the same posterior over hidden
    state paths given the history)."""
    import fba_pomdp_b200 as fba
    import pyoracle as O
    g = G.load("mh")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par, max_structures=len(g.t_par))
    m = O.Model(g.desc)
    P = 4096
    priors = g["priors/counts"]
    k = np.zeros(P, np.int32)
    prop = fba.BARejectionSampling(P)
    prop.initiate(sim, struct_id=k, counts=priors[k], state=np.zeros(P, np.int32))
    ln, ac, ob = g["history/len"], g["history/a"], g["history/o"]
    prop.replay_history(ln, ac, ob, fba.Rng.philox(77))
    d = prop.download()
    steps = int(ln.sum())
    J = sim.FS + sim.FO
    inc = (d["counts"][:, :priors.shape[1]].astype(np.float64) - priors[0].astype(np.float64)).sum(1)
    assert np.all(np.abs(inc - J * steps) < 0.01 * J * steps)
    rs = np.random.RandomState(2)
    ref_last = []
    for i in range(600):
        c = priors[0].copy()
        w = rs.randint(0, 2**32, size=16384, dtype=np.uint64).astype(np.uint32)
        ref_last.append(O.mh_replay_history(m, g.t_par[0], g.o_par[0], c, ln, ac, ob, O.Rng(w))[1])
    top_bit = int(np.log2(sim.S)) - 1                     # feature 0 (the tiger's side) is the most significant
    p_gpu = np.mean((d["state"] >> top_bit) & 1)
    p_ref = np.mean((np.array(ref_last) >> top_bit) & 1)
    se = np.sqrt(p_ref * (1 - p_ref) / 600 + p_gpu * (1 - p_gpu) / P)
    assert abs(p_gpu - p_ref) < 5 * se + 0.01, (p_gpu, p_ref, se)
    prop.free()
    sim.close()


def test_whole_mh_replays_the_reference(ctx):
    """MHNIPS2018::MH (MHNIPS2018.cpp:188-255) END TO END on the device bricks, proposal after proposal, fed the exact
    mt19937 words the reference's private MH consumed (tests/golden/mh.npz, 266 445 words): the weighted draw of the
    source particle, computePosterior (a single proposal consumes exactly the words it draws) and both BD scores run
    on the GPU; the keep-or-mutate coin, the domain's mutate and the accept uniform are drawn on the host from the
    same stream at the same cursor. Every word is used and the new belief — structures, domain states, count blocks —
    is the reference's bit for bit."""
    import math
    import fba_pomdp_b200 as fba
    import pyoracle as O
    from fba_pomdp_b200.capi import ptr
    g = G.load("mh")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par, max_structures=len(g.t_par))
    m = O.Model(g.desc)
    key = {(g.t_par[k].tobytes(), g.o_par[k].tobytes()): k for k in range(len(g.t_par))}
    priors = g["priors/counts"]
    stride = priors.shape[1]
    hist = (g["history/len"], g["history/a"], g["history/o"])
    words = g["mh/words"]
    rng, orng = fba.Rng.replay(words), O.Rng(words)

    def host(fn):                      # a host-side draw from the shared stream
        orng.c.cur = rng.cursor
        out = fn(orng)
        rng.cursor = orng.c.cur
        return out

    def single(k, counts):
        c = np.zeros((1, stride), np.float32)
        c[0, :len(counts)] = counts
        b = fba.BARejectionSampling(1)
        b.initiate(sim, struct_id=np.array([k], np.int32), counts=c, state=np.zeros(1, np.int32), stride=stride)
        return b

    n = g["old/counts"].shape[0]
    old = fba.BAImportanceSampling(n)
    old.initiate(sim, struct_id=g["old/struct_id"], counts=g["old/counts"], state=g["old/state"], stride=stride)
    w = np.ascontiguousarray(g["old/w"], np.float64)
    assert old.L.fba_belief_upload(old.h, 0, n, None, None, None, ptr(w)) == 0
    prior_of = {k: single(k, priors[k]) for k in range(len(g.t_par))}
    sid, state, counts = [], [], []
    proposals = 0
    while len(sid) < n:
        i = old.sample(rng)                                                              # :200
        k = int(g["old/struct_id"][i])
        if host(lambda r: O.lib().orc_boolean(r.ref())):                                 # :206-208
            k2 = k
        else:
            tp2, op2 = host(lambda r: O.mutate_structure(m, g.t_par[k], g.o_par[k], 0, r))
            k2 = key[(tp2.tobytes(), op2.tobytes())]
        prop = single(k2, priors[k2])
        prop.replay_history(*hist, rng)                                                  # :215, computePosterior
        src = single(k, g["old/counts"][i])
        old_score = fba.log_bd_score(src, prior_of[k])[0]                                # :237
        new_score = fba.log_bd_score(prop, prior_of[k2])[0]                              # :238
        proposals += 1
        if math.log(host(lambda r: O.lib().orc_uniform01(r.ref()))) < new_score - old_score:   # :240
            d = prop.download()
            sid.append(k2), state.append(int(d["state"][0])), counts.append(d["counts"][0])
        prop.free()
        src.free()
    assert rng.exhausted and proposals > n
    np.testing.assert_array_equal(sid, g["new/struct_id"])
    np.testing.assert_array_equal(state, g["new/state"])
    np.testing.assert_array_equal(np.stack(counts)[:, :g["new/counts"].shape[1]], g["new/counts"])
    for b in prior_of.values():
        b.free()
    old.free()
    sim.close()
