"""GPU: the device bricks of the reference's MHwithinGibbs structure belief
(src/beliefs/bayes-adaptive/factored/MHwithinGibbs.cpp) against the CPU oracle, whose restatement of the whole
reinvigorate() is pinned to the reference's private function bit for bit by tests/test_oracle_composite.py
(tests/golden/gibbs.npz):

  * fba_belief_sample_state_history, both methods (backward messages + forward sampling; rejection sampling), in
    REPLAY mode — particle i draws from the i-th slice of the stream, the oracle is fed the same slice — gives the
    oracle's state sequences state for state, on a small model (factored tiger, S = 16) and on sysadmin-10
    (S = 1024, 20 actions);
  * fba_belief_add_history_counts = computePosteriorCounts, bit-exact, per-particle and shared histories;
  * PHILOX mode: the two methods sample the same conditional distribution."""
import os

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu

GIBBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gibbs.npz")


@pytest.fixture(scope="module")
def env():
    import fba_pomdp_b200 as fba
    import pyoracle as O
    g = np.load(GIBBS)
    desc = {k[len("model/"):]: g[k] for k in g.files if k.startswith("model/")}
    ctx = fba.Context(0)
    sim = fba.BAPOMDP(ctx, desc, g["structs/t_par"], g["structs/o_par"])
    yield fba, O, g, O.Model(desc), sim
    sim.close()
    ctx.close()


def belief_of(fba, sim, g, tag="msg"):
    n, stride = g[tag + "/old_counts"].shape[0], g["priors/counts"].shape[1]
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, struct_id=g[tag + "/old_struct_id"], counts=g[tag + "/old_counts"], state=g[tag + "/old_state"],
               stride=stride)
    return b


@pytest.mark.parametrize("method", ["msg", "rs"])
def test_state_history_replay_is_bit_exact(env, method):
    fba, O, g, m, sim = env
    b = belief_of(fba, sim, g)
    n = b.size()
    hist = (g["history/len"], g["history/a"], g["history/o"])
    per = 60000 if method == "rs" else 2 * (int(hist[0].sum()) + len(hist[0])) + 7
    words = np.random.RandomState(9).randint(0, 1 << 32, size=per * n, dtype=np.uint64).astype(np.uint32)
    rng = fba.Rng.replay(words)
    got = b.sample_state_history(method, *hist, rng, state_prior=g["model_state_prior"])
    assert rng.cursor == per * n
    counts, sid = g["msg/old_counts"], g["msg/old_struct_id"]
    for i in range(n):
        k = int(sid[i])
        want = O.state_history(m, g["structs/t_par"][k], g["structs/o_par"][k], counts[i], *hist,
                               O.Rng(words[i * per:(i + 1) * per]), method, g["model_state_prior"])
        np.testing.assert_array_equal(got[i], want, err_msg="particle %d" % i)
    assert len({tuple(r) for r in got.tolist()}) > 1
    np.testing.assert_array_equal(b.download()["counts"][:, :counts.shape[1]], counts)     # counts untouched
    b.free()


def test_posterior_counts_are_bit_exact(env):
    fba, O, g, m, sim = env
    hist = (g["history/len"], g["history/a"], g["history/o"])
    L = int(hist[0].sum()) + len(hist[0])
    K = len(g["structs/t_par"])
    rs = np.random.RandomState(3)
    for shared in (False, True):
        b = fba.BARejectionSampling(K)           # one particle per structure, each holding that structure's prior
        b.initiate(sim, struct_id=np.arange(K, dtype=np.int32), counts=g["priors/counts"], state=np.zeros(K, np.int32))
        states = rs.randint(0, m.S, size=L if shared else (K, L)).astype(np.int32)
        b.add_history_counts(*hist, states)
        got = b.download()["counts"]
        for k in range(K):
            want = g["priors/counts"][k].copy()
            O.add_history_counts(m, g["structs/t_par"][k], g["structs/o_par"][k], want, *hist,
                                 states if shared else states[k])
            np.testing.assert_array_equal(got[k, :len(want)], want)
        assert got.sum() - g["priors/counts"].sum() == pytest.approx(K * int(hist[0].sum()) * (m.FS + m.FO), rel=1e-6)
        b.free()


@pytest.mark.parametrize("cluster", [1, 0])
def test_state_history_on_sysadmin_10(env, cluster):
    """S = 1024 states, 20 actions: the flattened tables are 80 MB per particle and a backward step is a 1024 x 1024
    matrix-vector product in the reference's summation order — with one model spread over a cluster of 4 CTAs
    (the default) and with one CTA per model: both bit-identical to the oracle"""
    fba, O, _, _, _ = env
    g = G.load("sysadmin")
    m = O.Model(g.desc)
    ctx = fba.Context(0)
    ctx.set_option("msg_cluster", cluster)
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    n = 3
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, struct_id=g["is/init_struct_id"][:n], counts=g["is/init_counts"][:n], state=g["is/init_state"][:n])
    upd = [t for t in range(len(g.a)) if not (g.flags[t] & 1)][:10]
    hist = (np.array([4, 6], np.int32), g.a[upd].astype(np.int32), g.o[upd].astype(np.int32))
    prior = np.zeros(m.S, np.float32)
    prior[m.S - 1] = 1.0                        # SysAdminFBAExtension.cpp:7-11: every computer starts working
    per = 2 * 12 + 3
    words = np.random.RandomState(4).randint(0, 1 << 32, size=per * n, dtype=np.uint64).astype(np.uint32)
    got = b.sample_state_history("msg", *hist, fba.Rng.replay(words), state_prior=prior)
    for i in range(n):
        k = int(g["is/init_struct_id"][i])
        want = O.state_history(m, g.t_par[k], g.o_par[k], g["is/init_counts"][i], *hist,
                               O.Rng(words[i * per:(i + 1) * per]), "msg", prior)
        np.testing.assert_array_equal(got[i], want)
    assert got[0][0] == m.S - 1 and got[0][5] == m.S - 1      # both episodes start from the prior's only state
    b.free()
    sim.close()
    ctx.close()


def test_messages_and_rejection_sample_the_same_conditional(env):
    """PHILOX: 4096 copies of one model; the state histories drawn by backward messages and by rejection sampling
    have the same distribution (per position, the frequency of every domain state within 5 standard errors)"""
    fba, O, g, m, sim = env
    n = 4096
    k = int(g["msg/old_struct_id"][0])
    stride = g["priors/counts"].shape[1]
    c = np.zeros((n, stride), np.float32)
    c[:, :g["msg/old_counts"].shape[1]] = g["msg/old_counts"][0]
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, struct_id=np.full(n, k, np.int32), counts=c, state=np.zeros(n, np.int32), stride=stride)
    hist = (g["history/len"], g["history/a"], g["history/o"])
    a = b.sample_state_history("msg", *hist, fba.Rng.philox(5), state_prior=g["model_state_prior"])
    r = b.sample_state_history("rs", *hist, fba.Rng.philox(6))
    assert a.shape == r.shape
    for pos in range(a.shape[1]):
        pa, pr = np.bincount(a[:, pos], minlength=m.S) / n, np.bincount(r[:, pos], minlength=m.S) / n
        p = (pa + pr) / 2
        assert np.all(np.abs(pa - pr) <= 5 * np.sqrt(2 * p * (1 - p) / n) + 1e-9), pos
    with pytest.raises(fba.FbaError):
        b.sample_state_history("rs", *hist, fba.Rng.philox(7), max_attempts=1)
    with pytest.raises(fba.FbaError):
        b.sample_state_history("msg", *hist, fba.Rng.philox(7))          # no state prior
    b.free()


@pytest.mark.parametrize("tag", ["msg", "rs"])
def test_whole_reinvigorate_replays_the_reference(env, tag):
    """MHwithinGibbs::reinvigorate (MHwithinGibbs.cpp:334-395) END TO END on the GPU bricks, one chain, fed the exact
    mt19937 words the reference's private reinvigorate consumed (tests/golden/gibbs.npz): the weighted draw of the
    start particle, every state history (a single model consumes exactly the words it draws), the posterior counts
    and the BD scores run on the device; only the domain's mutate and the accept uniform are drawn on the host, from
    the same stream at the same cursor. Every word is used and the new belief — structures, domain states, count
    blocks — is the reference's bit for bit."""
    import math
    fba, O, g, m, sim = env
    words = g[tag + "/words"]
    rng, orng = fba.Rng.replay(words), O.Rng(words)
    t_par, o_par, priors = g["structs/t_par"], g["structs/o_par"], g["priors/counts"]
    key = {(t_par[k].tobytes(), o_par[k].tobytes()): k for k in range(len(t_par))}
    hist = (g["history/len"], g["history/a"], g["history/o"])
    stride = priors.shape[1]

    def host(fn):                      # a host-side draw from the shared stream
        orng.c.cur = rng.cursor
        out = fn(orng)
        rng.cursor = orng.c.cur
        return out

    def single(k, counts):             # a one-particle belief holding (structure k, counts)
        c = np.zeros((1, stride), np.float32)
        c[0, :len(counts)] = counts
        b = fba.BARejectionSampling(1)
        b.initiate(sim, struct_id=np.array([k], np.int32), counts=c, state=np.zeros(1, np.int32), stride=stride)
        return b

    def posterior(k, seq):             # computePosteriorCounts on the prior model of structure k
        b = single(k, priors[k])
        b.add_history_counts(*hist, seq)
        return b

    def score(model, k):               # model.LogBDScore(prior_model)
        pb = single(k, priors[k])
        s = fba.log_bd_score(model, pb)[0]
        pb.free()
        return s

    def history(model):                # sampleStateHistory
        return model.sample_state_history(tag, *hist, rng, state_prior=g["model_state_prior"])[0]

    old = belief_of(fba, sim, g, tag)
    from fba_pomdp_b200.capi import ptr
    w = np.ascontiguousarray(g[tag + "/old_w"], np.float64)
    assert old.L.fba_belief_upload(old.h, 0, len(w), None, None, None, ptr(w)) == 0
    i = old.sample(rng)                                                        # :344
    k = int(g[tag + "/old_struct_id"][i])
    first = single(k, g[tag + "/old_counts"][i])
    seq = history(first)                                                       # :345-346
    first.free()
    model = posterior(k, seq)                                                  # :348-350
    sc = score(model, k)
    n_out = len(g[tag + "/new_state"])
    sid, state, counts = [], [], []
    proposals = 0
    while len(sid) < n_out:                                                    # :356
        tp2, op2 = host(lambda r: O.mutate_structure(m, t_par[k], o_par[k], 0, r))      # :360
        k2 = key[(tp2.tobytes(), op2.tobytes())]
        new_model = posterior(k2, seq)                                         # :362
        new_sc = score(new_model, k2)
        proposals += 1
        if math.log(host(lambda r: O.lib().orc_uniform01(r.ref()))) < new_sc - sc:      # :367
            sid.append(k2), state.append(int(seq[-1]))
            counts.append(new_model.download()["counts"][0])
            seq = history(model)                                               # :377-378: the model BEFORE the move
            model.free()
            k, model = k2, posterior(k2, seq)                                  # :381
            sc = score(model, k2)
        new_model.free()
    assert rng.exhausted and proposals > n_out
    np.testing.assert_array_equal(sid, g[tag + "/new_struct_id"])
    np.testing.assert_array_equal(state, g[tag + "/new_state"])
    np.testing.assert_array_equal(np.stack(counts)[:, :g[tag + "/new_counts"].shape[1]], g[tag + "/new_counts"])
    model.free()
    old.free()
