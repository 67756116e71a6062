"""CPU: the oracle's restatement of the reference's composite structure beliefs (oracle/pycomposite.py over
oracle/fba_oracle.c) against the UNMODIFIED reference's own classes — CheatingReinvigoration and
StructureIncubatorSampling — through tests/golden/composite.npz (oracle/gen_composite.py): fed the exact
mt19937 words each call consumed, every updateEstimation / resetDomainStateDistribution reproduces the
reference's filters bit for bit (states, structures, counts, weights, _total_weight, likelihood) and
consumes every word."""
import os

import numpy as np
import pytest

import pycomposite as PC
import pyoracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "composite.npz")
MUT_FACTORED_TIGER = 0


def load():
    g = np.load(GOLDEN)
    desc = {k[len("model/"):]: g[k] for k in g.files if k.startswith("model/")}
    m = O.Model(desc)
    st = O.Structs(m, g["structs/t_par"], g["structs/o_par"], cap=len(g["structs/t_par"]) + 512)
    return g, m, st


def used(rng):
    return not rng.overrun and rng.cur == len(rng.words)


def test_cheating_reinvigoration_equals_the_reference():
    g, m, st = load()
    stride = int(g["meta/stride"])
    b = PC.Cheating(m, st, PC.belief_from(g, "cheat/init_b", stride, True),
                    PC.belief_from(g, "cheat/init_c", stride, False), int(g["meta/amount"]),
                    float(g["meta/cheat_threshold"]))
    a_, o_, fl_ = g["script/a"], g["script/o"], g["script/flags"]
    updates = cheated = 0
    for t in range(len(a_)):
        if fl_[t] & 2 and t > 0:
            rng = O.Rng(g["cheat/%d/reset_words" % t])
            b.reset(rng)
            assert used(rng)
            np.testing.assert_array_equal(b.belief.state, g["cheat/%d/reset_b_state" % t])
            np.testing.assert_array_equal(b.correct.state, g["cheat/%d/reset_c_state" % t])
        if fl_[t] & 1:
            continue
        rng = O.Rng(g["cheat/%d/words" % t])
        b.update(int(a_[t]), int(o_[t]), rng)
        assert used(rng), t
        assert b.likelihood == float(g["cheat/%d/likelihood" % t])
        PC.assert_matches(b.belief, g, "cheat/%d/b" % t, weighted=True)
        PC.assert_matches(b.correct, g, "cheat/%d/c" % t)
        updates += 1
        cheated += int(b.likelihood == 1.0)
        if b.likelihood == 1.0:     # a cheat leaves _total_weight / N in the replaced slots: weights are no longer all equal
            assert len(np.unique(b.belief.w)) > 1
    PC.assert_matches(b.belief, g, "cheat/final_b", weighted=True)
    PC.assert_matches(b.correct, g, "cheat/final_c")
    assert updates >= 12 and cheated >= 3       # the cheat path (WeightedFilter::replace weights) was exercised


def test_structure_incubator_equals_the_reference():
    g, m, st = load()
    stride, N = int(g["meta/stride"]), int(g["meta/N"])
    inc = PC.Incubator(m, st, PC.belief_from(g, "inc/init_b", stride, False),
                       PC.belief_from(g, "inc/init_fc", stride, False),
                       PC.belief_from(g, "inc/init_s", stride, True), int(g["meta/amount"]),
                       float(g["meta/inc_threshold"]), MUT_FACTORED_TIGER)
    a_, o_, fl_ = g["script/a"], g["script/o"], g["script/flags"]
    done = 0
    for t in range(int(g["inc/last_step"]) + 1):
        if fl_[t] & 2 and t > 0:
            rng = O.Rng(g["inc/%d/reset_words" % t])
            inc.reset(rng)
            assert used(rng)
            for tag, b in (("b", inc.belief), ("fc", inc.fc), ("s", inc.shadow)):
                np.testing.assert_array_equal(b.state, g["inc/%d/reset_%s_state" % (t, tag)])
        if fl_[t] & 1:
            continue
        rng = O.Rng(g["inc/%d/words" % t])
        inc.update(int(a_[t]), int(o_[t]), rng)
        assert used(rng), t
        PC.assert_matches(inc.belief, g, "inc/%d/b" % t)
        PC.assert_matches(inc.fc, g, "inc/%d/fc" % t)
        PC.assert_matches(inc.shadow, g, "inc/%d/s" % t, weighted=True)
        done += 1
    assert done == 10


def test_incubator_parts_on_distinct_shadow_weights():
    """promotion into the belief, leastLikely and the replacement weights on NON-uniform shadow weights"""
    g, m, st = load()
    stride = int(g["meta/stride"])
    inc = PC.Incubator(m, st, PC.belief_from(g, "inc/parts/before_b", stride, False),
                       PC.belief_from(g, "inc/parts/before_fc", stride, False),
                       PC.belief_from(g, "inc/parts/before_s", stride, True), int(g["meta/amount"]),
                       float(g["meta/inc_threshold"]), MUT_FACTORED_TIGER)
    a, o = int(g["inc/parts/a"]), int(g["inc/parts/o"])
    steps = [lambda r: inc.reinvigorate_belief(r), lambda r: inc.reinvigorate_shadow(r),
             lambda r: inc.reject("belief", a, o, r), lambda r: inc.reject("fc", a, o, r),
             lambda r: O.is_update(m, st, inc.shadow, a, o, r),
             lambda r: setattr(inc, "shadow", O.is_resample(inc.shadow, r)[0])]
    for part, fn in enumerate(steps):
        rng = O.Rng(g["inc/parts/%d/words" % part])
        res = fn(rng)
        assert used(rng), part
        if part == 0:
            assert res == 5                      # the five heavy particles moved into the belief
        if part == 4:
            assert res == float(g["inc/parts/4/likelihood"])
        PC.assert_matches(inc.belief, g, "inc/parts/%d/b" % part)
        PC.assert_matches(inc.fc, g, "inc/parts/%d/fc" % part)
        PC.assert_matches(inc.shadow, g, "inc/parts/%d/s" % part, weighted=True)


@pytest.mark.parametrize("n,k", [(48, 6), (48, 47), (1000, 1), (1000, 128)])
def test_least_likely_ties_and_distinct_weights(n, k):
    """orc_least_likely restates libstdc++'s heap. On distinct weights the result does not depend on the
    container: WeightedFilter::leastLikely (WeightedFilter.cpp:206-243) seeds its queue with the first k
    particles and then walks ALL particles again from 0 — so one of the first k can be kept twice —
    replacing the heaviest kept one whenever a strictly lighter particle comes by; heaviest first."""
    rs = np.random.RandomState(n + k)
    w = rs.rand(n)
    kept = [(w[i], i) for i in range(k)]
    for i in range(n):
        top = max(kept)
        if w[i] < top[0]:
            kept.remove(top)
            kept.append((w[i], i))
    want = [i for _, i in sorted(kept, reverse=True)]
    np.testing.assert_array_equal(O.least_likely(w, k), want)
    # all equal: nothing is strictly lighter, the first k stay
    assert sorted(O.least_likely(np.full(n, 1.0 / n), k).tolist()) == list(range(k))


NESTED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nested.npz")


def load_nested(name):
    g = np.load(NESTED)
    P = name + "/"
    desc = {k[len(P + "model/"):]: g[k] for k in g.files if k.startswith(P + "model/")}
    m = O.Model(desc)
    st = O.Structs(m, g[P + "structs/t_par"], g[P + "structs/o_par"])
    counts = g[P + "init_counts"]
    top = O.Belief(counts.shape[0], counts.shape[1], True)
    top.counts[:] = counts
    top.struct_id[:] = g[P + "init_struct_id"]
    top.w[:] = g[P + "init_w"]
    top.total_weight = float(g[P + "init_total_weight"])
    return g, P, m, st, PC.Nested(m, st, top, g[P + "init_states"])


@pytest.mark.parametrize("name", ["tiger", "ftiger"])
def test_nested_belief_equals_the_reference(name):
    """NestedBelief (NestedBelief.cpp) on a tabular and a factored model: bottom filters, counts (raised by
    1 / n_bottom per accepted bottom particle), top weights and _total_weight, resets and sample() draws"""
    g, P, m, st, nb = load_nested(name)
    a_, o_, fl_ = g[P + "script/a"], g[P + "script/o"], g[P + "script/flags"]
    done = 0
    for t in range(len(a_)):
        if fl_[t] & 2 and t > 0:
            rng = O.Rng(g[P + "%d/reset_words" % t])
            nb.reset(rng)
            assert used(rng)
            np.testing.assert_array_equal(nb.states, g[P + "%d/reset_states" % t])
        if fl_[t] & 1:
            continue
        rng = O.Rng(g[P + "%d/words" % t])
        nb.update(int(a_[t]), int(o_[t]), rng)
        assert used(rng), t
        np.testing.assert_array_equal(nb.states, g[P + "%d/states" % t])
        np.testing.assert_array_equal(nb.top.counts, g[P + "%d/counts" % t])
        np.testing.assert_array_equal(nb.top.w, g[P + "%d/w" % t])
        assert nb.top.total_weight == float(g[P + "%d/total_weight" % t])
        rng = O.Rng(g[P + "%d/sample_words" % t])
        assert nb.sample(rng) == tuple(int(x) for x in g[P + "%d/sample" % t])
        assert used(rng)
        done += 1
    assert done >= 8
    assert len(np.unique(nb.top.w)) > 1      # the top weights moved apart (1 / attempts differs per particle)


GIBBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gibbs.npz")


def load_gibbs(tag):
    g = np.load(GIBBS)
    desc = {k[len("model/"):]: g[k] for k in g.files if k.startswith("model/")}
    m = O.Model(desc)
    n = g[tag + "/old_counts"].shape[0]
    old = O.Belief(n, g[tag + "/old_counts"].shape[1])
    old.counts[:], old.state[:], old.struct_id[:] = g[tag + "/old_counts"], g[tag + "/old_state"], g[tag + "/old_struct_id"]
    old.w[:] = g[tag + "/old_w"]
    old.total_weight = float(g[tag + "/old_total_weight"])
    return g, m, old


@pytest.mark.parametrize("tag", ["msg", "rs"])
def test_mh_within_gibbs_restated_over_oracle_primitives_equals_the_reference(tag):
    """MHwithinGibbs::reinvigorate (MHwithinGibbs.cpp:334-395) as a loop over oracle primitives — weighted draw,
    state history by backward messages / rejection sampling, computePosteriorCounts, the domain's mutate,
    LogBDScore, the accept test — fed the exact words the reference's private reinvigorate consumed: every word
    is used and the new belief equals the reference's bit for bit. Pins orc_state_history_msg (float tables,
    double messages, sequential sums), orc_state_history_rs and orc_add_history_counts."""
    g, m, old = load_gibbs(tag)
    rng = O.Rng(g[tag + "/words"])
    sid, state, counts = PC.gibbs_reinvigorate(
        m, old, g["priors/counts"], g["structs/t_par"], g["structs/o_par"],
        (g["history/len"], g["history/a"], g["history/o"]), rng, tag, g["model_state_prior"], MUT_FACTORED_TIGER,
        len(g[tag + "/new_state"]))
    assert used(rng)
    np.testing.assert_array_equal(sid, g[tag + "/new_struct_id"])
    np.testing.assert_array_equal(state, g[tag + "/new_state"])
    np.testing.assert_array_equal(counts[:, :g[tag + "/new_counts"].shape[1]], g[tag + "/new_counts"])
    assert len(np.unique(sid)) > 1


def test_flattened_model_rows_are_distributions():
    g, m, old = load_gibbs("msg")
    k = int(old.struct_id[0])
    T, Ob = O.flatten_model(m, g["structs/t_par"][k], g["structs/o_par"][k], old.counts[0])
    np.testing.assert_allclose(T.sum(2), 1.0, atol=1e-5)
    np.testing.assert_allclose(Ob.sum(2), 1.0, atol=1e-5)


MUTATE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mutate.npz")


@pytest.mark.parametrize("name", ["gridworld", "ftiger", "ca", "sysadmin"])
def test_domain_mutate_equals_the_reference(name):
    """FBAPOMDP::mutate of the four domains that have one, 48 chained calls each, against the reference's own
    results and word counts (tests/golden/mutate.npz) — the gridworld one (GridWorldBAPriors.cpp:200-225) cannot be
    reached through the reference's reinvigoration, so it is pinned here"""
    g = np.load(MUTATE)
    P = name + "/"
    m = O.Model({k[len(P + "model/"):]: g[k] for k in g.files if k.startswith(P + "model/")})
    rng = O.Rng(g[P + "words"])
    used_words = 0
    for k in range(len(g[P + "t_in"])):
        tp, op = O.mutate_structure(m, g[P + "t_in"][k], g[P + "o_in"][k], int(g[P + "kind"]), rng)
        used_words += int(g[P + "n_words"][k])
        assert rng.cur == used_words, k
        np.testing.assert_array_equal(tp, g[P + "t_out"][k])
        np.testing.assert_array_equal(op, g[P + "o_out"][k])
    assert used(rng)


def test_oracle_messages_and_rejection_sample_the_same_conditional():
    """a size-independent property of the two state-history samplers (MHwithinGibbs.cpp:38-213): both draw from
    p(states | model, actions, observations), so over 3000 draws each, from unrelated word streams, the frequency
    of every domain state at every position of the history agrees within 5 standard errors"""
    g, m, old = load_gibbs("msg")
    k = int(old.struct_id[0])
    tp, op = g["structs/t_par"][k], g["structs/o_par"][k]
    counts = np.zeros(g["priors/counts"].shape[1], np.float32)
    counts[:old.counts.shape[1]] = old.counts[0]
    hist = (g["history/len"][:3], g["history/a"][:int(g["history/len"][:3].sum())],
            g["history/o"][:int(g["history/len"][:3].sum())])
    n = 3000
    rs = np.random.RandomState(12)
    out = {}
    for method, words_per in (("msg", 64), ("rs", 20000)):
        rows = []
        for i in range(n):
            rng = O.Rng(rs.randint(0, 1 << 32, size=words_per, dtype=np.uint64).astype(np.uint32))
            rows.append(O.state_history(m, tp, op, counts, *hist, rng, method, g["model_state_prior"]))
        out[method] = np.stack(rows)
    for pos in range(out["msg"].shape[1]):
        pa = np.bincount(out["msg"][:, pos], minlength=m.S) / n
        pr = np.bincount(out["rs"][:, pos], minlength=m.S) / n
        p = (pa + pr) / 2
        assert np.all(np.abs(pa - pr) <= 5 * np.sqrt(2 * p * (1 - p) / n) + 1e-9), pos
