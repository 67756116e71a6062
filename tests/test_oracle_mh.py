"""CPU: the oracle's pieces of the reference's Metropolis-Hastings structure belief
(src/beliefs/bayes-adaptive/factored/MHNIPS2018.cpp) against tests/golden/mh.npz, a fixture oracle/gen_mh.py
produced by running the unmodified reference's MHNIPS2018 under seed "42".

MHNIPS2018::MH (MHNIPS2018.cpp:188-255) is restated here, in the test, as a loop over oracle primitives —
weighted draw, boolean, the domain's mutate, computePosterior (orc_mh_replay_history), LogBDScore, the
accept test — fed with the exact mt19937 words the reference's MH consumed. It must consume ALL of them
and produce the reference's new belief bit for bit: that pins orc_mh_replay_history (episode retries,
the -1 undo, start-state draws), orc_mutate_structure and orc_log_bd_score in the order MH uses them."""
import math

import numpy as np
import pytest

import golden_util as G

MUT_FACTORED_TIGER = 0


def load():
    import pyoracle as O
    g = G.load("mh")
    m = O.Model(g.desc)
    key = {(g.t_par[k].tobytes(), g.o_par[k].tobytes()): k for k in range(len(g.t_par))}
    return O, g, m, key


def mh(O, g, m, key, words):
    """-> (struct ids, states, counts) of the new belief, words consumed"""
    n, stride = g["old/counts"].shape[0], g["priors/counts"].shape[1]
    old = O.Belief(n, g["old/counts"].shape[1])
    old.counts[:], old.state[:], old.struct_id[:], old.w[:] = (g["old/counts"], g["old/state"], g["old/struct_id"],
                                                                g["old/w"])
    old.total_weight = O.sequential_uniform_total(n)          # n x add(.., 1/n) (ImportanceSampler.hpp:79-92)
    rng = O.Rng(words)
    priors = g["priors/counts"]
    sid, state, counts = [], [], []
    while len(sid) < n:
        i = O.weighted_sample(old, rng)                                            # :200
        k = int(old.struct_id[i])
        tp, op = g.t_par[k], g.o_par[k]
        if O.lib().orc_boolean(rng.ref()):                                         # :206-208
            tp2, op2 = tp, op
        else:
            tp2, op2 = O.mutate_structure(m, tp, op, MUT_FACTORED_TIGER, rng)
        k2 = key[(tp2.tobytes(), op2.tobytes())]
        c = priors[k2].copy()
        _, last = O.mh_replay_history(m, tp2, op2, c, g["history/len"], g["history/a"], g["history/o"], rng)  # :215
        sz, sz2 = m.struct_size(tp, op), m.struct_size(tp2, op2)
        old_c = np.zeros(stride, np.float32)
        old_c[:old.counts.shape[1]] = old.counts[i]
        old_score = O.log_bd_score(m, tp, op, old_c[:sz].copy(), priors[k][:sz].copy())      # :237
        new_score = O.log_bd_score(m, tp2, op2, c[:sz2].copy(), priors[k2][:sz2].copy())     # :238
        if math.log(O.lib().orc_uniform01(rng.ref())) < new_score - old_score:     # :240
            sid.append(k2), state.append(last), counts.append(c)
    return np.array(sid), np.array(state), np.stack(counts), rng.cur


def test_mh_restated_over_oracle_primitives_equals_the_reference():
    O, g, m, key = load()
    sid, state, counts, used = mh(O, g, m, key, g["mh/words"])
    assert used == len(g["mh/words"])                 # the same number of mt19937 words
    np.testing.assert_array_equal(sid, g["new/struct_id"])
    np.testing.assert_array_equal(state, g["new/state"])
    np.testing.assert_array_equal(counts, g["new/counts"])
    assert len(np.unique(sid)) > 1                    # MH did move between structures


def test_replay_history_retries_episodes_and_restores_counts():
    """computePosterior on a model that cannot explain the first observation at once: attempts exceed
    the episode count, and every cell ends at prior + (its increments of the successful attempts): the
    total mass grows by exactly J per history step."""
    O, g, m, key = load()
    tp, op = g.t_par[0], g.o_par[0]
    c = g["priors/counts"][0].copy()
    before = c.astype(np.float64).sum()
    rs = np.random.RandomState(4)
    words = rs.randint(0, 2**32, size=200000, dtype=np.uint64).astype(np.uint32)
    attempts, last = O.mh_replay_history(m, tp, op, c, g["history/len"], g["history/a"], g["history/o"], O.Rng(words))
    steps = int(g["history/len"].sum())
    assert attempts >= len(g["history/len"]) and 0 <= last < m.c.S
    J = m.c.FS + m.c.FO
    assert abs(c.astype(np.float64).sum() - before - J * steps) < 1e-3 * J * steps
    # too few attempts allowed: reported, not looped forever
    c2 = g["priors/counts"][0].copy()
    assert O.mh_replay_history(m, tp, op, c2, g["history/len"], g["history/a"], g["history/o"], O.Rng(words),
                               max_attempts=1)[0] == -1
