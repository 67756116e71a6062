"""TEST INFRASTRUCTURE — the CPU oracle's restatement of the reference's COMPOSITE structure beliefs as
fixed sequences of the oracle primitives (oracle/fba_oracle.c through pyoracle), every step citing the
reference lines it follows. Pinned bit for bit, word for word, against the reference's own classes by
tests/test_oracle_composite.py through tests/golden/composite.npz (oracle/gen_composite.py).
Only tests/ may import this; never on the product path."""
import math

import numpy as np

import pyoracle as O


class Cheating:
    """beliefs::bayes_adaptive::prototypes::CheatingReinvigoration (prototypes/CheatingReinvigoration.cpp)"""

    def __init__(self, model, structs, belief, correct, cheat_amount, threshold):
        self.m, self.st = model, structs
        self.belief, self.correct = belief, correct          # weighted, flat
        self.amount, self.threshold = cheat_amount, threshold
        self.likelihood = 1.0

    def update(self, a, o, rng):
        """:107-134"""
        self.correct, _, _ = O.reject_sample(self.m, self.st, self.correct, a, o, rng)
        l = O.is_update(self.m, self.st, self.belief, a, o, rng)
        self.belief, _ = O.is_resample(self.belief, rng)
        self.likelihood *= l
        if (math.log(self.likelihood) if self.likelihood > 0 else -math.inf) < self.threshold:
            O.cheat(self.belief, self.correct, self.amount, rng)          # :136-147
            self.likelihood = 1.0

    def reset(self, rng):
        """:50-66: resetDomainState on the cheating filter's particles, then on the belief's"""
        O.flat_reset_domain_states(self.m, self.correct, rng)
        O.flat_reset_domain_states(self.m, self.belief, rng)


class Incubator:
    """beliefs::bayes_adaptive::factored::StructureIncubatorSampling (factored/StructureIncubatorSampling.cpp)"""

    def __init__(self, model, structs, belief, fc, shadow, amount, threshold, mutate_kind):
        self.m, self.st = model, structs
        self.belief, self.fc, self.shadow = belief, fc, shadow   # flat, flat, weighted
        self.amount, self.threshold, self.mutate = amount, threshold, mutate_kind

    def reinvigorate_belief(self, rng):                          # :155-187
        return O.promote(self.shadow, self.belief, self.threshold, rng)

    def reinvigorate_shadow(self, rng):                          # :139-153
        idx = O.least_likely(self.shadow.w, self.amount)
        O.breed_into(self.m, self.st, self.shadow, idx, self.belief, self.fc, self.mutate, rng)
        return idx

    def reject(self, which, a, o, rng):
        b, _, _ = O.reject_sample(self.m, self.st, getattr(self, which), a, o, rng)
        setattr(self, which, b)

    def update(self, a, o, rng):
        """:107-137"""
        self.reinvigorate_belief(rng)
        self.reinvigorate_shadow(rng)
        self.reject("belief", a, o, rng)
        self.reject("fc", a, o, rng)
        O.is_update(self.m, self.st, self.shadow, a, o, rng)
        self.shadow, _ = O.is_resample(self.shadow, rng)

    def reset(self, rng):
        """:46-61"""
        for b in (self.belief, self.fc, self.shadow):
            O.flat_reset_domain_states(self.m, b, rng)


class Nested:
    """beliefs::bayes_adaptive::NestedBelief (src/beliefs/bayes-adaptive/NestedBelief.cpp): `top` is a weighted
    oracle belief (counts, structures, weights), `states` the bottom filters [n_top, n_bottom]"""

    def __init__(self, model, structs, top, states):
        self.m, self.st, self.top = model, structs, top
        self.states = np.ascontiguousarray(states, np.int32).copy()
        self.attempts = np.zeros(top.N, np.int64)

    def update(self, a, o, rng, slice_words=None):
        """:129-193. rng: ONE stream consumed particle after particle (the reference's order) — or, with
        slice_words = w, top particle i draws from words [i w, (i + 1) w) (how the GPU kernel splits it)."""
        for i in range(self.top.N):
            g = rng if slice_words is None else O.Rng(rng.words[i * slice_words:(i + 1) * slice_words])
            sid = int(self.top.struct_id[i])
            new, n = O.nested_update_particle(self.m, self.st.t_par[sid], self.st.o_par[sid], self.top.counts[i],
                                              self.states[i], a, o, g)
            assert n > 0
            self.states[i] = new
            self.attempts[i] = n
            self.top.w[i] *= 1.0 / float(n)                                   # :183
        total = 0.0
        for x in self.top.w:                                                  # WeightedFilter::normalize()
            total += x
        self.top.total_weight = O.normalize(self.top.w, total)

    def reset(self, rng):
        """:33-61: n_bottom fresh domain start states per top particle"""
        for i in range(self.top.N):
            for j in range(self.states.shape[1]):
                self.states[i, j] = self.m.sample_start_state(rng)

    def sample(self, rng):
        """:117-127 -> (top index, domain state)"""
        i = O.weighted_sample(self.top, rng)
        j = O.lib().orc_uniform_int(rng.ref(), self.states.shape[1])
        return int(i), int(self.states[i, j])


def belief_from(g, prefix, stride, weighted):
    """an oracle belief from a fixture dump with full counts (<prefix>_counts / _state / _struct_id [/ _w])"""
    counts = g[prefix + "_counts"]
    b = O.Belief(counts.shape[0], stride, weighted)
    b.counts[:, :counts.shape[1]] = counts
    b.state[:] = g[prefix + "_state"]
    b.struct_id[:] = g[prefix + "_struct_id"]
    if weighted:
        b.w[:] = g[prefix + "_w"]
        b.total_weight = float(g[prefix + "_total_weight"])
    return b


def assert_matches(b, g, prefix, weighted=False):
    np.testing.assert_array_equal(b.state, g[prefix + "_state"], err_msg=prefix)
    np.testing.assert_array_equal(b.struct_id, g[prefix + "_struct_id"], err_msg=prefix)
    np.testing.assert_array_equal(b.counts.astype(np.float64).sum(1), g[prefix + "_count_sums"], err_msg=prefix)
    if (prefix + "_counts") in g.files:
        np.testing.assert_array_equal(b.counts, g[prefix + "_counts"], err_msg=prefix)
    if weighted:
        np.testing.assert_array_equal(b.w, g[prefix + "_w"], err_msg=prefix)
        assert b.total_weight == float(g[prefix + "_total_weight"]), prefix


def gibbs_reinvigorate(m, old, priors, t_par, o_par, hist, rng, method, state_prior, mutate_kind, n_out):
    """MHwithinGibbs::reinvigorate (factored/MHwithinGibbs.cpp:334-395) over the oracle primitives.
    old: the weighted oracle belief; priors[k]: FBAPOMDPPrior::computePriorModel of structure k (t_par[k],
    o_par[k]); hist = (episode lengths, actions, observations). -> (struct ids, states, counts) of the new belief."""
    key = {(t_par[k].tobytes(), o_par[k].tobytes()): k for k in range(len(t_par))}
    ln, ac, ob = hist

    def history(k, counts):                                                   # sampleStateHistory, :215-232
        return O.state_history(m, t_par[k], o_par[k], counts, ln, ac, ob, rng, method, state_prior)

    def posterior(k, seq):                                                    # computePosteriorCounts, :397-436
        c = priors[k].copy()
        O.add_history_counts(m, t_par[k], o_par[k], c, ln, ac, ob, seq)
        return c

    def score(k, c):                                                          # BABNModel::LogBDScore
        sz = m.struct_size(t_par[k], o_par[k])
        return O.log_bd_score(m, t_par[k], o_par[k], c[:sz].copy(), priors[k][:sz].copy())

    i = O.weighted_sample(old, rng)                                           # old_belief.sample(), :344
    k = int(old.struct_id[i])
    c0 = np.zeros(priors.shape[1], np.float32)
    c0[:old.counts.shape[1]] = old.counts[i]
    seq = history(k, c0)                                                      # :345-346
    model = posterior(k, seq)                                                 # :348-350
    sc = score(k, model)                                                      # :352
    sid, state, counts = [], [], []
    while len(sid) < n_out:                                                   # :356
        tp2, op2 = O.mutate_structure(m, t_par[k], o_par[k], mutate_kind, rng)    # :360
        k2 = key[(tp2.tobytes(), op2.tobytes())]
        new_model = posterior(k2, seq)                                        # :362
        new_sc = score(k2, new_model)                                         # :365
        if math.log(O.lib().orc_uniform01(rng.ref())) < new_sc - sc:          # :367
            sid.append(k2), state.append(int(seq[-1])), counts.append(new_model)  # :370-374
            seq = history(k, model)                                           # :377-378: from the model BEFORE the move
            k, model = k2, posterior(k2, seq)                                 # :381
            sc = score(k2, model)                                             # :383
    return np.array(sid), np.array(state), np.stack(counts)
