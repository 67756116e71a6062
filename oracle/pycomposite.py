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
