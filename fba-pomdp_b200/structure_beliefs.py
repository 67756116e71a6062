"""Host-side mirror of the reference's COMPOSITE structure beliefs over the C ABI: the
classes whose update is a fixed sequence of the primitives of beliefs.py on two or three particle filters.
Same names, members and call order as the reference (paths relative to samkatt/fba-pomdp):

  CheatingReinvigoration      src/beliefs/bayes-adaptive/prototypes/CheatingReinvigoration.cpp:27-147
  StructureIncubatorSampling  src/beliefs/bayes-adaptive/factored/StructureIncubatorSampling.cpp:20-187

Every method ends in CUDA calls on the filters' fba_belief handles; there is no CPU path. One Rng is
passed per call and consumed by the primitives in the reference's draw order, so that in REPLAY mode a
whole updateEstimation reproduces the reference's bit for bit (tests/test_cuda_composite.py).
"""
import ctypes as C
import math

import numpy as np

from . import capi
from .beliefs import BAImportanceSampling, BARejectionSampling, _check
from .capi import FbaError, ptr


def _flat(sim, n, particles, stride):
    b = BARejectionSampling(n)
    b.initiate(sim, struct_id=particles["struct_id"], counts=particles["counts"], state=particles["state"],
               stride=stride)
    return b


def _weighted(sim, n, particles, stride):
    b = BAImportanceSampling(n)
    b.initiate(sim, struct_id=particles["struct_id"], counts=particles["counts"], state=particles["state"],
               stride=stride)
    return b


def replace_from(dst, dst_index, src, src_index):
    """src[src_index[j]] -> dst[dst_index[j]], in order (WeightedFilter::replace / FlatFilter slot writes)."""
    di = np.ascontiguousarray(dst_index, np.int64)
    si = np.ascontiguousarray(src_index, np.int64)
    _check(dst.ctx.h, dst.L.fba_belief_replace_from(dst.h, ptr(di), src.h, ptr(si), len(di)))


class CheatingReinvigoration:
    """beliefs::bayes_adaptive::prototypes::CheatingReinvigoration: a weighted belief tracked by importance
    sampling next to a flat filter of correctly structured particles tracked by rejection sampling; when the
    accumulated likelihood drops below the threshold, `cheat_amount` particles of the latter are copied
    into the former."""

    def __init__(self, size, cheat_amount, resample_threshold):
        if size < 1 or cheat_amount < 1:  # CheatingReinvigoration.cpp:34-38
            raise FbaError(capi.ERR_INVALID, "CheatingReinvigoration::cannot initiate belief of size < 1 ("
                           + str(size) + "), or resample size of < 1 (" + str(cheat_amount) + ")")
        if resample_threshold >= 0:  # :40-44
            raise FbaError(capi.ERR_INVALID, "CheatingReinvigoration::cannot initiate with resample_threshold >= 0 (is:"
                           + str(resample_threshold) + ")")
        self._size, self._cheat_amount, self._resample_threshold = size, cheat_amount, resample_threshold
        self._belief = self._correct_structured_belief = None
        self._likelihood = 1.0
        self.cheats = 0

    def initiate(self, simulator, *, belief, correct_structured, stride):
        """belief / correct_structured: dicts (struct_id, counts, state) of what sampleStartState /
        sampleCorrectGraphState produced on the host (:68-93)."""
        self._correct_structured_belief = _flat(simulator, self._size, correct_structured, stride)
        self._belief = _weighted(simulator, self._size, belief, stride)
        self._likelihood = 1.0

    def updateEstimation(self, a, o, rng):
        """:107-134"""
        self._correct_structured_belief.updateEstimation(a, o, rng)
        l = self._belief.update(a, o, rng)
        self._belief.resample(rng)
        self._likelihood *= l
        if (math.log(self._likelihood) if self._likelihood > 0 else -math.inf) < self._resample_threshold:
            self.cheat(rng)
            self._likelihood = 1.0

    def cheat(self, rng):
        """:136-147"""
        b = self._belief
        _check(b.ctx.h, b.L.fba_belief_cheat(b.h, self._correct_structured_belief.h, self._cheat_amount,
                                             C.byref(rng)))
        self.cheats += 1

    def sample(self, rng):
        return self._belief.sample(rng)

    def resetDomainStateDistribution(self, rng):
        """:50-66: resetDomainState on every particle of the cheating filter, then of the belief (its
        weights and order stay as they are)."""
        for b in (self._correct_structured_belief, self._belief):
            _check(b.ctx.h, b.L.fba_belief_redraw_domain_states(b.h, C.byref(rng)))

    def free(self, _simulator=None):
        for b in (self._belief, self._correct_structured_belief):
            if b is not None:
                b.free()
        self._belief = self._correct_structured_belief = None


class StructureIncubatorSampling:
    """beliefs::bayes_adaptive::factored::StructureIncubatorSampling: the belief and a fully connected
    belief (flat, rejection sampling) plus a weighted SHADOW belief of bred particles (importance
    sampling) whose heavy particles are promoted into the belief."""

    def __init__(self, size, reinvigor_amount, threshold, mutate_kind):
        if size < 1 or reinvigor_amount < 1:  # StructureIncubatorSampling.cpp:29-34
            raise FbaError(capi.ERR_INVALID,
                           "StructureIncubatorSampling::Cannot initiate Incubator belief update with size < 1 ("
                           + str(size) + ") or resample size < 1 (" + str(reinvigor_amount) + ")")
        if threshold <= 0 or threshold > 1:  # :36-40
            raise FbaError(capi.ERR_INVALID, "StructureIncubatorSampling::must initiate with 1 < threshold <= 0 (is:"
                           + str(threshold) + ")")
        self._size, self._shadow_reinvigor_amount, self._real_reinvigor_threshold = size, reinvigor_amount, threshold
        self._mutate = mutate_kind
        self._belief = self._fully_connected_belief = self._shadow_belief = None
        self.promoted = 0

    def initiate(self, simulator, *, belief, fully_connected, stride, rng=None, shadow=None):
        """belief / fully_connected: host particles as above (:65-73). The shadow belief is bred from them
        (:74-80) with `rng`, or — to start from the reference's own shadow particles — uploaded as is."""
        self._belief = _flat(simulator, self._size, belief, stride)
        self._fully_connected_belief = _flat(simulator, self._size, fully_connected, stride)
        if shadow is not None:
            self._shadow_belief = _weighted(simulator, self._size, shadow, stride)
        else:
            self._shadow_belief = _weighted(simulator, self._size, belief, stride)  # placeholders, all overwritten
            self._breed_into(np.arange(self._size), rng)
            w = np.full(self._size, 1.0 / self._size)  # add(s, 1 / size) x size
            b = self._shadow_belief
            _check(b.ctx.h, b.L.fba_belief_upload(b.h, 0, self._size, None, None, None, ptr(w)))

    def _breed_into(self, slots, rng):
        s = np.ascontiguousarray(slots, np.int64)
        b = self._shadow_belief
        _check(b.ctx.h, b.L.fba_belief_breed_into(b.h, ptr(s), len(s), self._belief.h, self._fully_connected_belief.h,
                                                  self._mutate, C.byref(rng)))

    def reinvigorateBelief(self, rng):
        """:155-187"""
        n = C.c_int64(0)
        b = self._shadow_belief
        _check(b.ctx.h, b.L.fba_belief_promote(b.h, self._belief.h, self._real_reinvigor_threshold, C.byref(rng),
                                               C.byref(n)))
        self.promoted += n.value
        return n.value

    def reinvigorateShadowBelief(self, rng):
        """:139-153"""
        idx = np.zeros(self._shadow_reinvigor_amount, np.int64)
        b = self._shadow_belief
        _check(b.ctx.h, b.L.fba_belief_least_likely(b.h, len(idx), ptr(idx)))
        self._breed_into(idx, rng)
        return idx

    def updateEstimation(self, a, o, rng):
        """:107-137"""
        self.reinvigorateBelief(rng)
        self.reinvigorateShadowBelief(rng)
        self._belief.updateEstimation(a, o, rng)
        self._fully_connected_belief.updateEstimation(a, o, rng)
        self._shadow_belief.update(a, o, rng)
        self._shadow_belief.resample(rng)

    def sample(self, rng):
        return self._belief.sample(rng)

    def resetDomainStateDistribution(self, rng):
        """:46-61: resetDomainState on every particle of the three filters, in this order"""
        for b in (self._belief, self._fully_connected_belief, self._shadow_belief):
            _check(b.ctx.h, b.L.fba_belief_redraw_domain_states(b.h, C.byref(rng)))

    def free(self, _simulator=None):
        for b in (self._belief, self._fully_connected_belief, self._shadow_belief):
            if b is not None:
                b.free()
        self._belief = self._fully_connected_belief = self._shadow_belief = None


class NestedBelief:
    """beliefs::bayes_adaptive::NestedBelief (src/beliefs/bayes-adaptive/NestedBelief.cpp): a weighted top
    filter of count blocks, each with its own flat bottom filter of domain states (fba_nested_*)."""

    def __init__(self, top_filter_size, bottom_filter_size):
        if top_filter_size < 1 or bottom_filter_size < 1:  # NestedBelief.cpp:19-26
            raise FbaError(capi.ERR_INVALID, "NestedBelief: cannot initiate with filter size < 1 (top: "
                           + str(top_filter_size) + ", bottom: " + str(bottom_filter_size) + ")")
        self._top_filter_size, self._bottom_filter_size = top_filter_size, bottom_filter_size
        self.h = None
        self.attempts = None

    def initiate(self, simulator, *, struct_id, counts, states, stride=0):
        """the top particles the host-side prior produced (structure ids, count blocks) and the bottom filters'
        domain start states [top, bottom] (NestedBelief.cpp:63-90)"""
        self.sim, self.ctx, self.L = simulator, simulator.ctx, simulator.L
        h = C.c_void_p()
        _check(self.ctx.h, self.L.fba_nested_create(self.ctx.h, simulator.h, self._top_filter_size,
                                                    self._bottom_filter_size, stride, C.byref(h)))
        self.h = h
        self._top = BAImportanceSampling(self._top_filter_size)
        self._top.sim, self._top.ctx, self._top.L = simulator, self.ctx, self.L
        self._top.h = C.c_void_p(self.L.fba_nested_top(self.h))     # owned by the nested object
        self._top._init_explicit(self._top.h, struct_id, counts, np.zeros(self._top_filter_size, np.int32))
        self.upload_states(states)

    def upload_states(self, states):
        s = np.ascontiguousarray(states, np.int32).reshape(self._top_filter_size, self._bottom_filter_size)
        _check(self.ctx.h, self.L.fba_nested_upload_states(self.h, 0, self._top_filter_size, ptr(s)))

    def updateEstimation(self, a, o, rng, max_attempts=1 << 40):
        """:129-193"""
        att = np.zeros(self._top_filter_size, np.int64)
        _check(self.ctx.h, self.L.fba_nested_update(self.h, a, o, C.byref(rng), int(max_attempts), ptr(att)))
        self.attempts = att

    def resetDomainStateDistribution(self, rng):
        """:33-61"""
        _check(self.ctx.h, self.L.fba_nested_reset_domain_states(self.h, C.byref(rng)))

    def sample(self, rng):
        """:117-127 -> (index of the drawn top particle, its domain state)"""
        i, s = C.c_int64(0), C.c_int32(0)
        _check(self.ctx.h, self.L.fba_nested_sample(self.h, C.byref(rng), C.byref(i), C.byref(s)))
        return i.value, s.value

    def download(self):
        d = self._top.download()
        s = np.zeros((self._top_filter_size, self._bottom_filter_size), np.int32)
        _check(self.ctx.h, self.L.fba_nested_download_states(self.h, 0, self._top_filter_size, ptr(s)))
        d["states"] = s
        del d["state"]
        return d

    def free(self, _simulator=None):
        if self.h:
            self.L.fba_nested_destroy(self.h)
            self.h = None
