"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the plain-C restatement,
oracle/fba_oracle.c). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this. Never on the product path."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAXF = 16

DOM_TABLE, DOM_TIGER, DOM_FACTORED_TIGER, DOM_SYSADMIN, DOM_GRIDWORLD, DOM_CA = range(6)
ACT_UNIFORM_INT, ACT_SLOW_INT = 0, 1
START_CONST, START_BOOL, START_UNIFORM_INT, START_SLOW2, START_CATEGORICAL = range(5)
MUT_FACTORED_TIGER, MUT_CA, MUT_SYSADMIN, MUT_GRIDWORLD = range(4)


class CModel(C.Structure):
    _fields_ = [
        ("S", C.c_int32), ("A", C.c_int32), ("O", C.c_int32), ("FS", C.c_int32), ("FO", C.c_int32),
        ("feat_s", C.c_int32 * MAXF), ("feat_o", C.c_int32 * MAXF),
        ("tabular", C.c_int32), ("domain", C.c_int32),
        ("dom_ip", C.c_int32 * 32), ("dom_dp", C.c_double * 8),
        ("rew_sa", C.c_void_p), ("rew_as2", C.c_void_p), ("term_sa", C.c_void_p),
        ("term_as2", C.c_void_p),
        ("action_draw", C.c_int32), ("start_kind", C.c_int32), ("start_ip", C.c_int32 * 4),
        ("start_values", C.c_void_p), ("start_total", C.c_double), ("start_table", C.c_void_p),
    ]


class CRng(C.Structure):
    _fields_ = [("words", C.c_void_p), ("n", C.c_int64), ("cur", C.c_int64), ("overrun", C.c_int32)]


class CStructs(C.Structure):
    _fields_ = [("n_structs", C.c_int32), ("cap", C.c_int32), ("t_par", C.c_void_p),
                ("o_par", C.c_void_p)]


class CBelief(C.Structure):
    _fields_ = [("N", C.c_int64), ("stride", C.c_int64), ("counts", C.c_void_p),
                ("state", C.c_void_p), ("struct_id", C.c_void_p), ("w", C.c_void_p),
                ("total_weight", C.c_double)]


def build(force=False):
    """gcc oracle/fba_oracle.c -> oracle/liboracle.so (building the checker is not using it)."""
    src = os.path.join(_HERE, "fba_oracle.c")
    hdr = os.path.join(_HERE, "fba_oracle.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                           "-Wall", "-Wextra", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.orc_uniform01.restype = dbl
        L.orc_uniform01.argtypes = [vp]
        L.orc_boolean.restype = C.c_int
        L.orc_boolean.argtypes = [vp]
        L.orc_uniform_int.restype = i32
        L.orc_uniform_int.argtypes = [vp, C.c_uint32]
        L.orc_struct_size.restype = i64
        L.orc_struct_size.argtypes = [vp, vp, vp]
        L.orc_struct_offsets.restype = i64
        L.orc_struct_offsets.argtypes = [vp, vp, vp, vp]
        L.orc_reward.restype = dbl
        L.orc_reward.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.orc_sample_start_state.restype = C.c_int
        L.orc_sample_start_state.argtypes = [vp, vp]
        L.orc_step.restype = dbl
        L.orc_step.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
        L.orc_obs_prob.restype = dbl
        L.orc_obs_prob.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int]
        L.orc_is_update.restype = dbl
        L.orc_is_update.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        L.orc_weighted_sample.restype = i64
        L.orc_weighted_sample.argtypes = [vp, vp]
        L.orc_is_propose.restype = dbl
        L.orc_is_propose.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, dbl]
        L.orc_normalize.restype = dbl
        L.orc_normalize.argtypes = [vp, i64, dbl]
        L.orc_weighted_sample_many.argtypes = [vp, i64, dbl, vp, i64, vp]
        L.orc_block_checksums.argtypes = [vp, i64, i64, vp]
        L.orc_mutate_structure.argtypes = [vp, vp, vp, C.c_int, vp]
        L.orc_increment_counts.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_mh_replay_history.restype = i64
        L.orc_mh_replay_history.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, i64, vp]
        L.orc_is_resample.argtypes = [vp, vp, vp, vp]
        L.orc_is_reset_domain_states.argtypes = [vp, vp, vp, vp, vp]
        L.orc_reject_sample.restype = i64
        L.orc_reject_sample.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
        L.orc_flat_reset_domain_states.argtypes = [vp, vp, vp]
        L.orc_marginalize_node.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_int, vp]
        L.orc_reinvigorate.restype = C.c_int
        L.orc_reinvigorate.argtypes = [vp, vp, vp, vp, i64, C.c_int, vp]
        L.orc_rollout.restype = dbl
        L.orc_rollout.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, dbl, vp]
        L.orc_breed_into.restype = C.c_int
        L.orc_breed_into.argtypes = [vp, vp, vp, vp, vp, vp, i64, C.c_int, vp]
        L.orc_replace_weight.argtypes = [vp, i64]
        L.orc_cheat.argtypes = [vp, vp, i64, vp]
        L.orc_least_likely.argtypes = [vp, i64, i64, vp]
        L.orc_promote.restype = i64
        L.orc_promote.argtypes = [vp, vp, dbl, vp]
        L.orc_state_history_rs.restype = i64
        L.orc_state_history_rs.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, i64, vp]
        L.orc_state_history_msg.restype = C.c_int
        L.orc_state_history_msg.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]
        L.orc_flatten_model.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_add_history_counts.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        L.orc_nested_update_particle.restype = i64
        L.orc_nested_update_particle.argtypes = [vp, vp, vp, vp, vp, vp, i64, C.c_int, C.c_int, vp, i64]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Rng:
    """A replay stream of mt19937 words."""

    def __init__(self, words):
        self.words = np.ascontiguousarray(words, np.uint32)
        self.c = CRng(_p(self.words), len(self.words), 0, 0)

    @property
    def cur(self):
        return self.c.cur

    @property
    def overrun(self):
        return bool(self.c.overrun)

    def ref(self):
        return C.byref(self.c)


class Model:
    """Domain + feature description. `desc` is the plain dict stored in fixtures
    (see oracle/gen_golden.py / tests/golden/*.npz)."""

    def __init__(self, desc):
        d = self.desc = dict(desc)
        c = self.c = CModel()
        c.S, c.A, c.O = int(d["S"]), int(d["A"]), int(d["O"])
        fs = np.asarray(d["feat_s"], np.int32)
        fo = np.asarray(d["feat_o"], np.int32)
        c.FS, c.FO = len(fs), len(fo)
        for i, v in enumerate(fs):
            c.feat_s[i] = int(v)
        for i, v in enumerate(fo):
            c.feat_o[i] = int(v)
        c.tabular = int(d["tabular"])
        c.domain = int(d["domain"])
        for i, v in enumerate(np.asarray(d.get("dom_ip", []), np.int32)):
            c.dom_ip[i] = int(v)
        for i, v in enumerate(np.asarray(d.get("dom_dp", []), np.float64)):
            c.dom_dp[i] = float(v)
        self._keep = {}
        for k, dt in (("rew_sa", np.float64), ("rew_as2", np.float64), ("term_sa", np.uint8),
                      ("term_as2", np.uint8), ("start_values", np.float32),
                      ("start_table", np.int32)):
            v = d.get(k)
            if v is not None and len(np.atleast_1d(v)):
                self._keep[k] = np.ascontiguousarray(v, dt)
                setattr(c, k, _p(self._keep[k]))
        c.action_draw = int(d.get("action_draw", 0))
        c.start_kind = int(d.get("start_kind", 0))
        for i, v in enumerate(np.asarray(d.get("start_ip", []), np.int32)):
            c.start_ip[i] = int(v)
        c.start_total = float(d.get("start_total", 0.0))
        self.S, self.A, self.O, self.FS, self.FO = c.S, c.A, c.O, c.FS, c.FO
        self.feat_s, self.feat_o = fs, fo

    def ref(self):
        return C.byref(self.c)

    def struct_size(self, t_par, o_par):
        t = np.ascontiguousarray(t_par, np.uint32)
        o = np.ascontiguousarray(o_par, np.uint32)
        return lib().orc_struct_size(self.ref(), _p(t), _p(o))

    def struct_offsets(self, t_par, o_par):
        t = np.ascontiguousarray(t_par, np.uint32)
        o = np.ascontiguousarray(o_par, np.uint32)
        off = np.zeros(self.A * (self.FS + self.FO), np.int64)
        n = lib().orc_struct_offsets(self.ref(), _p(t), _p(o), _p(off))
        return off, n

    def reward(self, s, a, s2):
        t = C.c_int(0)
        r = lib().orc_reward(self.ref(), s, a, s2, C.byref(t))
        return r, bool(t.value)

    def sample_start_state(self, rng):
        return lib().orc_sample_start_state(self.ref(), rng.ref())


class Structs:
    """Structure table: parent bitmasks per (structure, action, node)."""

    def __init__(self, model, t_par, o_par, cap=None):
        t_par = np.asarray(t_par, np.uint32).reshape(-1, model.A * model.FS)
        o_par = np.asarray(o_par, np.uint32).reshape(-1, model.A * model.FO)
        n = len(t_par)
        cap = max(cap or n, n)
        self.model = model
        self.t_par = np.zeros((cap, model.A * model.FS), np.uint32)
        self.o_par = np.zeros((cap, model.A * model.FO), np.uint32)
        self.t_par[:n] = t_par
        self.o_par[:n] = o_par
        self.c = CStructs(n, cap, _p(self.t_par), _p(self.o_par))

    @property
    def n(self):
        return self.c.n_structs

    def ref(self):
        return C.byref(self.c)

    def sizes(self):
        return np.array([self.model.struct_size(self.t_par[i], self.o_par[i])
                         for i in range(self.n)], np.int64)


class Belief:
    """N particles as flat arrays (the same arrays the CUDA path downloads)."""

    def __init__(self, N, stride, weighted=True):
        self.N, self.stride = int(N), int(stride)
        self.counts = np.zeros((self.N, self.stride), np.float32)
        self.state = np.zeros(self.N, np.int32)
        self.struct_id = np.zeros(self.N, np.int32)
        self.w = np.full(self.N, 1.0 / self.N, np.float64) if weighted else None
        self.c = CBelief(self.N, self.stride, _p(self.counts), _p(self.state), _p(self.struct_id),
                         _p(self.w), 1.0)

    @property
    def total_weight(self):
        return self.c.total_weight

    @total_weight.setter
    def total_weight(self, v):
        self.c.total_weight = float(v)

    def ref(self):
        return C.byref(self.c)

    def clone_empty(self):
        return Belief(self.N, self.stride, self.w is not None)

    def copy(self):
        b = self.clone_empty()
        b.counts[:] = self.counts
        b.state[:] = self.state
        b.struct_id[:] = self.struct_id
        if self.w is not None:
            b.w[:] = self.w
        b.total_weight = self.total_weight
        return b


def sequential_uniform_total(n):
    """WeightedFilter::_total_weight after n add(s, 1/n) calls (WeightedFilter.cpp:60-66)."""
    w = 1.0 / float(n)
    return float(np.cumsum(np.full(n, w, np.float64))[-1]) if n else 0.0


def step(model, t_par, o_par, counts, state, a, update_counts, rng):
    st = np.array([state], np.int32)
    o = C.c_int(0)
    t = C.c_int(0)
    tp = np.ascontiguousarray(t_par, np.uint32)
    op = np.ascontiguousarray(o_par, np.uint32)
    r = lib().orc_step(model.ref(), _p(tp), _p(op), _p(counts), _p(st), a, int(update_counts),
                       rng.ref(), C.byref(o), C.byref(t))
    return int(st[0]), o.value, bool(t.value), r


def log_bd_score(model, t_par, o_par, counts, prior_counts):
    L = lib()
    L.orc_log_bd_score.restype = C.c_double
    L.orc_log_bd_score.argtypes = [C.c_void_p] * 5
    tp = np.ascontiguousarray(t_par, np.uint32)
    op = np.ascontiguousarray(o_par, np.uint32)
    c = np.ascontiguousarray(counts, np.float32)
    pc = np.ascontiguousarray(prior_counts, np.float32)
    return L.orc_log_bd_score(model.ref(), _p(tp), _p(op), _p(c), _p(pc))


def obs_prob(model, t_par, o_par, counts, state, a, o):
    tp = np.ascontiguousarray(t_par, np.uint32)
    op = np.ascontiguousarray(o_par, np.uint32)
    return lib().orc_obs_prob(model.ref(), _p(tp), _p(op), _p(counts), state, a, o)


def is_update(model, structs, belief, a, o, rng):
    return lib().orc_is_update(model.ref(), structs.ref(), belief.ref(), a, o, rng.ref())


def is_propose(model, structs, belief, a, o, rng, running_total=0.0):
    """the per-particle loop of importance_sampling::update over one block of particles; returns the
    running un-normalised total including this block"""
    return lib().orc_is_propose(model.ref(), structs.ref(), belief.ref(), a, o, rng.ref(), float(running_total))


def normalize(w, total):
    """WeightedFilter::normalize in place; returns the new _total_weight"""
    assert w.dtype == np.float64 and w.flags.c_contiguous
    return lib().orc_normalize(_p(w), len(w), float(total))


def weighted_sample_many(w, total_weight, rng, n_draws):
    """n_draws x WeightedFilter::sample, identical results, O(n + n_draws log n)"""
    w = np.ascontiguousarray(w, np.float64)
    out = np.zeros(int(n_draws), np.int64)
    lib().orc_weighted_sample_many(_p(w), len(w), float(total_weight), rng.ref(), int(n_draws), _p(out))
    return out


def block_checksums(counts):
    c = np.ascontiguousarray(counts, np.float32)
    out = np.zeros(c.shape[0], np.uint64)
    lib().orc_block_checksums(_p(c), c.shape[0], c.shape[1], _p(out))
    return out


def mutate_structure(model, t_par, o_par, mutate_kind, rng):
    """FBAPOMDP::mutate on one structure's parent bitmasks -> new (t_par, o_par)"""
    tp = np.ascontiguousarray(t_par, np.uint32).copy()
    op = np.ascontiguousarray(o_par, np.uint32).copy()
    lib().orc_mutate_structure(model.ref(), _p(tp), _p(op), int(mutate_kind), rng.ref())
    return tp, op


def mh_replay_history(model, t_par, o_par, counts, episode_len, actions, observations, rng, max_attempts=1 << 40):
    """computePosterior (MHNIPS2018.cpp:41-109) in place on `counts` -> (episode attempts, last state)"""
    tp = np.ascontiguousarray(t_par, np.uint32)
    op = np.ascontiguousarray(o_par, np.uint32)
    ln = np.ascontiguousarray(episode_len, np.int32)
    ac = np.ascontiguousarray(actions, np.int32)
    ob = np.ascontiguousarray(observations, np.int32)
    assert counts.dtype == np.float32 and counts.flags.c_contiguous
    last = C.c_int32(0)
    n = lib().orc_mh_replay_history(model.ref(), _p(tp), _p(op), _p(counts), len(ln), _p(ln), _p(ac), _p(ob),
                                    rng.ref(), int(max_attempts), C.byref(last))
    return int(n), int(last.value)


def is_resample(src, rng):
    dst = src.clone_empty()
    anc = np.zeros(src.N, np.int64)
    lib().orc_is_resample(src.ref(), dst.ref(), rng.ref(), _p(anc))
    return dst, anc


def is_reset_domain_states(model, src, rng):
    dst = src.clone_empty()
    anc = np.zeros(src.N, np.int64)
    lib().orc_is_reset_domain_states(model.ref(), src.ref(), dst.ref(), rng.ref(), _p(anc))
    return dst, anc


def weighted_sample(belief, rng):
    return lib().orc_weighted_sample(belief.ref(), rng.ref())


def reject_sample(model, structs, src, a, o, rng):
    dst = src.clone_empty()
    anc = np.zeros(src.N, np.int64)
    attempts = lib().orc_reject_sample(model.ref(), structs.ref(), src.ref(), dst.ref(), a, o,
                                       rng.ref(), _p(anc))
    return dst, anc, attempts


def flat_reset_domain_states(model, belief, rng):
    lib().orc_flat_reset_domain_states(model.ref(), belief.ref(), rng.ref())


def reinvigorate(model, structs, belief, fc, amount, mutate_kind, rng):
    rc = lib().orc_reinvigorate(model.ref(), structs.ref(), belief.ref(), fc.ref(), amount,
                                mutate_kind, rng.ref())
    if rc:
        raise RuntimeError("orc_reinvigorate failed: %d" % rc)


def breed_into(model, structs, dst, dst_slots, belief, fc, mutate_kind, rng):
    """len(dst_slots) x breed into the given slots of dst (structure donors from belief, counts from fc)"""
    s = np.ascontiguousarray(dst_slots, np.int64)
    rc = lib().orc_breed_into(model.ref(), structs.ref(), dst.ref(), _p(s), belief.ref(), fc.ref(), len(s),
                              mutate_kind, rng.ref())
    if rc:
        raise RuntimeError("orc_breed_into failed: %d" % rc)


def cheat(belief, correct, amount, rng):
    lib().orc_cheat(belief.ref(), correct.ref(), amount, rng.ref())


def least_likely(w, n):
    w = np.ascontiguousarray(w, np.float64)
    out = np.zeros(n, np.int64)
    lib().orc_least_likely(_p(w), len(w), n, _p(out))
    return out


def promote(shadow, belief, threshold, rng):
    return lib().orc_promote(shadow.ref(), belief.ref(), threshold, rng.ref())


def _hist(episode_len, actions, observations):
    return (np.ascontiguousarray(episode_len, np.int32), np.ascontiguousarray(actions, np.int32),
            np.ascontiguousarray(observations, np.int32))


def state_history(model, t_par, o_par, counts, episode_len, actions, observations, rng, method, state_prior=None,
                  max_attempts=1 << 40):
    """MHwithinGibbs' sampleStateHistory (MHwithinGibbs.cpp:38-213): method "msg" (backward messages, forward
    sampling) or "rs" (rejection sampling) -> states, episode_len[e] + 1 per episode"""
    tp, op = np.ascontiguousarray(t_par, np.uint32), np.ascontiguousarray(o_par, np.uint32)
    ln, ac, ob = _hist(episode_len, actions, observations)
    c = np.ascontiguousarray(counts, np.float32)
    out = np.zeros(int(ln.sum()) + len(ln), np.int32)
    if method == "rs":
        n = lib().orc_state_history_rs(model.ref(), _p(tp), _p(op), _p(c), len(ln), _p(ln), _p(ac), _p(ob), rng.ref(),
                                       int(max_attempts), _p(out))
        if n < 0:
            raise RuntimeError("orc_state_history_rs: out of attempts / words")
    else:
        sp = np.ascontiguousarray(state_prior, np.float32)
        if lib().orc_state_history_msg(model.ref(), _p(tp), _p(op), _p(c), _p(sp), len(ln), _p(ln), _p(ac), _p(ob),
                                       rng.ref(), _p(out)):
            raise RuntimeError("orc_state_history_msg: out of words")
    return out


def flatten_model(model, t_par, o_par, counts):
    """BABNModel::flattenT / flattenO -> (T[S, A, S], O[A, S, O]) float32"""
    tp, op = np.ascontiguousarray(t_par, np.uint32), np.ascontiguousarray(o_par, np.uint32)
    c = np.ascontiguousarray(counts, np.float32)
    T = np.zeros((model.S, model.A, model.S), np.float32)
    Ob = np.zeros((model.A, model.S, model.O), np.float32)
    lib().orc_flatten_model(model.ref(), _p(tp), _p(op), _p(c), _p(T), _p(Ob))
    return T, Ob


def add_history_counts(model, t_par, o_par, counts, episode_len, actions, observations, states):
    """MHwithinGibbs::computePosteriorCounts, in place on `counts`"""
    tp, op = np.ascontiguousarray(t_par, np.uint32), np.ascontiguousarray(o_par, np.uint32)
    ln, ac, ob = _hist(episode_len, actions, observations)
    st = np.ascontiguousarray(states, np.int32)
    assert counts.dtype == np.float32 and counts.flags.c_contiguous
    lib().orc_add_history_counts(model.ref(), _p(tp), _p(op), _p(counts), len(ln), _p(ln), _p(ac), _p(ob), _p(st))


def nested_update_particle(model, t_par, o_par, counts, states_in, a, o, rng, max_attempts=1 << 40):
    """NestedBelief::updateEstimation for one top particle, counts updated in place -> (new bottom states, attempts)"""
    tp = np.ascontiguousarray(t_par, np.uint32)
    op = np.ascontiguousarray(o_par, np.uint32)
    si = np.ascontiguousarray(states_in, np.int32)
    so = np.zeros_like(si)
    assert counts.dtype == np.float32 and counts.flags.c_contiguous
    n = lib().orc_nested_update_particle(model.ref(), _p(tp), _p(op), _p(counts), _p(si), _p(so), len(si), a, o,
                                         rng.ref(), int(max_attempts))
    return so, int(n)


def rollout(model, t_par, o_par, counts, start_state, depth, discount, rng):
    tp = np.ascontiguousarray(t_par, np.uint32)
    op = np.ascontiguousarray(o_par, np.uint32)
    return lib().orc_rollout(model.ref(), _p(tp), _p(op), _p(counts), start_state, depth, discount,
                             rng.ref())
