"""Host-side mirror of the reference's Belief / BABelief / rollout interfaces over the C ABI.

Same names, argument meaning and error behaviour as the reference classes they stand in for
(paths relative to samkatt/fba-pomdp):

  BAPOMDP                          src/bayes-adaptive/models/table/BAPOMDP.hpp:44-149 (and FBAPOMDP)
  BAImportanceSampling             src/beliefs/bayes-adaptive/BAImportanceSampling.cpp:17-111
  BARejectionSampling              src/beliefs/bayes-adaptive/BARejectionSampling.cpp:10-60
  ReinvigoratingRejectionSampling  src/beliefs/bayes-adaptive/factored/ReinvigoratingRejectionSampling.cpp:37-131
  rollouts                         src/planners/bayes-adaptive/RBAPOUCT.cpp:295-323

The C++ adapters a maintainer would compile into the reference are the same thin
layer in C++; this Python layer exists so the parity tests read like the reference's own tests.
Every method ends in a CUDA call: there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import FbaError, Rng, ptr


def _check(ctx, rc):
    if rc != capi.OK:
        msg = capi.lib().fba_last_error(ctx).decode() if ctx else "fba error"
        raise FbaError(rc, msg or "fba status %d" % rc)


class Context:
    """One GPU + one stream (fba_ctx)."""

    def __init__(self, device=0):
        self.L = capi.lib()
        h = C.c_void_p()
        rc = self.L.fba_ctx_create(device, C.byref(h))
        if rc == capi.ERR_NO_DEVICE:
            raise FbaError(rc, "no CUDA device: fba_pomdp_b200 has no CPU fallback")
        if rc != capi.OK:
            raise FbaError(rc, "fba_ctx_create failed with status %d" % rc)
        self.h = h
        self.device = device

    @property
    def stream(self):
        return self.L.fba_ctx_stream(self.h)

    @property
    def launches(self):
        return self.L.fba_ctx_launch_count(self.h)

    def counter(self, which=0):
        """device-side counters (0: simulated steps executed by the rollout kernels)."""
        return int(self.L.fba_ctx_counter(self.h, which))

    def synchronize(self):
        _check(self.h, self.L.fba_ctx_synchronize(self.h))

    def set_option(self, name, value):
        _check(self.h, self.L.fba_ctx_set_option(self.h, name.encode(), int(value)))

    def profile_begin(self):
        _check(self.h, self.L.fba_ctx_profile_begin(self.h))

    def profile_end(self):
        _check(self.h, self.L.fba_ctx_profile_end(self.h))

    def kernel_time(self, prefix):
        """(total ms, launches) of the kernels whose name starts with prefix, from CUDA events."""
        ms, n = C.c_double(0), C.c_int64(0)
        _check(self.h, self.L.fba_ctx_profile_get(self.h, prefix.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_times(self):
        """{kernel name: (total ms, launches)} since profile_begin."""
        n = self.L.fba_ctx_profile_list(self.h, None, 0)
        buf = C.create_string_buffer(int(n))
        self.L.fba_ctx_profile_list(self.h, buf, n)
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                name, ms, cnt = item.rsplit(":", 2)
                out[name] = (float(ms), int(cnt))
        return out

    def close(self):
        if self.h:
            self.L.fba_ctx_destroy(self.h)
            self.h = None


class BAPOMDP:
    """The hyper-state simulator description: sizes, features, domain functor, start distribution,
    structure table, and the prior's prototype particles (host-built)."""

    def __init__(self, ctx, desc, t_par, o_par, max_structures=None):
        self.ctx, self.L, self.desc = ctx, ctx.L, dict(desc)
        d = self.desc
        md = capi.ModelDesc()
        md.S, md.A, md.O = int(d["S"]), int(d["A"]), int(d["O"])
        fs = np.asarray(d["feat_s"], np.int32).reshape(-1)
        fo = np.asarray(d["feat_o"], np.int32).reshape(-1)
        if len(fs) > capi.MAX_FEATURES or len(fo) > capi.MAX_FEATURES:
            raise FbaError(capi.ERR_INVALID, "at most %d features" % capi.MAX_FEATURES)
        md.n_state_features, md.n_obs_features = len(fs), len(fo)
        for i, v in enumerate(fs):
            md.state_feature_sizes[i] = int(v)
        for i, v in enumerate(fo):
            md.obs_feature_sizes[i] = int(v)
        md.tabular, md.domain = int(d["tabular"]), int(d["domain"])
        for i, v in enumerate(np.asarray(d.get("dom_ip", []), np.int32).reshape(-1)[:32]):
            md.dom_ip[i] = int(v)
        for i, v in enumerate(np.asarray(d.get("dom_dp", []), np.float64).reshape(-1)[:8]):
            md.dom_dp[i] = float(v)
        self._keep = {}
        for k, dt in (("rew_sa", np.float64), ("rew_as2", np.float64), ("term_sa", np.uint8),
                      ("term_as2", np.uint8), ("start_values", np.float32), ("start_table", np.int32)):
            v = d.get(k)
            if v is not None and np.size(v):
                self._keep[k] = np.ascontiguousarray(v, dt)
                setattr(md, k, ptr(self._keep[k]))
        md.action_draw = int(d.get("action_draw", 0))
        md.start_kind = int(d.get("start_kind", 0))
        for i, v in enumerate(np.asarray(d.get("start_ip", []), np.int32).reshape(-1)[:4]):
            md.start_ip[i] = int(v)
        md.start_total = float(d.get("start_total", 0.0))
        md.delta_capacity = int(d.get("delta_capacity", 0))
        md.dirichlet_sampling = int(d.get("dirichlet_sampling", 0))
        self.S, self.A, self.O, self.FS, self.FO = md.S, md.A, md.O, len(fs), len(fo)
        t_par = np.ascontiguousarray(t_par, np.uint32).reshape(-1, self.A * self.FS)
        o_par = np.ascontiguousarray(o_par, np.uint32).reshape(-1, self.A * self.FO)
        h = C.c_void_p()
        _check(ctx.h, self.L.fba_model_create(ctx.h, C.byref(md), int(max_structures or len(t_par)),
                                              C.byref(h)))
        self.h = h
        self.add_structures(t_par, o_par)

    def add_structures(self, t_par, o_par):
        t_par = np.ascontiguousarray(t_par, np.uint32).reshape(-1, self.A * self.FS)
        o_par = np.ascontiguousarray(o_par, np.uint32).reshape(-1, self.A * self.FO)
        ids = np.zeros(len(t_par), np.int32)
        _check(self.ctx.h, self.L.fba_model_add_structures(self.h, len(t_par), ptr(t_par), ptr(o_par),
                                                           ptr(ids)))
        return ids

    @property
    def num_structures(self):
        return self.L.fba_model_num_structures(self.h)

    def structure_size(self, i):
        return self.L.fba_model_structure_size(self.h, i)

    def structure(self, i):
        t = np.zeros(self.A * self.FS, np.uint32)
        o = np.zeros(self.A * self.FO, np.uint32)
        _check(self.ctx.h, self.L.fba_model_get_structure(self.h, i, ptr(t), ptr(o)))
        return t, o

    def max_structure_size(self):
        return max(self.structure_size(i) for i in range(self.num_structures))

    def close(self):
        if self.h:
            self.L.fba_model_destroy(self.h)
            self.h = None


class _ParticleBelief:
    """Shared plumbing of the belief kinds: one fba_belief handle."""
    _weighted = True

    def __init__(self):
        self.h = None
        self.sim = None

    def _create(self, sim, n, stride=0):
        self.sim, self.ctx, self.L = sim, sim.ctx, sim.L
        h = C.c_void_p()
        _check(self.ctx.h, self.L.fba_belief_create(self.ctx.h, sim.h, n, stride, int(self._weighted),
                                                    C.byref(h)))
        return h

    def _init(self, h, proto_struct_id, proto_counts, particle_proto, particle_state):
        stride = self.L.fba_belief_stride(h)
        psid = np.ascontiguousarray(proto_struct_id, np.int32).reshape(-1)
        pc = np.zeros((len(psid), stride), np.float32)
        src = np.asarray(proto_counts, np.float32).reshape(len(psid), -1)
        pc[:, :src.shape[1]] = src
        pp = None if particle_proto is None else np.ascontiguousarray(particle_proto, np.int32)
        ps = np.ascontiguousarray(particle_state, np.int32)
        _check(self.ctx.h, self.L.fba_belief_init(h, len(psid), ptr(psid), ptr(pc), ptr(pp), ptr(ps)))

    def _init_explicit(self, h, struct_id, counts, state):
        """Upload fully explicit particles (what the reference prior produced, one by one)."""
        n = self.L.fba_belief_size(h)
        stride = self.L.fba_belief_stride(h)
        c = np.zeros((n, stride), np.float32)
        src = np.asarray(counts, np.float32)
        c[:, :src.shape[1]] = src
        sid = np.ascontiguousarray(struct_id, np.int32)
        st = np.ascontiguousarray(state, np.int32)
        w = np.full(n, 1.0 / n, np.float64) if self._weighted else None
        _check(self.ctx.h, self.L.fba_belief_upload(h, 0, n, ptr(st), ptr(sid), ptr(c), ptr(w)))

    @staticmethod
    def _download(L, ctx, h, weighted, first=0, count=None, counts=True):
        n = L.fba_belief_size(h)
        count = n - first if count is None else count
        stride = L.fba_belief_stride(h)
        st = np.zeros(count, np.int32)
        sid = np.zeros(count, np.int32)
        c = np.zeros((count, stride), np.float32) if counts else None
        w = np.zeros(count, np.float64) if weighted else None
        _check(ctx.h, L.fba_belief_download(h, first, count, ptr(st), ptr(sid), ptr(c), ptr(w)))
        tw = C.c_double(0)
        L.fba_belief_total_weight(h, C.byref(tw))
        return dict(state=st, struct_id=sid, counts=c, w=w, total_weight=tw.value)

    def download(self, first=0, count=None, counts=True):
        return self._download(self.L, self.ctx, self.h, self._weighted, first, count, counts)

    def size(self):
        return self.L.fba_belief_size(self.h)

    def resample_stats(self):
        """(count blocks copied, resamples run) by the in-place resampler since creation."""
        c, r = C.c_int64(0), C.c_int64(0)
        _check(self.ctx.h, self.L.fba_belief_resample_stats(self.h, C.byref(c), C.byref(r)))
        return c.value, r.value

    def free(self, _simulator=None):
        """Belief::free (Belief.hpp:29)."""
        if self.h:
            self.L.fba_belief_destroy(self.h)
            self.h = None

    def sample(self, rng):
        """Belief::sample (Belief.hpp:34): index of the drawn particle."""
        i = C.c_int64(0)
        _check(self.ctx.h, self.L.fba_belief_sample(self.h, C.byref(rng), C.byref(i)))
        return i.value

    def replay_history(self, episode_len, actions, observations, rng, max_attempts=1_000_000):
        """computePosterior of the MH structure beliefs (MHNIPS2018.cpp:41-109) on every particle: the whole
        (action, observation) history replayed with episode retries; each particle's domain state becomes
        the state after the last step."""
        ln = np.ascontiguousarray(episode_len, np.int32)
        ac = np.ascontiguousarray(actions, np.int32)
        ob = np.ascontiguousarray(observations, np.int32)
        _check(self.ctx.h, self.L.fba_belief_replay_history(self.h, len(ln), ptr(ln), ptr(ac), ptr(ob), C.byref(rng),
                                                            int(max_attempts)))

    def sample_state_history(self, method, episode_len, actions, observations, rng, state_prior=None,
                             max_attempts=1_000_000):
        """MHwithinGibbs' sampleStateHistory (MHwithinGibbs.cpp:38-232) for every particle: method "msg" (backward
        messages + forward sampling; needs state_prior, S floats) or "rs" (rejection sampling) ->
        states [N, n_steps + n_episodes]; counts untouched"""
        ln = np.ascontiguousarray(episode_len, np.int32)
        ac = np.ascontiguousarray(actions, np.int32)
        ob = np.ascontiguousarray(observations, np.int32)
        sp = None if state_prior is None else np.ascontiguousarray(state_prior, np.float32)
        out = np.zeros((self.size(), int(ln.sum()) + len(ln)), np.int32)
        _check(self.ctx.h, self.L.fba_belief_sample_state_history(
            self.h, {"msg": 0, "rs": 1}[method], len(ln), ptr(ln), ptr(ac), ptr(ob), ptr(sp), C.byref(rng),
            int(max_attempts), ptr(out)))
        return out

    def add_history_counts(self, episode_len, actions, observations, states):
        """MHwithinGibbs::computePosteriorCounts (MHwithinGibbs.cpp:397-436): +1 per transition of the state
        history — states [N, L] (one history per particle) or [L] (one history shared by all particles)"""
        ln = np.ascontiguousarray(episode_len, np.int32)
        ac = np.ascontiguousarray(actions, np.int32)
        ob = np.ascontiguousarray(observations, np.int32)
        st = np.ascontiguousarray(states, np.int32)
        _check(self.ctx.h, self.L.fba_belief_add_history_counts(self.h, len(ln), ptr(ln), ptr(ac), ptr(ob), ptr(st),
                                                                int(st.ndim == 1)))

    def assign_from(self, first, src, src_index):
        """particles src[src_index[j]] -> self[first + j] (MHNIPS2018.cpp:241-246: accepted proposals)"""
        idx = np.ascontiguousarray(src_index, np.int64)
        _check(self.ctx.h, self.L.fba_belief_assign_from(self.h, int(first), src.h, len(idx), ptr(idx)))

    def resetDomainStateDistribution(self, rng):
        """BABelief::resetDomainStateDistribution (BABelief.hpp:30)."""
        _check(self.ctx.h, self.L.fba_belief_reset_domain_states(self.h, C.byref(rng)))


class BAImportanceSampling(_ParticleBelief):
    """beliefs::BAImportanceSampling (BAImportanceSampling.cpp:17-111)."""

    def __init__(self, n):
        super().__init__()
        if n < 1:  # BAImportanceSampling.cpp:19-22
            raise FbaError(capi.ERR_INVALID, "cannot initiate BAImportanceSampling with n " + str(n))
        self._n = n

    def initiate(self, simulator, *, struct_id=None, counts=None, state=None, proto_struct_id=None,
                 proto_counts=None, particle_proto=None, stride=0):
        """Belief::initiate with the particles the (host-side) prior produced: either explicit
        per-particle arrays, or prototypes + per-particle (prototype, start state)."""
        self.h = self._create(simulator, self._n, stride)
        if proto_struct_id is not None:
            self._init(self.h, proto_struct_id, proto_counts, particle_proto, state)
        else:
            self._init_explicit(self.h, struct_id, counts, state)

    def initiate_sampled(self, simulator, proto_struct_id, proto_counts, proto_probs, rng, stride=0):
        """initiate with start states (and prototypes) drawn on device (PHILOX only)."""
        self.h = self._create(simulator, self._n, stride)
        s = self.L.fba_belief_stride(self.h)
        psid = np.ascontiguousarray(proto_struct_id, np.int32).reshape(-1)
        pc = np.zeros((len(psid), s), np.float32)
        src = np.asarray(proto_counts, np.float32).reshape(len(psid), -1)
        pc[:, :src.shape[1]] = src
        pr = None if proto_probs is None else np.ascontiguousarray(proto_probs, np.float64)
        _check(self.ctx.h, self.L.fba_belief_init_sampled(self.h, len(psid), ptr(psid), ptr(pc), ptr(pr),
                                                          C.byref(rng)))

    def update(self, a, o, rng):
        """importance_sampling::update (ImportanceSampler.hpp:31-62); returns the step likelihood."""
        lik = C.c_double(0)
        _check(self.ctx.h, self.L.fba_belief_update(self.h, a, o, C.byref(rng), C.byref(lik)))
        return lik.value

    def resample(self, rng):
        """importance_sampling::resample (ImportanceSampler.hpp:71-94)."""
        _check(self.ctx.h, self.L.fba_belief_resample(self.h, C.byref(rng)))

    def updateEstimation(self, a, o, rng, want_likelihood=True):
        """BAImportanceSampling::updateEstimation (BAImportanceSampling.cpp:74-88). Without the
        likelihood read-back a PHILOX-mode call only enqueues work (no host sync)."""
        lik = C.c_double(0)
        _check(self.ctx.h, self.L.fba_belief_update_estimation(
            self.h, a, o, C.byref(rng), C.byref(lik) if want_likelihood else None))
        return lik.value if want_likelihood else None


class BARejectionSampling(_ParticleBelief):
    """beliefs::BARejectionSampling (BARejectionSampling.cpp:10-60)."""
    _weighted = False

    def __init__(self, n):
        super().__init__()
        if n < 1:  # BARejectionSampling.cpp:13-16
            raise FbaError(capi.ERR_INVALID, "cannot initiate RejectionSampling with n = " + str(n))
        self._n = n
        self.attempts = 0

    def initiate(self, simulator, *, struct_id=None, counts=None, state=None, proto_struct_id=None,
                 proto_counts=None, particle_proto=None, stride=0):
        self.h = self._create(simulator, self._n, stride)
        if proto_struct_id is not None:
            self._init(self.h, proto_struct_id, proto_counts, particle_proto, state)
        else:
            self._init_explicit(self.h, struct_id, counts, state)

    def updateEstimation(self, a, o, rng):
        """rejectSample (RejectionSampling.hpp:26-72)."""
        n = C.c_int64(0)
        _check(self.ctx.h, self.L.fba_belief_reject_sample(self.h, a, o, C.byref(rng), C.byref(n)))
        self.attempts = n.value
        return n.value


class ReinvigoratingRejectionSampling(_ParticleBelief):
    """beliefs::bayes_adaptive::factored::ReinvigoratingRejectionSampling: two flat filters, the
    learned-structure belief and the fully connected one."""
    _weighted = False

    def __init__(self, size, reinvigoration_amount, mutate_kind):
        super().__init__()
        if size < 1 or reinvigoration_amount < 1:  # ReinvigoratingRejectionSampling.cpp:43-48
            raise FbaError(capi.ERR_INVALID,
                           "ReinvigoratingRejectionSampling::cannot initiate belief of size < 1 ("
                           + str(size) + "), or resample size of < 1 (" + str(reinvigoration_amount) + ")")
        self._size, self._amount, self._mutate = size, reinvigoration_amount, mutate_kind
        self.fc = None

    def initiate(self, simulator, *, belief, fully_connected, stride):
        """belief / fully_connected: dicts with struct_id, counts, state (what
        sampleStartState / sampleFullyConnectedState produced on the host)."""
        self.h = self._create(simulator, self._size, stride)
        self.fc = self._create(simulator, self._size, stride)
        self._init_explicit(self.h, belief["struct_id"], belief["counts"], belief["state"])
        self._init_explicit(self.fc, fully_connected["struct_id"], fully_connected["counts"],
                            fully_connected["state"])

    def reinvigorateParticles(self, rng):
        _check(self.ctx.h, self.L.fba_belief_reinvigorate(self.h, self.fc, self._amount, self._mutate,
                                                          C.byref(rng)))

    def updateEstimation(self, a, o, rng):
        """ReinvigoratingRejectionSampling::updateEstimation (…:89-106)."""
        self.reinvigorateParticles(rng)
        n = C.c_int64(0)
        _check(self.ctx.h, self.L.fba_belief_reject_sample(self.h, a, o, C.byref(rng), C.byref(n)))
        _check(self.ctx.h, self.L.fba_belief_reject_sample(self.fc, a, o, C.byref(rng), C.byref(n)))

    def resetDomainStateDistribution(self, rng):
        _check(self.ctx.h, self.L.fba_belief_reset_domain_states(self.h, C.byref(rng)))
        _check(self.ctx.h, self.L.fba_belief_reset_domain_states(self.fc, C.byref(rng)))

    def download_fully_connected(self, counts=True):
        return self._download(self.L, self.ctx, self.fc, False, 0, None, counts)

    def free(self, _simulator=None):
        super().free()
        if self.fc:
            self.L.fba_belief_destroy(self.fc)
            self.fc = None


def log_bd_score(belief, prior):
    """BABNModel::LogBDScore of every particle of `belief` against `prior` (a belief of the same size
    or of one particle, same structures) -> N doubles."""
    out = np.zeros(belief.size(), np.float64)
    _check(belief.ctx.h, belief.L.fba_belief_log_bd_score(belief.h, prior.h, ptr(out)))
    return out


class SearchTree:
    """Device-resident POMCP search tree (fba_tree_*): planners::RBAPOUCT::selectAction
    (RBAPOUCT.cpp:67-153) with whole simulations inside one kernel, `wave` at a time."""

    def __init__(self, simulator, max_simulations, max_depth):
        self.sim, self.ctx, self.L = simulator, simulator.ctx, simulator.L
        h = C.c_void_p()
        _check(self.ctx.h, self.L.fba_tree_create(self.ctx.h, simulator.h, int(max_simulations), int(max_depth),
                                                  C.byref(h)))
        self.h = h

    def selectAction(self, belief, n_simulations, depth, u, discount, wave, rng):
        """-> (action, q[A], visits[A])"""
        act = C.c_int32(-1)
        q = np.zeros(self.sim.A, np.float64)
        n = np.zeros(self.sim.A, np.int64)
        _check(self.ctx.h, self.L.fba_tree_search(self.h, belief.h, int(n_simulations), int(depth), float(u),
                                                  float(discount), int(wave), C.byref(rng), C.byref(act), ptr(q),
                                                  ptr(n)))
        return int(act.value), q, n

    def free(self):
        if self.h:
            self.L.fba_tree_destroy(self.h)
            self.h = None


class BatchedBAImportanceSampling:
    """The importance-sampling beliefs of `n_runs` INDEPENDENT runs (the reference's `--runs`,
    src/experiments/BAPOMDPExperiment.cpp:32-78), `n` particles each, advanced together: every call
    is one kernel launch with one CTA per run, each run with its own action and observation. Run r is
    bit-identical to a stand-alone BAImportanceSampling(n) driven by Rng.philox(seed + r) through the
    same calls. `active`: optional boolean mask; runs with False are left untouched."""

    def __init__(self, n_runs, n):
        if n_runs < 1 or n < 1:
            raise FbaError(capi.ERR_INVALID, "cannot initiate %d runs of %d particles" % (n_runs, n))
        self.n_runs, self.n = int(n_runs), int(n)
        self.h = self.L = self.ctx = self.sim = None

    def initiate_sampled(self, simulator, proto_struct_id, proto_counts, proto_probs, rng, stride=0):
        self.sim, self.ctx, self.L = simulator, simulator.ctx, simulator.L
        if stride <= 0:
            stride = simulator.max_structure_size()
        h = C.c_void_p()
        _check(self.ctx.h, self.L.fba_runs_create(self.ctx.h, simulator.h, self.n_runs, self.n, stride, C.byref(h)))
        self.h = h
        self.storage = self.L.fba_runs_belief(self.h)
        s = self.L.fba_belief_stride(self.storage)
        psid = np.ascontiguousarray(proto_struct_id, np.int32).reshape(-1)
        pc = np.zeros((len(psid), s), np.float32)
        src = np.asarray(proto_counts, np.float32).reshape(len(psid), -1)
        pc[:, :src.shape[1]] = src
        pr = None if proto_probs is None else np.ascontiguousarray(proto_probs, np.float64)
        _check(self.ctx.h, self.L.fba_runs_init_sampled(self.h, len(psid), ptr(psid), ptr(pc), ptr(pr), C.byref(rng)))

    def _mask(self, active):
        if active is None:
            return None
        m = np.ascontiguousarray(active, np.uint8).reshape(-1)
        assert len(m) == self.n_runs
        return m

    def updateEstimation(self, a, o, rng, active=None, want_likelihood=True):
        a = np.ascontiguousarray(a, np.int32).reshape(-1)
        o = np.ascontiguousarray(o, np.int32).reshape(-1)
        assert len(a) == self.n_runs and len(o) == self.n_runs
        m = self._mask(active)
        lik = np.full(self.n_runs, np.nan) if want_likelihood else None
        _check(self.ctx.h, self.L.fba_runs_update_estimation(self.h, ptr(a), ptr(o), ptr(m), C.byref(rng), ptr(lik)))
        return lik

    def resetDomainStateDistribution(self, rng, active=None):
        _check(self.ctx.h, self.L.fba_runs_reset_domain_states(self.h, ptr(self._mask(active)), C.byref(rng)))

    def sample(self, rng, active=None):
        """one particle index per run (inside the run); -1 for inactive runs"""
        idx = np.full(self.n_runs, -1, np.int64)
        _check(self.ctx.h, self.L.fba_runs_sample(self.h, ptr(self._mask(active)), C.byref(rng), ptr(idx)))
        return idx

    def selectAction(self, n_simulations, depth, u, discount, rng, sims_per_wave=1, active=None):
        """Planner::selectAction of every run (one device tree per run) -> (action[R], q[R, A], visits[R, A])"""
        depth = np.ascontiguousarray(np.broadcast_to(np.asarray(depth, np.int32), (self.n_runs,)))
        act = np.full(self.n_runs, -1, np.int32)
        q = np.zeros((self.n_runs, self.sim.A), np.float64)
        n = np.zeros((self.n_runs, self.sim.A), np.int64)
        _check(self.ctx.h, self.L.fba_runs_plan(self.h, int(n_simulations), ptr(depth), float(u), float(discount),
                                                int(sims_per_wave), ptr(self._mask(active)), C.byref(rng), ptr(act),
                                                ptr(q), ptr(n)))
        return act, q, n

    def download(self, run, counts=True):
        """the particles of one run, as _ParticleBelief.download"""
        return _ParticleBelief._download(self.L, self.ctx, self.storage, True, run * self.n, self.n, counts)

    def copies(self):
        return int(self.L.fba_runs_copies(self.h))

    def free(self, _simulator=None):
        if self.h:
            self.L.fba_runs_destroy(self.h)
            self.h = None


def rollouts(belief, particle, start_state, depth, discount, rng, word_offset=None):
    """n x RBAPOUCT::rollout (RBAPOUCT.cpp:295-323) in one launch; returns the n returns."""
    p = np.ascontiguousarray(particle, np.int64)
    s = np.ascontiguousarray(start_state, np.int32)
    d = np.ascontiguousarray(depth, np.int32)
    off = None if word_offset is None else np.ascontiguousarray(word_offset, np.int64)
    out = np.zeros(len(p), np.float64)
    _check(belief.ctx.h, belief.L.fba_rollouts(belief.h, len(p), ptr(p), ptr(s), ptr(d), float(discount),
                                               C.byref(rng), ptr(off), ptr(out)))
    return out
