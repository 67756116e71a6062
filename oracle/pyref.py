"""TEST INFRASTRUCTURE: ctypes window onto oracle/_ref/libfba_ref.so (the unmodified reference,
compiled by oracle/Makefile + oracle/ref_harness.cpp). Only tests/, bench.py's reference/cpu_baseline
legs and oracle/gen_golden.py may import this. Never on the product path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfba_ref.so")

F_IS, F_RS, F_REINV, F_REINV_FC, F_MH = 0, 1, 2, 3, 4
F_CHEAT, F_CHEAT_CORRECT, F_INC, F_INC_FC, F_INC_SHADOW, F_GIBBS, F_NESTED = 5, 6, 7, 8, 9, 10, 11


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_open_ex.restype = C.c_void_p
        L.ref_open_ex.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                  C.c_double, C.c_int, C.c_char_p, C.c_int]
        L.ref_error.restype = C.c_char_p
        L.ref_error.argtypes = [C.c_void_p]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_reseed.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_feature_sizes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_rng_mark.argtypes = [C.c_void_p]
        L.ref_rng_words_since_mark.restype = C.c_long
        L.ref_rng_words_since_mark.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long]
        L.ref_rng_draw_words.argtypes = [C.c_void_p, C.c_long]
        L.ref_belief_init.restype = C.c_int
        L.ref_belief_init.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_long]
        L.ref_filter_size.restype = C.c_long
        L.ref_filter_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_filter_states.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_is_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_particle_num_counts.restype = C.c_long
        L.ref_particle_num_counts.argtypes = [C.c_void_p, C.c_int, C.c_long]
        L.ref_particle_dump.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        L.ref_is_update.restype = C.c_double
        L.ref_is_update.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_is_resample.argtypes = [C.c_void_p]
        L.ref_update_estimation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_reset_domain_states.argtypes = [C.c_void_p, C.c_int]
        L.ref_reinvigorate_only.argtypes = [C.c_void_p]
        L.ref_particle_step.restype = C.c_double
        L.ref_particle_step.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_int, C.c_void_p]
        L.ref_particle_obs_prob.restype = C.c_double
        L.ref_particle_obs_prob.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_int]
        L.ref_rollout.restype = C.c_double
        L.ref_rollout.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_int]
        L.ref_reward.restype = C.c_double
        L.ref_reward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_env_script.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
        L.ref_adapter_episodes.restype = C.c_int
        L.ref_adapter_episodes.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_char_p, C.c_int, C.c_int,
                                           C.c_void_p]
        L.ref_plan_seconds.restype = C.c_double
        L.ref_plan_seconds.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_char_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Ref:
    """One reference simulator (BAPOMDP or FBAPOMDP) + its beliefs. NOTE: the reference keeps ONE
    global RNG, so only one Ref should be active at a time."""

    def __init__(self, domain, size=0, width=0, height=0, factored=False, structure_prior="",
                 discount=0.95, horizon=20, seed="42", sampled=False):
        self.L = lib()
        self.h = self.L.ref_open_ex(domain.encode(), size, width, height, int(factored),
                                    structure_prior.encode(), discount, horizon, seed.encode(), int(sampled))
        err = self.L.ref_error(self.h).decode()
        if err:
            raise RuntimeError("reference: " + err)
        sz = np.zeros(5, np.int32)
        self.L.ref_sizes(self.h, _p(sz))
        self.S, self.A, self.O, self.FS, self.FO = (int(x) for x in sz)
        fs, fo = np.zeros(self.FS, np.int32), np.zeros(self.FO, np.int32)
        self.L.ref_feature_sizes(self.h, _p(fs), _p(fo))
        self.feat_s, self.feat_o = fs, fo
        self.factored = factored

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    def reseed(self, seed):
        self.L.ref_reseed(self.h, seed.encode())

    # RNG tap
    def mark(self):
        self.L.ref_rng_mark(self.h)

    def words_since_mark(self, limit=1 << 31):
        n = self.L.ref_rng_words_since_mark(self.h, None, 0, limit)
        if n < 0:
            raise RuntimeError("rng tap lost sync")
        out = np.zeros(n, np.uint32)
        if n:
            self.L.ref_rng_words_since_mark(self.h, _p(out), n, limit)
        return out

    def draw_words(self, n):
        out = np.zeros(n, np.uint32)
        self.L.ref_rng_draw_words(_p(out), n)
        return out

    # beliefs
    def belief_init(self, kind, n, resample_amount=0):
        if self.L.ref_belief_init(self.h, kind, n, resample_amount):
            raise RuntimeError("reference: " + self.L.ref_error(self.h).decode())

    def size(self, filt):
        return self.L.ref_filter_size(self.h, filt)

    def states(self, filt):
        out = np.zeros(self.size(filt), np.int32)
        self.L.ref_filter_states(self.h, filt, _p(out))
        return out

    def is_weights(self):
        w = np.zeros(self.size(F_IS), np.float64)
        tot = np.zeros(1, np.float64)
        self.L.ref_is_weights(self.h, _p(w), _p(tot))
        return w, float(tot[0])

    def particle(self, filt, i):
        """(t_parent_masks [A,FS], o_parent_masks [A,FO], counts float32[C])"""
        n = self.L.ref_particle_num_counts(self.h, filt, i)
        tp = np.zeros((self.A, self.FS), np.uint32)
        op = np.zeros((self.A, self.FO), np.uint32)
        c = np.zeros(n, np.float32)
        self.L.ref_particle_dump(self.h, filt, i, _p(tp), _p(op), _p(c))
        return tp, op, c

    def particles(self, filt):
        return [self.particle(filt, i) for i in range(self.size(filt))]

    def is_update(self, a, o):
        return self.L.ref_is_update(self.h, a, o)

    def is_resample(self):
        self.L.ref_is_resample(self.h)

    def update_estimation(self, kind, a, o):
        self.L.ref_update_estimation(self.h, kind, a, o)

    def reset_domain_states(self, kind):
        self.L.ref_reset_domain_states(self.h, kind)

    def reinvigorate_only(self):
        self.L.ref_reinvigorate_only(self.h)

    def particle_step(self, filt, i, a, keep_counts=False):
        out = np.zeros(3, np.int32)
        r = self.L.ref_particle_step(self.h, filt, i, a, int(keep_counts), _p(out))
        return int(out[0]), int(out[1]), bool(out[2]), r

    def obs_prob(self, filt, i, a, o):
        return self.L.ref_particle_obs_prob(self.h, filt, i, a, o)

    def log_bd_score(self, filt, i, prior_filt, j):
        self.L.ref_log_bd_score.restype = C.c_double
        self.L.ref_log_bd_score.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_long]
        return self.L.ref_log_bd_score(self.h, filt, i, prior_filt, j)

    def rollout(self, filt, i, start_state, depth):
        return self.L.ref_rollout(self.h, filt, i, start_state, depth)

    def reward(self, s, a, s2):
        t = np.zeros(1, np.int32)
        r = self.L.ref_reward(self.h, s, a, s2, _p(t))
        return r, bool(t[0])

    def env_script(self, steps, horizon):
        a = np.zeros(steps, np.int32)
        o = np.zeros(steps, np.int32)
        f = np.zeros(steps, np.int32)
        self.L.ref_env_script(self.h, steps, horizon, _p(a), _p(o), _p(f))
        return a, o, f

    def adapter_episodes(self, kind, n, planner="po-uct", sims=64, episodes=10):
        """The reference's own episode loop + planner with belief `kind` (0: reference IS, 1: this
        repo's CudaBAImportanceSampling, 2: reference RS, 3: CudaBARejectionSampling)."""
        out = np.zeros(episodes, np.float64)
        if self.L.ref_adapter_episodes(self.h, kind, n, planner.encode(), sims, episodes, _p(out)):
            raise RuntimeError("reference/adapter: " + self.L.ref_error(self.h).decode())
        return out

    def adapter_events(self):
        """MH runs / Gibbs chains / cheats of the CUDA belief in the last adapter_episodes call (-1: not such a belief)"""
        self.L.ref_adapter_events.restype = C.c_long
        self.L.ref_adapter_events.argtypes = [C.c_void_p]
        return self.L.ref_adapter_events(self.h)

    def adapter_update_seconds(self):
        """(seconds inside Belief::updateEstimation, calls) of the last adapter_episodes call"""
        f = self.L.ref_adapter_update_seconds
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_void_p]
        n = C.c_long(0)
        s = f(self.h, C.byref(n))
        return s, n.value

    def adapter_update_times(self):
        """seconds of every Belief::updateEstimation call of the last adapter_episodes call"""
        f = self.L.ref_adapter_update_times
        f.restype = C.c_long
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        n = f(self.h, None, 0)
        out = np.zeros(max(n, 1), np.float64)
        f(self.h, _p(out), n)
        return out[:n]

    def batched_episodes(self, n, runs, sims, episodes, sims_per_wave=1, device=0, seed=4711):
        """fba_b200::runBatchedExperiment on GPU `device` -> (returns[episodes, runs], seconds)"""
        f = self.L.ref_batched_episodes_on
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_void_p]
        out = np.zeros((episodes, runs), np.float64)
        dt = f(self.h, n, runs, sims, episodes, sims_per_wave, device, seed, _p(out))
        if dt < 0:
            raise RuntimeError("reference/adapter: " + self.L.ref_error(self.h).decode())
        return out, dt

    # ---- MHNIPS2018 ----
    def mh_init(self, n, threshold=-1e300):
        self.L.ref_mh_init.restype = C.c_int
        self.L.ref_mh_init.argtypes = [C.c_void_p, C.c_long, C.c_double]
        if self.L.ref_mh_init(self.h, n, threshold):
            raise RuntimeError(self.L.ref_error(self.h).decode())

    def mh_run(self):
        """the private MHNIPS2018::MH on the belief as it is"""
        self.L.ref_mh_run.argtypes = [C.c_void_p]
        self.L.ref_mh_run(self.h)

    def mh_log_likelihood(self):
        self.L.ref_mh_log_likelihood.restype = C.c_double
        self.L.ref_mh_log_likelihood.argtypes = [C.c_void_p]
        return self.L.ref_mh_log_likelihood(self.h)

    def prior_model(self, t_par, o_par, cap=1 << 20):
        """FBAPOMDPPrior::computePriorModel(structure) -> counts in this repo's layout"""
        f = self.L.ref_prior_model
        f.restype = C.c_long
        f.argtypes = [C.c_void_p] * 4
        tp, op = np.ascontiguousarray(t_par, np.uint32), np.ascontiguousarray(o_par, np.uint32)
        out = np.zeros(cap, np.float32)
        n = f(self.h, _p(tp), _p(op), _p(out))
        if n < 0:
            raise RuntimeError(self.L.ref_error(self.h).decode())
        return out[:n].copy()

    def batched_experiment_file(self, n, runs, sims, episodes, path, seed=4711):
        """runBatchedExperiment writing the reference's result file -> (returns[episodes, runs], seconds)"""
        f = self.L.ref_batched_experiment_file
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_char_p, C.c_void_p]
        out = np.zeros((episodes, runs), np.float64)
        dt = f(self.h, n, runs, sims, episodes, seed, str(path).encode(), _p(out))
        if dt < 0:
            raise RuntimeError(self.L.ref_error(self.h).decode())
        return out, dt

    def adapter_initiate(self, n, n_ref):
        """CudaBAImportanceSampling(n).initiate next to n_ref reference sampleStartState draws ->
        dict(seconds, host_samples, cuda_state, cuda_sid, ref_state, ref_sid)"""
        f = self.L.ref_adapter_initiate
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_long, C.c_long] + [C.c_void_p] * 6
        sec, hs = C.c_double(0), C.c_long(0)
        cs, ci = np.zeros(n, np.int32), np.zeros(n, np.int32)
        rs, ri = np.zeros(n_ref, np.int32), np.zeros(n_ref, np.int32)
        if f(self.h, n, n_ref, C.byref(sec), C.byref(hs), _p(cs), _p(ci), _p(rs), _p(ri)):
            raise RuntimeError(self.L.ref_error(self.h).decode())
        return dict(seconds=sec.value, host_samples=hs.value, cuda_state=cs, cuda_sid=ci, ref_state=rs, ref_sid=ri)

    # ---- composite structure beliefs (CheatingReinvigoration, StructureIncubatorSampling, MHwithinGibbs, NestedBelief)
    def composite_init(self, kind, n, amount, threshold):
        f = self.L.ref_composite_init
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_double]
        if f(self.h, kind, n, amount, threshold):
            raise RuntimeError("reference: " + self.L.ref_error(self.h).decode())

    def composite_update(self, kind, a, o):
        self.L.ref_composite_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.L.ref_composite_update(self.h, kind, a, o)

    def composite_reset(self, kind):
        self.L.ref_composite_reset.argtypes = [C.c_void_p, C.c_int]
        self.L.ref_composite_reset(self.h, kind)

    def weights(self, filt):
        """(weights, _total_weight) of a weighted filter"""
        self.L.ref_filter_weights.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        w = np.zeros(self.size(filt), np.float64)
        tot = np.zeros(1, np.float64)
        self.L.ref_filter_weights(self.h, filt, _p(w), _p(tot))
        return w, float(tot[0])

    def cheat_likelihood(self):
        self.L.ref_cheat_likelihood.restype = C.c_double
        self.L.ref_cheat_likelihood.argtypes = [C.c_void_p]
        return self.L.ref_cheat_likelihood(self.h)

    def cheat_only(self):
        self.L.ref_cheat_only.argtypes = [C.c_void_p]
        self.L.ref_cheat_only(self.h)

    def incubator_part(self, part, a=0, o=0):
        f = self.L.ref_incubator_part
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        return f(self.h, part, a, o)

    def incubator_set_shadow_weights(self, w):
        w = np.ascontiguousarray(w, np.float64)
        self.L.ref_incubator_set_shadow_weights.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_incubator_set_shadow_weights(self.h, _p(w))

    def nested_states(self, n_top, n_bottom):
        out = np.zeros((n_top, n_bottom), np.int32)
        self.L.ref_nested_states.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_nested_states(self.h, _p(out))
        return out

    def nested_sample(self):
        out = np.zeros(2, np.int32)
        self.L.ref_nested_sample.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_nested_sample(self.h, _p(out))
        return int(out[0]), int(out[1])

    def mutate(self, t_par, o_par):
        """FBAPOMDP::mutate on one structure (parent bitmasks) -> (t_par, o_par) of the mutated structure"""
        f = self.L.ref_mutate
        f.restype = C.c_int
        f.argtypes = [C.c_void_p] * 5
        tp, op = np.ascontiguousarray(t_par, np.uint32), np.ascontiguousarray(o_par, np.uint32)
        to, oo = np.zeros_like(tp), np.zeros_like(op)
        if f(self.h, _p(tp), _p(op), _p(to), _p(oo)):
            raise RuntimeError(self.L.ref_error(self.h).decode())
        return to, oo

    def gibbs_run(self):
        """the private MHwithinGibbs::reinvigorate on the belief as it is"""
        self.L.ref_gibbs_run.argtypes = [C.c_void_p]
        self.L.ref_gibbs_run(self.h)

    def gibbs_log_likelihood(self):
        self.L.ref_gibbs_log_likelihood.restype = C.c_double
        self.L.ref_gibbs_log_likelihood.argtypes = [C.c_void_p]
        return self.L.ref_gibbs_log_likelihood(self.h)

    def state_prior(self):
        out = np.zeros(self.S, np.float32)
        self.L.ref_state_prior.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_state_prior(self.h, _p(out))
        return out

    def plan_seconds(self, kind, n, planner, sims, reps=3):
        """Seconds per Planner::selectAction with `sims` simulations (empty history)."""
        v = self.L.ref_plan_seconds(self.h, kind, n, planner.encode(), sims, reps)
        if v < 0:
            raise RuntimeError("reference/adapter: " + self.L.ref_error(self.h).decode())
        return v
