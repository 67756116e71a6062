"""ctypes binding of include/fba_pomdp_b200.h (libfba_b200.so). Nothing here computes: every call
goes to the CUDA library, and a missing library or missing GPU raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FBA_B200_LIB") or os.path.join(HERE, "libfba_b200.so")

MAX_FEATURES = 16

OK, ERR_INVALID, ERR_CUDA, ERR_RNG_UNDERRUN, ERR_CAPACITY, ERR_NO_DEVICE = range(6)
DOM_TABLE, DOM_TIGER, DOM_FACTORED_TIGER, DOM_SYSADMIN, DOM_GRIDWORLD, DOM_COLLISION_AVOIDANCE = range(6)
ACT_UNIFORM_INT, ACT_SLOW_INT = 0, 1
START_CONST, START_BOOL, START_UNIFORM_INT, START_SLOW2, START_CATEGORICAL = range(5)
MUT_FACTORED_TIGER, MUT_COLLISION_AVOIDANCE, MUT_SYSADMIN, MUT_GRIDWORLD = range(4)
RNG_REPLAY, RNG_PHILOX = 0, 1

P2P_BLOB_BYTES = 320  # FBA_P2P_BLOB_BYTES

# every symbol include/fba_pomdp_b200.h declares (tests check the library exports all of them)
ABI_VERSION = 12  # must equal FBA_ABI_VERSION in include/fba_pomdp_b200.h

SYMBOLS = [
    "fba_abi_version", "fba_ctx_create", "fba_ctx_destroy", "fba_last_error", "fba_ctx_stream", "fba_ctx_synchronize",
    "fba_ctx_launch_count", "fba_ctx_set_option", "fba_ctx_profile_begin", "fba_ctx_profile_end", "fba_ctx_profile_get", "fba_ctx_profile_list",
    "fba_model_create", "fba_model_destroy", "fba_model_add_structures",
    "fba_model_num_structures", "fba_model_structure_size", "fba_model_get_structure",
    "fba_belief_create", "fba_belief_destroy", "fba_belief_size", "fba_belief_stride",
    "fba_belief_init", "fba_belief_init_sampled", "fba_belief_upload", "fba_belief_download",
    "fba_belief_total_weight", "fba_belief_update", "fba_belief_resample",
    "fba_belief_update_estimation", "fba_belief_reset_domain_states", "fba_belief_sample",
    "fba_belief_reject_sample", "fba_belief_reinvigorate", "fba_rollouts",
    "fba_belief_sample_batch", "fba_belief_gather_states", "fba_step_batch", "fba_belief_propose",
    "fba_belief_normalize", "fba_belief_resample_shard", "fba_belief_resample_stats",
    "fba_belief_shard_resample", "fba_belief_shard_resample_async", "fba_belief_shard_plan",
    "fba_belief_p2p_export", "fba_belief_p2p_open", "fba_belief_sharded_update", "fba_belief_p2p_timeouts",
    "fba_belief_p2p_set_timeout", "fba_ctx_counter", "fba_belief_replay_history", "fba_belief_assign_from", "fba_belief_aux_ptr", "fba_belief_chain_recomputed",
    "fba_belief_dropped_records", "fba_belief_reserve_export", "fba_belief_import_from", "fba_belief_export_count",
    "fba_belief_export_ptr", "fba_belief_import_ptr", "fba_belief_record_bytes", "fba_belief_import",
    "fba_belief_counts_ptr", "fba_belief_state_ptr", "fba_belief_weight_ptr",
    "fba_belief_scalars_ptr",
    "fba_belief_log_bd_score",
    "fba_tree_create", "fba_tree_destroy", "fba_tree_search",
    "fba_runs_create", "fba_runs_destroy", "fba_runs_belief", "fba_runs_init_sampled",
    "fba_runs_update_estimation", "fba_runs_reset_domain_states", "fba_runs_sample", "fba_runs_copies", "fba_runs_plan", "fba_runs_init",
    "fba_belief_replace_from", "fba_belief_cheat", "fba_belief_breed_into", "fba_belief_least_likely", "fba_belief_promote", "fba_belief_compact", "fba_belief_delta_capacity", "fba_belief_redraw_domain_states", "fba_belief_sample_state_history", "fba_belief_add_history_counts",
    "fba_nested_create", "fba_nested_destroy", "fba_nested_top", "fba_nested_bottom_size", "fba_nested_upload_states",
    "fba_nested_download_states", "fba_nested_reset_domain_states", "fba_nested_update", "fba_nested_sample",
]


class ModelDesc(C.Structure):
    _fields_ = [
        ("S", C.c_int32), ("A", C.c_int32), ("O", C.c_int32),
        ("n_state_features", C.c_int32), ("n_obs_features", C.c_int32),
        ("state_feature_sizes", C.c_int32 * MAX_FEATURES),
        ("obs_feature_sizes", C.c_int32 * MAX_FEATURES),
        ("tabular", C.c_int32), ("domain", C.c_int32),
        ("dom_ip", C.c_int32 * 32), ("dom_dp", C.c_double * 8),
        ("rew_sa", C.c_void_p), ("rew_as2", C.c_void_p), ("term_sa", C.c_void_p),
        ("term_as2", C.c_void_p),
        ("action_draw", C.c_int32), ("start_kind", C.c_int32), ("start_ip", C.c_int32 * 4),
        ("start_values", C.c_void_p), ("start_total", C.c_double), ("start_table", C.c_void_p),
        ("delta_capacity", C.c_int32), ("dirichlet_sampling", C.c_int32),
    ]


class Rng(C.Structure):
    """fba_rng. Replay: Rng.replay(words); native: Rng.philox(seed)."""
    _fields_ = [("mode", C.c_int32), ("words", C.c_void_p), ("n_words", C.c_int64),
                ("cursor", C.c_int64), ("seed", C.c_uint64), ("offset", C.c_uint64)]

    @classmethod
    def replay(cls, words):
        w = np.ascontiguousarray(words, np.uint32)
        r = cls(RNG_REPLAY, w.ctypes.data_as(C.c_void_p), len(w), 0, 0, 0)
        r._keep = w
        return r

    @classmethod
    def philox(cls, seed, offset=0):
        return cls(RNG_PHILOX, None, 0, 0, seed, offset)

    @property
    def exhausted(self):
        return self.cursor == self.n_words


class FbaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


_lib = None


def lib():
    """Loads libfba_b200.so; fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FbaError(ERR_NO_DEVICE, "%s is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        pp = C.POINTER(vp)
        sig = {
            "fba_abi_version": (C.c_int, []),
            "fba_ctx_create": (C.c_int, [C.c_int, pp]),
            "fba_ctx_destroy": (None, [vp]),
            "fba_last_error": (C.c_char_p, [vp]),
            "fba_ctx_stream": (vp, [vp]),
            "fba_ctx_synchronize": (C.c_int, [vp]),
            "fba_ctx_launch_count": (i64, [vp]),
            "fba_ctx_set_option": (C.c_int, [vp, C.c_char_p, i64]),
            "fba_ctx_profile_begin": (C.c_int, [vp]),
            "fba_ctx_profile_end": (C.c_int, [vp]),
            "fba_ctx_profile_get": (C.c_int, [vp, C.c_char_p, vp, vp]),
            "fba_ctx_profile_list": (i64, [vp, vp, i64]),
            "fba_belief_resample_stats": (C.c_int, [vp, vp, vp]),
            "fba_belief_shard_resample": (C.c_int, [vp, vp, i32, i32, dbl, vp, vp, vp]),
            "fba_belief_shard_resample_async": (C.c_int, [vp, vp, i32, i32, dbl, vp]),
            "fba_belief_shard_plan": (C.c_int, [vp, vp, i32, i32, dbl, vp, vp]),
            "fba_ctx_counter": (i64, [vp, i32]),
            "fba_belief_replay_history": (C.c_int, [vp, i32, vp, vp, vp, vp, i64]),
            "fba_belief_assign_from": (C.c_int, [vp, i64, vp, i64, vp]),
            "fba_belief_aux_ptr": (vp, [vp]),
            "fba_belief_chain_recomputed": (i64, [vp]),
            "fba_belief_p2p_export": (C.c_int, [vp, vp]),
            "fba_belief_p2p_open": (C.c_int, [vp, vp, i32, i32]),
            "fba_belief_sharded_update": (C.c_int, [vp, i32, i32, vp, dbl, vp]),
            "fba_belief_p2p_timeouts": (i64, [vp]),
            "fba_belief_p2p_set_timeout": (C.c_int, [vp, dbl]),
            "fba_belief_dropped_records": (i64, [vp]),
            "fba_belief_reserve_export": (C.c_int, [vp, i64]),
            "fba_belief_import_from": (C.c_int, [vp, vp, i64]),
            "fba_model_create": (C.c_int, [vp, vp, i32, pp]),
            "fba_model_destroy": (None, [vp]),
            "fba_model_add_structures": (C.c_int, [vp, i32, vp, vp, vp]),
            "fba_model_num_structures": (i32, [vp]),
            "fba_model_structure_size": (i64, [vp, i32]),
            "fba_model_get_structure": (C.c_int, [vp, i32, vp, vp]),
            "fba_belief_create": (C.c_int, [vp, vp, i64, i64, i32, pp]),
            "fba_belief_destroy": (None, [vp]),
            "fba_belief_size": (i64, [vp]),
            "fba_belief_stride": (i64, [vp]),
            "fba_belief_init": (C.c_int, [vp, i32, vp, vp, vp, vp]),
            "fba_belief_init_sampled": (C.c_int, [vp, i32, vp, vp, vp, vp]),
            "fba_belief_upload": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
            "fba_belief_download": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
            "fba_belief_total_weight": (C.c_int, [vp, vp]),
            "fba_belief_update": (C.c_int, [vp, i32, i32, vp, vp]),
            "fba_belief_resample": (C.c_int, [vp, vp]),
            "fba_belief_update_estimation": (C.c_int, [vp, i32, i32, vp, vp]),
            "fba_belief_reset_domain_states": (C.c_int, [vp, vp]),
            "fba_belief_sample": (C.c_int, [vp, vp, vp]),
            "fba_belief_reject_sample": (C.c_int, [vp, i32, i32, vp, vp]),
            "fba_belief_reinvigorate": (C.c_int, [vp, vp, i64, i32, vp]),
            "fba_rollouts": (C.c_int, [vp, i64, vp, vp, vp, dbl, vp, vp, vp]),
            "fba_belief_sample_batch": (C.c_int, [vp, vp, i64, vp]),
            "fba_belief_gather_states": (C.c_int, [vp, i64, vp, vp]),
            "fba_step_batch": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]),
            "fba_belief_propose": (C.c_int, [vp, i32, i32, vp, vp]),
            "fba_belief_normalize": (C.c_int, [vp, dbl]),
            "fba_belief_resample_shard": (C.c_int, [vp, i64, vp]),
            "fba_belief_export_count": (i64, [vp]),
            "fba_belief_export_ptr": (vp, [vp]),
            "fba_belief_import_ptr": (vp, [vp, i64]),
            "fba_belief_record_bytes": (i64, [vp]),
            "fba_belief_import": (C.c_int, [vp, i64]),
            "fba_belief_counts_ptr": (vp, [vp]),
            "fba_belief_state_ptr": (vp, [vp]),
            "fba_belief_weight_ptr": (vp, [vp]),
            "fba_belief_scalars_ptr": (vp, [vp]),
            "fba_belief_log_bd_score": (C.c_int, [vp, vp, vp]),
            "fba_tree_create": (C.c_int, [vp, vp, i64, i32, pp]),
            "fba_tree_destroy": (None, [vp]),
            "fba_tree_search": (C.c_int, [vp, vp, i64, i32, dbl, dbl, i32, vp, vp, vp, vp]),
            "fba_runs_create": (C.c_int, [vp, vp, i32, i64, i64, pp]),
            "fba_runs_destroy": (None, [vp]),
            "fba_runs_belief": (vp, [vp]),
            "fba_runs_init_sampled": (C.c_int, [vp, i32, vp, vp, vp, vp]),
            "fba_runs_update_estimation": (C.c_int, [vp, vp, vp, vp, vp, vp]),
            "fba_runs_reset_domain_states": (C.c_int, [vp, vp, vp]),
            "fba_runs_sample": (C.c_int, [vp, vp, vp, vp]),
            "fba_runs_copies": (i64, [vp]),
            "fba_runs_init": (C.c_int, [vp, i32, vp, vp, vp, vp]),
            "fba_runs_plan": (C.c_int, [vp, i64, vp, dbl, dbl, i32, vp, vp, vp, vp, vp]),
            "fba_belief_replace_from": (C.c_int, [vp, vp, vp, vp, i64]),
            "fba_belief_cheat": (C.c_int, [vp, vp, i64, vp]),
            "fba_belief_breed_into": (C.c_int, [vp, vp, i64, vp, vp, i32, vp]),
            "fba_belief_least_likely": (C.c_int, [vp, i64, vp]),
            "fba_belief_promote": (C.c_int, [vp, vp, dbl, vp, vp]),
            "fba_belief_redraw_domain_states": (C.c_int, [vp, vp]),
            "fba_belief_compact": (C.c_int, [vp]),
            "fba_belief_delta_capacity": (i32, [vp]),
            "fba_belief_sample_state_history": (C.c_int, [vp, i32, i32, vp, vp, vp, vp, vp, i64, vp]),
            "fba_belief_add_history_counts": (C.c_int, [vp, i32, vp, vp, vp, vp, i32]),
            "fba_nested_create": (C.c_int, [vp, vp, i64, i64, i64, pp]),
            "fba_nested_destroy": (None, [vp]),
            "fba_nested_top": (vp, [vp]),
            "fba_nested_bottom_size": (i64, [vp]),
            "fba_nested_upload_states": (C.c_int, [vp, i64, i64, vp]),
            "fba_nested_download_states": (C.c_int, [vp, i64, i64, vp]),
            "fba_nested_reset_domain_states": (C.c_int, [vp, vp]),
            "fba_nested_update": (C.c_int, [vp, i32, i32, vp, i64, vp]),
            "fba_nested_sample": (C.c_int, [vp, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.fba_abi_version() != ABI_VERSION:
            raise FbaError(ERR_INVALID, "libfba_b200.so has ABI version %d, this binding expects %d: rebuild"
                           % (L.fba_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
