"""fba_pomdp_b200 — B200-native (sm_100a) particle-belief / rollout hot path of samkatt/fba-pomdp.

The product is libfba_b200.so (hand-written CUDA kernels behind the C ABI in
include/fba_pomdp_b200.h). This package is the thin host-side mirror of the reference's
Belief / BABelief / rollout interfaces used by the tests and the benchmark."""
from . import capi  # noqa: F401
from .beliefs import (BAImportanceSampling, BAPOMDP, BARejectionSampling, BatchedBAImportanceSampling,  # noqa: F401
                      Context, ReinvigoratingRejectionSampling, SearchTree, log_bd_score, rollouts)
from .capi import FbaError, Rng  # noqa: F401

__all__ = ["capi", "Context", "BAPOMDP", "BAImportanceSampling", "BARejectionSampling", "BatchedBAImportanceSampling", "SearchTree", "log_bd_score",
           "ReinvigoratingRejectionSampling", "rollouts", "Rng", "FbaError"]
from .sharded import ShardedBAImportanceSampling, exchange_plan, offspring_quotas  # noqa: F401,E402
from .sharded import exchange_records, gather_ragged, split_requests  # noqa: F401,E402
