/* fba_pomdp_b200 — C ABI of the B200 (sm_100a) particle-belief / rollout hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes, no
 * exceptions, no torch / CUDA types. Each entry point names the reference (samkatt/fba-pomdp)
 * interface it stands in for; paths are relative to the reference's root.
 *
 * Vocabulary
 *   particle   one BA-/FBA-POMDP hyper-state: a domain state index, a structure id and a private
 *              block of float32 Dirichlet counts (BAPOMDPState / FBAPOMDPState).
 *   structure  the parent sets of every DBN node, as bitmasks over state features. A tabular
 *              BA-POMDP is the one-structure case: one state feature of size S, parent mask 1.
 *   count block layout: for a in [0,A): transition nodes f = 0..FS-1, then observation nodes
 *              g = 0..FO-1; each node's CPT row-major [parent configuration][output], parent
 *              configuration = mixed radix over the node's parents in ascending feature order.
 *              Tabular: T(a)[s][s'] = phi[s][a][s'], O(a)[s'][o] = psi[a][s'][o].
 *   belief     N particles on one GPU (+ double weights when weighted).
 *   rng        REPLAY: the exact 32-bit words the reference's std::mt19937 would produce, consumed
 *              through the same libstdc++ distributions in the same order (bit-exact parity mode);
 *              PHILOX: counter-based Philox4x32-10 on device (production mode).
 *
 * Threading: one host thread drives one context; every call is synchronous on return unless noted.
 */
#ifndef FBA_POMDP_B200_H
#define FBA_POMDP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBA_MAX_FEATURES 16
/* bumped whenever a struct below changes layout; compare with fba_abi_version() after loading */
#define FBA_ABI_VERSION 12

typedef struct fba_ctx fba_ctx;
typedef struct fba_model fba_model;
typedef struct fba_belief fba_belief;

enum fba_status {
    FBA_OK = 0,
    FBA_ERR_INVALID = 1,      /* bad argument (the reference throws a string here) */
    FBA_ERR_CUDA = 2,         /* CUDA runtime failure; see fba_last_error */
    FBA_ERR_RNG_UNDERRUN = 3, /* replay stream too short for the operation */
    FBA_ERR_CAPACITY = 4,     /* structure table or count stride too small */
    FBA_ERR_NO_DEVICE = 5     /* no CUDA device: there is no CPU fallback */
};

/* BADomainExtension::{reward,terminal} and POMDP::generateRandomAction, per domain
 * (src/bayes-adaptive/models/table/BADomainExtension.hpp:22-47) */
enum fba_domain_kind {
    FBA_DOM_TABLE = 0,          /* reward = rew_sa[s*A+a] + rew_as2[a*S+s'], terminal = OR of the two */
    FBA_DOM_TIGER = 1,          /* ip[0] = episodic          src/domains/tiger/TigerBAExtension.cpp:21-44 */
    FBA_DOM_FACTORED_TIGER = 2, /* ip[0] = episodic  src/domains/tiger/FactoredTigerBAExtension.cpp:27-56 */
    FBA_DOM_SYSADMIN = 3,       /* ip[0] = #computers, dp[0] = reboot cost
                                   src/domains/sysadmin/SysAdminBAExtension.cpp:27-48 */
    FBA_DOM_GRIDWORLD = 4,      /* ip[0] = size, ip[1] = #goals, ip[2+2g], ip[3+2g] = goal g (x,y);
                                   dp[0] = goal reward, dp[1] = step reward
                                   src/domains/gridworld/GridWorldBAExtension.cpp:74-100 */
    FBA_DOM_COLLISION_AVOIDANCE = 5 /* ip[0] = width, ip[1] = height, ip[2] = #obstacles;
                                   dp[0] = move penalty, dp[1] = collide penalty
                                   src/domains/collision-avoidance/CollisionAvoidanceBAExtension.cpp:59-89 */
};

enum fba_action_draw { FBA_ACT_UNIFORM_INT = 0, FBA_ACT_SLOW_INT = 1 };

/* the domain's sampleStartState (Environment::sampleStartState, src/environment/Environment.hpp:70) */
enum fba_start_kind {
    FBA_START_CONST = 0,       /* start_ip[0] */
    FBA_START_BOOL = 1,        /* rnd::boolean() ? start_ip[0] : start_ip[1] */
    FBA_START_UNIFORM_INT = 2, /* uniform_int over start_ip[0] states */
    FBA_START_SLOW2 = 3,       /* start_table[floor(u*ip[0]) * ip[1] + floor(u*ip[1])] */
    FBA_START_CATEGORICAL = 4  /* categorical over start_values[0..start_ip[0]) with total start_total */
};

/* domain `mutate` used by reinvigoration (FBAPOMDP::mutate, src/bayes-adaptive/models/factored/FBAPOMDP.cpp:57-61) */
enum fba_mutate_kind {
    FBA_MUT_FACTORED_TIGER = 0,
    FBA_MUT_COLLISION_AVOIDANCE = 1,
    FBA_MUT_SYSADMIN = 2,
    FBA_MUT_GRIDWORLD = 3
};

/* Stands in for what factory::makeTBAPOMDP / makeFBAPOMDP assemble
 * (src/bayes-adaptive/models/table/BAPOMDP.cpp:189-209, …/factored/FBAPOMDP.cpp:79-101):
 * Domain_Size, Domain_Feature_Size, the BADomainExtension and the domain's start distribution. */
typedef struct fba_model_desc {
    int32_t S, A, O;
    int32_t n_state_features, n_obs_features;
    int32_t state_feature_sizes[FBA_MAX_FEATURES];
    int32_t obs_feature_sizes[FBA_MAX_FEATURES];
    int32_t tabular; /* 1: BAFlatModel semantics (psi keyed and incremented by the NEW state);
                        0: BABNModel semantics (observation increment keyed by the OLD state) */
    int32_t domain;  /* fba_domain_kind */
    int32_t dom_ip[32];
    double dom_dp[8];
    const double* rew_sa;    /* FBA_DOM_TABLE only, host pointers, may be NULL */
    const double* rew_as2;
    const uint8_t* term_sa;
    const uint8_t* term_as2;
    int32_t action_draw; /* fba_action_draw */
    int32_t start_kind;  /* fba_start_kind */
    int32_t start_ip[4];
    const float* start_values; /* host pointers */
    double start_total;
    const int32_t* start_table;
    /* 0: every particle owns a dense count block. > 0 (tabular models only): base+delta storage —
     * particles share the prior's dense tables and own only a list of at most delta_capacity
     * increments (2 per update). For tabular models whose dense block is too large to replicate
     * (gridworld --size 5: 720 KB per particle); stands in for BAFlatModel's copy-on-write rows
     * (src/bayes-adaptive/states/table/BAFlatModel.cpp:185-252,284-354). */
    int32_t delta_capacity;
    /* 0: expected-Dirichlet mode, the reference default (--dirichlet_sampling_method expected,
     * BAConf.hpp:22; sampleFromExpectedMult / expectedMult). 1: sampled mode (regular): the
     * multinomial is drawn from the Dirichlet for every step and every likelihood
     * (sampleFromSampledMult / sampleMult, src/utils/random.cpp:217-242,281-304). PHILOX mode only;
     * statistical parity (device libm differs from glibc in the last bits of log / pow). */
    int32_t dirichlet_sampling;
} fba_model_desc;

enum fba_rng_mode { FBA_RNG_REPLAY = 0, FBA_RNG_PHILOX = 1 };

/* The random source of one call. REPLAY: `words` (host memory) from `cursor` on; the call advances
 * `cursor` by exactly what the reference would have drawn. PHILOX: (seed, offset); the call
 * advances `offset`. */
typedef struct fba_rng {
    int32_t mode;
    const uint32_t* words;
    int64_t n_words;
    int64_t cursor;
    uint64_t seed;
    uint64_t offset;
} fba_rng;

/* ---- context ---- */
int fba_abi_version(void); /* FBA_ABI_VERSION the library was built with */
int fba_ctx_create(int device, fba_ctx** out);
void fba_ctx_destroy(fba_ctx* ctx);
const char* fba_last_error(const fba_ctx* ctx);
/* the cudaStream_t all work of this context is launched on (for event timing by the host) */
void* fba_ctx_stream(const fba_ctx* ctx);
int fba_ctx_synchronize(fba_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t fba_ctx_launch_count(const fba_ctx* ctx);
/* device-side counters since creation. which = 0: simulated steps executed by the rollout kernels
 * (RBAPOUCT::rollout ends early at terminal states, RBAPOUCT.cpp:306) */
int64_t fba_ctx_counter(fba_ctx* ctx, int32_t which);

/* options: "inplace_resample" (default 1): PHILOX-mode resampling keeps surviving particles in
 * their slot and copies only duplicates; 0 = gather every particle into the second buffer */
/*          "rollout_coop" (default -1 = thread per rollout): 1 = one warp per rollout with
 *          cooperatively loaded rows, 0 = one thread per rollout
 *          "fused_update" (default 1): fba_belief_update_estimation on a weighted PHILOX belief of at
 *          most 2048 particles runs update + resample in ONE launch (bit-identical to the
 *          launch-per-phase path); 0 = always launch per phase
 *          "bulk_copy" (default 0): 1 = full-copy gathers go through the TMA engine (cp.async.bulk)
 *          "auto_compact" (default 1): a belief in base + journal / base + delta storage whose increment lists are
 *          full continues in dense storage (fba_belief_compact) instead of returning FBA_ERR_CAPACITY
 *          "msg_cluster" (default 1): fba_belief_sample_state_history by messages spreads one model over a
 *          thread-block cluster of 4 CTAs (distributed shared memory) when its states split evenly; 0 = one CTA per model
 *          "nested_exact" (default 0): 1 = PHILOX-mode fba_nested_update runs the reference's loop attempt by
 *          attempt, one thread per top particle (what REPLAY mode always does); 0 = one warp per top particle,
 *          32 attempts per round that all see the counts as of the start of their round */
int fba_ctx_set_option(fba_ctx* ctx, const char* name, int64_t value);

/* per-kernel CUDA-event timing on the context's stream: begin, run calls, end, then query the
 * summed milliseconds and launch count of the kernels whose name starts with `prefix` */
int fba_ctx_profile_begin(fba_ctx* ctx);
int fba_ctx_profile_end(fba_ctx* ctx);
int fba_ctx_profile_get(const fba_ctx* ctx, const char* prefix, double* total_ms, int64_t* count);
/* "name:ms:count;..." of every kernel timed since fba_ctx_profile_begin; returns the bytes needed */
int64_t fba_ctx_profile_list(const fba_ctx* ctx, char* buf, int64_t cap);

/* ---- model ---- */
int fba_model_create(fba_ctx* ctx, const fba_model_desc* desc, int32_t max_structures, fba_model** out);
void fba_model_destroy(fba_model* m);
/* Registers n structures (t_par: [n][A*FS], o_par: [n][A*FO] parent bitmasks); duplicates map to the
 * existing id. Stands in for BABNModel::structure() (src/bayes-adaptive/states/factored/BABNModel.cpp:263-290). */
int fba_model_add_structures(fba_model* m, int32_t n, const uint32_t* t_par, const uint32_t* o_par,
                             int32_t* ids_out);
int32_t fba_model_num_structures(const fba_model* m);
int64_t fba_model_structure_size(const fba_model* m, int32_t id); /* float cells */
int fba_model_get_structure(const fba_model* m, int32_t id, uint32_t* t_par, uint32_t* o_par);

/* ---- belief ---- */
/* stride_floats = 0: the largest registered structure, rounded up to 4 floats. weighted: 1 for
 * WeightedFilter semantics (importance sampling), 0 for FlatFilter (rejection sampling). */
int fba_belief_create(fba_ctx* ctx, fba_model* m, int64_t n_particles, int64_t stride_floats,
                      int32_t weighted, fba_belief** out);
void fba_belief_destroy(fba_belief* b);
int64_t fba_belief_size(const fba_belief* b);
int64_t fba_belief_stride(const fba_belief* b);

/* Belief::initiate (src/beliefs/Belief.hpp:24) with the particles the host-side prior produced:
 * n_protos distinct (structure id, count block) prototypes and, per particle, which prototype it
 * clones and its domain start state. particle_proto may be NULL (all prototype 0). Weights 1/N. */
int fba_belief_init(fba_belief* b, int32_t n_protos, const int32_t* proto_struct_id,
                    const float* proto_counts /* [n_protos][stride] */,
                    const int32_t* particle_proto, const int32_t* particle_state);
/* same, but domain start states (and, if proto_probs != NULL, prototypes) drawn on device */
int fba_belief_init_sampled(fba_belief* b, int32_t n_protos, const int32_t* proto_struct_id,
                            const float* proto_counts, const double* proto_probs, fba_rng* rng);
/* raw particle I/O (debugging, oracle diffing, Belief::sample() materialisation); NULLs skipped */
int fba_belief_upload(fba_belief* b, int64_t first, int64_t count, const int32_t* state,
                      const int32_t* struct_id, const float* counts, const double* w);
int fba_belief_download(fba_belief* b, int64_t first, int64_t count, int32_t* state,
                        int32_t* struct_id, float* counts, double* w);
int fba_belief_total_weight(fba_belief* b, double* total); /* WeightedFilter::_total_weight */
/* base + delta / base + journal storage -> dense private blocks, in place: every particle's block becomes its base
 * table with its own increments applied (bit-identical to what fba_belief_download reports), the belief continues as
 * a dense one of any age. What BAFlatModel::operator= does when a tenth of a table's rows are private
 * (src/bayes-adaptive/states/table/BAFlatModel.cpp:284-354). Called automatically by the importance-sampling update
 * when the increment lists are full (option "auto_compact", default 1) if dense storage fits the free memory;
 * FBA_ERR_CAPACITY if it does not. No-op on a dense belief. */
int fba_belief_compact(fba_belief* b);
/* increments a particle of this belief can still hold in total (0: dense storage) */
int32_t fba_belief_delta_capacity(const fba_belief* b);

/* beliefs::importance_sampling::update (src/beliefs/particle_filters/ImportanceSampler.hpp:31-62):
 * step every particle with `action`, weight by P(observation | action, particle), normalise.
 * *likelihood = the un-normalised weight total (the function's return value). */
int fba_belief_update(fba_belief* b, int32_t action, int32_t observation, fba_rng* rng,
                      double* likelihood);
/* beliefs::importance_sampling::resample (ImportanceSampler.hpp:71-94). PHILOX mode resamples
 * systematically (one uniform, sorted ancestors); REPLAY draws the reference's N multinomial picks. */
int fba_belief_resample(fba_belief* b, fba_rng* rng);
/* BAImportanceSampling::updateEstimation (src/beliefs/bayes-adaptive/BAImportanceSampling.cpp:74-88) */
int fba_belief_update_estimation(fba_belief* b, int32_t action, int32_t observation, fba_rng* rng,
                                 double* likelihood);
/* BABelief::resetDomainStateDistribution (src/beliefs/bayes-adaptive/BABelief.hpp:30):
 * weighted: BAImportanceSampling.cpp:90-111; flat: BARejectionSampling.cpp:47-58 */
int fba_belief_reset_domain_states(fba_belief* b, fba_rng* rng);
/* BAPOMDP::resetDomainState (src/bayes-adaptive/models/table/BAPOMDP.cpp:69-77) on every particle where it
 * is: a fresh domain start state each, counts, weights and order untouched — the reset of the beliefs
 * that keep their weighted filter as it is (MHNIPS2018.cpp:132-147, CheatingReinvigoration.cpp:50-66,
 * StructureIncubatorSampling.cpp:46-61). For a flat belief the same as fba_belief_reset_domain_states. */
int fba_belief_redraw_domain_states(fba_belief* b, fba_rng* rng);
/* Belief::sample (Belief.hpp:34): index of the drawn particle (use fba_belief_download to view it). PHILOX mode,
 * right after init / a resample / a reset — every weight is 1 / N, which the library tracks — the weighted draw is
 * floor(u N) computed on the host: no kernel, no read-back (fba_runs_sample uses the same rule and the same u).
 * REPLAY mode always evaluates WeightedFilter::sample's arithmetic (WeightedFilter.cpp:163-191). */
int fba_belief_sample(fba_belief* b, fba_rng* rng, int64_t* index);

/* beliefs::rejectSample (src/beliefs/particle_filters/RejectionSampling.hpp:26-72) on a flat belief */
int fba_belief_reject_sample(fba_belief* b, int32_t action, int32_t observation, fba_rng* rng,
                             int64_t* attempts);
/* ReinvigoratingRejectionSampling::reinvigorateParticles
 * (src/beliefs/bayes-adaptive/factored/ReinvigoratingRejectionSampling.cpp:121-131) */
int fba_belief_reinvigorate(fba_belief* b, fba_belief* fully_connected, int64_t amount,
                            int32_t mutate_kind, fba_rng* rng);

/* RBAPOUCT::rollout (src/planners/bayes-adaptive/RBAPOUCT.cpp:295-323), n of them at once:
 * rollout i starts from particle[i]'s counts (read-only, KeepCounts) in domain state
 * start_state[i] and runs depth[i] steps or to a terminal. REPLAY: rollout i reads words from
 * rng->cursor + word_offset[i]; PHILOX: word_offset ignored. returns: n doubles (host). */
int fba_rollouts(fba_belief* b, int64_t n, const int64_t* particle, const int32_t* start_state,
                 const int32_t* depth, double discount, fba_rng* rng, const int64_t* word_offset,
                 double* returns);

/* ---- planner support (wave-parallel POMCP on the host, its simulator calls batched on the GPU) ---- */
/* n x Belief::sample (Belief.hpp:34) in one launch: the root particles of n simulations
 * (RBAPOUCT.cpp:92). PHILOX mode. indices: n host int64. */
int fba_belief_sample_batch(fba_belief* b, fba_rng* rng, int64_t n, int64_t* indices);
/* the domain states of n particles (BAState::_domain_state, src/bayes-adaptive/states/BAState.hpp:20) */
int fba_belief_gather_states(fba_belief* b, int64_t n, const int64_t* indices, int32_t* states);
/* n x BAPOMDP::step in KeepCounts mode (BAPOMDP.cpp:111-143) from (particle, domain state, action):
 * the in-tree steps of a wave of simulations (RBAPOUCT::traverseChanceNode, RBAPOUCT.cpp:249).
 * Outputs (host, n each): new domain state, observation, reward, terminal flag. PHILOX mode. */
int fba_step_batch(fba_belief* b, int64_t n, const int64_t* particle, const int32_t* state,
                   const int32_t* action, fba_rng* rng, int32_t* new_state, int32_t* observation,
                   double* reward, int32_t* terminal);

/* ---- multi-GPU phases (one process per GPU; the host runs the collective between them) ---- */
/* phase 1: step + weight, no normalisation; *local_total = this shard's weight sum. With
 * local_total == NULL the call only enqueues work: the total is then the first double at
 * fba_belief_scalars_ptr(), valid in stream order (for a device-side all-gather). */
int fba_belief_propose(fba_belief* b, int32_t action, int32_t observation, fba_rng* rng,
                       double* local_total);
/* phase 2: divide by the global total (all-gathered by the host) */
int fba_belief_normalize(fba_belief* b, double global_total);
/* phase 3: resample this shard to n_offspring particles, the first min(n_offspring, N) stay here,
 * the surplus lands in an export buffer (fba_belief_export_ptr) for the host to ship over NCCL */
int fba_belief_resample_shard(fba_belief* b, int64_t n_offspring, fba_rng* rng);
/* REPLAY mode, beliefs of at least `parallel_chains_min` particles (fba_ctx_set_option, default 8192): the
 * reference's three sequential double chains (ImportanceSampler.hpp:51, WeightedFilter.cpp:130-143,
 * 168-183) are evaluated in parallel segments with bit-identical results; this returns how many segments
 * (of 1024 weights) had to be recomputed sequentially since creation, -1 if that path never ran */
int64_t fba_belief_chain_recomputed(fba_belief* b);
/* in-place resampler statistics since creation: count blocks copied, resamples run */
int fba_belief_resample_stats(fba_belief* b, int64_t* copies, int64_t* resamples);
/* phases 2+3 with the quota allocation (systematic over ranks, shared offset u in [0,1)) and the
 * exchange plan computed inside: totals = the all-gathered shard totals; send_plan (n_ranks x
 * n_ranks, row-major, may be NULL) receives the records rank g ships to rank h */
int fba_belief_shard_resample(fba_belief* b, const double* totals, int32_t n_ranks, int32_t rank, double u,
                              fba_rng* rng, int64_t* send_plan, double* global_total);
/* the same, split so that nothing waits for the GPU before the resampling kernels are enqueued:
 * _async reads the shard totals from DEVICE memory (where the all-gather left them) and computes
 * this rank's quota on device; _plan (host, any time later) returns the exchange plan for the same
 * totals and sets this rank's export / import bookkeeping. The export buffer must have been
 * reserved (fba_belief_reserve_export); a surplus beyond it makes _plan return FBA_ERR_CAPACITY. */
int fba_belief_shard_resample_async(fba_belief* b, const double* totals_device, int32_t n_ranks,
                                    int32_t rank, double u, fba_rng* rng);
int fba_belief_shard_plan(fba_belief* b, const double* totals_host, int32_t n_ranks, int32_t rank, double u,
                          int64_t* send_plan, double* global_total);
/* Peer-to-peer sharded update over NVLink / NVSwitch (one process per GPU, one box). The reference has
 * no distributed path; this is the multi-GPU form of BAImportanceSampling::updateEstimation
 * (src/beliefs/bayes-adaptive/BAImportanceSampling.cpp:74-88) named by SURVEY.md §8e. Setup, once:
 * every rank calls _p2p_export (allocates its control block, returns a blob of CUDA IPC handles), the
 * host all-gathers the blobs (the only collective, at setup), every rank calls _p2p_open. Per update
 * ONE call, fba_belief_sharded_update, enqueues everything on the context's stream with no host
 * synchronisation and no NCCL: the shard totals, "dead-slot list complete" and "records landed" travel
 * between GPUs as step-stamped flags in peer-mapped memory, and the surplus blocks of an over-quota
 * shard are stored by the resampling kernel straight into dead slots of the destination GPU's particle
 * array — no staging buffer, so no capacity limit however skewed the shard weights are.
 * u in [0,1): the shared systematic offset of the quota allocation, the same on every rank.
 * likelihood (may be NULL = fully asynchronous): the GLOBAL un-normalised weight total.
 * Every rank must make the same sequence of sharded_update calls. A cross-rank wait longer than the
 * timeout (default 20 s) is abandoned and counted (fba_belief_p2p_timeouts; the belief is then
 * invalid) instead of hanging the GPU. */
#define FBA_P2P_BLOB_BYTES 320
int fba_belief_p2p_export(fba_belief* b, void* blob);
int fba_belief_p2p_open(fba_belief* b, const void* blobs, int32_t n_ranks, int32_t rank);
int fba_belief_sharded_update(fba_belief* b, int32_t action, int32_t observation, fba_rng* rng, double u,
                              double* likelihood);
int64_t fba_belief_p2p_timeouts(fba_belief* b);
int fba_belief_p2p_set_timeout(fba_belief* b, double seconds);
/* surplus records that did not fit an export / import buffer since creation (0 in a healthy run) */
int64_t fba_belief_dropped_records(fba_belief* b);
int fba_belief_reserve_export(fba_belief* b, int64_t records);
/* phase 4 from an arbitrary device address (e.g. a segment of an all-gathered export window);
 * successive calls fill successive dead slots. Asynchronous. */
int fba_belief_import_from(fba_belief* b, const void* records_device, int64_t n_records);
int64_t fba_belief_export_count(const fba_belief* b);
/* device pointers of the export / import staging area: particle records of
 * fba_belief_record_bytes() each (count block, then state, structure id) */
void* fba_belief_export_ptr(fba_belief* b);
void* fba_belief_import_ptr(fba_belief* b, int64_t n_records);
int64_t fba_belief_record_bytes(const fba_belief* b);
/* phase 4: place n_records imported records into the slots the local resample left empty */
int fba_belief_import(fba_belief* b, int64_t n_records);

/* BABNModel::LogBDScore (src/bayes-adaptive/states/factored/BABNModel.cpp:451-478, DBNNode.cpp:82-117):
 * scores[i] (host, N doubles) = log Bayesian-Dirichlet score of particle i of `b` against the prior
 * counts in `prior` — particle i of `prior`, or its only particle if it has one. Both beliefs share a
 * model and the compared particles a structure (FBA_ERR_INVALID otherwise). Factored models, dense
 * storage. First brick of the reference's MCMC structure beliefs. */
int fba_belief_log_bd_score(fba_belief* b, fba_belief* prior, double* scores);

/* ---- MH structure beliefs ----------------------------------------------------
 * computePosterior of the reference's Metropolis-Hastings structure beliefs
 * (src/beliefs/bayes-adaptive/factored/MHNIPS2018.cpp:41-109): the whole (action, observation) history —
 * n_episodes episodes, episode e of episode_len[e] steps, actions / observations concatenated (host) —
 * replayed on EVERY particle of `b` (the proposals: each holds the prior model of its proposed
 * structure), one thread per particle: per episode attempt a domain start state, per step s' and o from
 * the particle's counts (expected Dirichlets), +1 on (s, a, o, s') if o is the observed one, otherwise
 * the episode's increments are taken back (-1) and the episode is tried again. Each particle's domain
 * state becomes the state after the last step. More than max_attempts attempts in one particle:
 * FBA_ERR_CAPACITY. REPLAY mode: particle i draws from the i-th equal slice of the remaining words (a belief of ONE
 * particle consumes exactly the words it drew, like the reference). */
int fba_belief_replay_history(fba_belief* b, int32_t n_episodes, const int32_t* episode_len,
                              const int32_t* actions, const int32_t* observations, fba_rng* rng,
                              int64_t max_attempts);
/* particles src[src_index[j]] -> dst[first + j], j < n (count block, domain state, structure id); both
 * beliefs share context, model and stride. How accepted proposals become the new belief
 * (MHNIPS2018.cpp:241-246). Weights are untouched. */
int fba_belief_assign_from(fba_belief* dst, int64_t first, fba_belief* src, int64_t n, const int64_t* src_index);

/* ---- MHwithinGibbs -------------------------------------------------------------
 * sampleStateHistory of beliefs::bayes_adaptive::factored::MHwithinGibbs
 * (src/beliefs/bayes-adaptive/factored/MHwithinGibbs.cpp:38-232): for EVERY particle of `b` (each one a model)
 * a sequence of domain states for the whole (action, observation) history — episode_len[e] + 1 states per
 * episode, so states is [N][n_steps + n_episodes] (host) — conditioned on the observations:
 *   method 0 (MSG, :96-213): BABNModel::flattenT / flattenO as float tables, per episode a backward pass of
 *     normalised double messages (one GPU thread per domain state, each running the reference's sequential
 *     inner product; sequential totals) and forward sampling; state_prior = S floats,
 *     FBAPOMDP::domainStatePrior()->prob(s) (the prior of s_0); one CTA per particle;
 *   method 1 (RS, :38-94): rejection sampling with the particle's counts, one thread per particle; more than
 *     max_attempts episode attempts in one particle: FBA_ERR_CAPACITY.
 * Counts are not modified. REPLAY: particle i draws from the i-th equal slice of the remaining words and the call
 * takes all of them; a belief of ONE particle consumes exactly the words it drew, like the reference. */
int fba_belief_sample_state_history(fba_belief* b, int32_t method, int32_t n_episodes, const int32_t* episode_len,
                                    const int32_t* actions, const int32_t* observations, const float* state_prior,
                                    fba_rng* rng, int64_t max_attempts, int32_t* states);
/* MHwithinGibbs::computePosteriorCounts (MHwithinGibbs.cpp:397-436): every particle of `b` (holding the prior
 * counts of its structure) gets incrementCountsOf(s_t, a_t, o_t, s_t+1) for every step of its state history:
 * states [N][n_steps + n_episodes], or with shared = 1 ONE history [n_steps + n_episodes] for all particles
 * (the proposals of one Gibbs sweep share the current state sequence). */
int fba_belief_add_history_counts(fba_belief* b, int32_t n_episodes, const int32_t* episode_len,
                                  const int32_t* actions, const int32_t* observations, const int32_t* states,
                                  int32_t shared);

/* ---- single particles between the filters of the composite structure beliefs ----
 * src[src_index[j]] -> dst[dst_index[j]] for j = 0 .. n-1 IN ORDER (a later j overwrites an earlier one on
 * the same slot): count block, domain state, structure id. A weighted dst follows WeightedFilter::replace
 * (src/beliefs/particle_filters/WeightedFilter.cpp:71-90): the replaced slot's weight becomes
 * _total_weight / N and _total_weight moves by the difference, slot after slot. Two different beliefs of
 * one context, model and stride, dense storage. */
int fba_belief_replace_from(fba_belief* dst, const int64_t* dst_index, fba_belief* src, const int64_t* src_index,
                            int64_t n);
/* CheatingReinvigoration::cheat (src/beliefs/bayes-adaptive/prototypes/CheatingReinvigoration.cpp:136-147):
 * `amount` times, a uniformly drawn particle of the flat correct-structure filter replaces a particle of the
 * weighted belief at slot slowRandomInt(0, N) (draw order: the source first, as g++ evaluates the call). */
int fba_belief_cheat(fba_belief* belief, fba_belief* correct, int64_t amount, fba_rng* rng);
/* beliefs::bayes_adaptive::factored::breed (factored/ReinvigoratingRejectionSampling.cpp:24-35) n times, in
 * order: counts donor drawn from `fully_connected`, structure donor from `structure_donors` (both flat),
 * the domain's mutate, BABNModel::marginalizeOut; the j-th bred particle replaces particle dst_slot[j] of
 * `dst` (weights as fba_belief_replace_from). What StructureIncubatorSampling does to its shadow belief
 * (factored/StructureIncubatorSampling.cpp:74-80,139-153). */
int fba_belief_breed_into(fba_belief* dst, const int64_t* dst_slot, int64_t n, fba_belief* structure_donors,
                          fba_belief* fully_connected, int32_t mutate_kind, fba_rng* rng);
/* WeightedFilter::leastLikely(n) (WeightedFilter.cpp:206-243): index[0..n) in the reference's order (the
 * same std::priority_queue, driven the same way, so ties resolve identically). n < N. */
int fba_belief_least_likely(fba_belief* b, int64_t n, int64_t* index);
/* StructureIncubatorSampling::reinvigorateBelief (factored/StructureIncubatorSampling.cpp:155-187): shadow
 * particles whose normalised weight exceeds `threshold` are copied over uniformly drawn particles of the
 * flat `belief` (one draw each, particle order), their weights zeroed and the shadow weights normalised. */
int fba_belief_promote(fba_belief* shadow, fba_belief* belief, double threshold, fba_rng* rng, int64_t* n_promoted);

/* ---- NestedBelief ---------------------------------------------------------------
 * beliefs::bayes_adaptive::NestedBelief (src/beliefs/bayes-adaptive/NestedBelief.{hpp,cpp}): P(counts) x
 * P(state | counts) as a weighted TOP filter of n_top count blocks, each with its own flat BOTTOM filter of
 * n_bottom domain states (the factory uses n and n^2, BABelief.cpp:66-69). Tabular and factored models,
 * dense storage. */
typedef struct fba_nested fba_nested;
int fba_nested_create(fba_ctx* ctx, fba_model* m, int64_t n_top, int64_t n_bottom, int64_t stride, fba_nested** out);
void fba_nested_destroy(fba_nested* n);
/* the top filter — a weighted fba_belief for init / upload / download of count blocks, structure ids and
 * weights (its per-particle domain state is unused, as in the reference, NestedBelief.cpp:84-88) */
fba_belief* fba_nested_top(fba_nested* n);
int64_t fba_nested_bottom_size(const fba_nested* n);
/* the bottom filters of top particles [first_top, first_top + count): count x n_bottom domain states */
int fba_nested_upload_states(fba_nested* n, int64_t first_top, int64_t count, const int32_t* states);
int fba_nested_download_states(fba_nested* n, int64_t first_top, int64_t count, int32_t* states);
/* NestedBelief::resetDomainStateDistribution (NestedBelief.cpp:33-61): fresh start states in every bottom
 * filter, drawn on the device from the model's start-state sampler */
int fba_nested_reset_domain_states(fba_nested* n, fba_rng* rng);
/* NestedBelief::updateEstimation (NestedBelief.cpp:129-193): per top particle, rejection sampling of its
 * bottom filter (KeepCounts steps on the particle's own counts) until n_bottom states are accepted, every
 * acceptance adding 1 / n_bottom to the counts it went through BEFORE the next attempt; weight *= 1 /
 * attempts; then WeightedFilter::normalize. One GPU thread per top particle runs that loop as written.
 * PHILOX mode by default gives each top particle a warp: 32 attempts per round from 32 streams, the round's
 * acceptances landing together (option "nested_exact" = 1 keeps the attempt-by-attempt loop).
 * attempts (n_top, may be NULL) receives the attempts per top particle. A particle that needs more than
 * max_attempts: FBA_ERR_CAPACITY. REPLAY: top particle i draws from the i-th equal slice of the remaining
 * words (the reference's single stream is data dependent across particles). */
int fba_nested_update(fba_nested* n, int32_t action, int32_t observation, fba_rng* rng, int64_t max_attempts,
                      int64_t* attempts);
/* NestedBelief::sample (NestedBelief.cpp:117-127): a weighted top draw, then a uniform bottom draw */
int fba_nested_sample(fba_nested* n, fba_rng* rng, int64_t* top_index, int32_t* state);

/* ---- POMCP with the search tree on the device --------------------------------
 * planners::RBAPOUCT::selectAction (src/planners/bayes-adaptive/RBAPOUCT.cpp:67-153) as waves of
 * `wave` concurrent simulations, each one entirely on the device: root particle from the belief
 * (weighted or flat), UCB descent with BAPOMDP::step in KeepCounts mode (RBAPOUCT.cpp:197-277), one
 * new leaf, random-policy rollout (RBAPOUCT.cpp:295-323), back-up. Simulations of one wave see each
 * other through atomic statistics (selection counts are raised when an action is chosen). PHILOX
 * mode. max_simulations bounds the tree (one node per simulation), max_depth the search depth. */
typedef struct fba_tree fba_tree;
int fba_tree_create(fba_ctx* ctx, fba_model* model, int64_t max_simulations, int32_t max_depth, fba_tree** out);
void fba_tree_destroy(fba_tree* tree);
/* depth = min(horizon - history length, max depth) as RBAPOUCT.cpp:80; u = UCB exploration constant;
 * *action = argmax_a Q(root, a), ties broken uniformly; q / visits (n_actions entries, may be NULL)
 * receive the root's mean returns and completed visit counts */
int fba_tree_search(fba_tree* tree, fba_belief* b, int64_t n_simulations, int32_t depth, double u,
                    double discount, int32_t wave, fba_rng* rng, int32_t* action, double* q, int64_t* visits);

/* ---- many independent runs on one GPU ---------------------------------------
 * The reference runs its `--runs` one after the other (src/experiments/BAPOMDPExperiment.cpp:32-78),
 * each with its own small belief (episodic tiger: 1024 particles), which on its own leaves a GPU
 * idle. An fba_runs object holds the importance-sampling beliefs of n_runs runs of
 * particles_per_run particles back to back (run r owns particles [r n, (r+1) n) of
 * fba_runs_belief()), and each call below advances EVERY run with one kernel launch, one CTA per
 * run, each run with its own action / observation. PHILOX mode, dense storage. Run r is
 * bit-identical to a stand-alone weighted fba_belief of particles_per_run particles driven by an
 * fba_rng with seed + r through the same sequence of calls (each call here advances rng->offset
 * exactly as its single-belief counterpart does).
 * active (n_runs bytes, may be NULL = all): runs with 0 are left untouched (e.g. finished episodes). */
typedef struct fba_runs fba_runs;
int fba_runs_create(fba_ctx* ctx, fba_model* model, int32_t n_runs, int64_t particles_per_run, int64_t stride,
                    fba_runs** out);
void fba_runs_destroy(fba_runs* runs);
/* the underlying particle storage, for fba_belief_download / fba_belief_upload (NOT a belief to
 * update on its own: its weights are normalised per run) */
fba_belief* fba_runs_belief(fba_runs* runs);
/* Belief::initiate of every run: as fba_belief_init_sampled */
int fba_runs_init_sampled(fba_runs* runs, int32_t n_protos, const int32_t* proto_struct_id,
                          const float* proto_counts, const double* proto_probs, fba_rng* rng);
/* the same from explicit host particles, as fba_belief_init: particle_proto / particle_state have
 * n_runs * particles_per_run entries (run-major) */
int fba_runs_init(fba_runs* runs, int32_t n_protos, const int32_t* proto_struct_id, const float* proto_counts,
                  const int32_t* particle_proto, const int32_t* particle_state);
/* Belief::updateEstimation of every run (BAImportanceSampling.cpp:74-88): action / observation are
 * n_runs entries on the host; likelihood (n_runs doubles, may be NULL) receives each run's step
 * likelihood */
int fba_runs_update_estimation(fba_runs* runs, const int32_t* action, const int32_t* observation,
                               const uint8_t* active, fba_rng* rng, double* likelihood);
/* BABelief::resetDomainStateDistribution of every (active) run */
int fba_runs_reset_domain_states(fba_runs* runs, const uint8_t* active, fba_rng* rng);
/* Belief::sample of every (active) run: index[r] = particle index inside run r (storage index
 * r * particles_per_run + index[r]); entries of inactive runs are left untouched */
int fba_runs_sample(fba_runs* runs, const uint8_t* active, fba_rng* rng, int64_t* index);
/* Planner::selectAction of every (active) run at once (RBAPOUCT.cpp:67-153, as fba_tree_search): each
 * run has its own search tree on the device; a wave advances sims_per_wave simulations of EVERY run,
 * so with sims_per_wave = 1 each run's search is the reference's sequential algorithm while the GPU
 * is kept busy by the number of runs. depth: n_runs entries (episodes may be at different steps).
 * action: n_runs entries; q / visits (n_runs x n_actions, may be NULL): the roots' mean returns and
 * visit counts. With sims_per_wave = 1, run r is bit-identical to fba_tree_search(wave = 1) on a
 * stand-alone belief with an fba_rng seeded seed + r. Call after init / update_estimation / reset
 * (uniform weights). */
int fba_runs_plan(fba_runs* runs, int64_t n_simulations, const int32_t* depth, double u, double discount,
                  int32_t sims_per_wave, const uint8_t* active, fba_rng* rng, int32_t* action, double* q,
                  int64_t* visits);
/* count blocks copied by the in-place resamples of all runs since creation */
int64_t fba_runs_copies(fba_runs* runs);

/* device pointers of the current particle arrays, for zero-copy views by the host language (once the weight pointer
 * has been handed out the library no longer assumes it knows the weights: the uniform-weights shortcuts are off) */
void* fba_belief_counts_ptr(fba_belief* b);
void* fba_belief_state_ptr(fba_belief* b);
void* fba_belief_weight_ptr(fba_belief* b);
void* fba_belief_aux_ptr(fba_belief* b); /* device double[N]: REPLAY: the remainders R_k of WeightedFilter::sample
                                           (valid after an update / a pick); PHILOX: the cdf */
void* fba_belief_scalars_ptr(fba_belief* b); /* device double[4]: [0] = last un-normalised total */

#ifdef __cplusplus
}
#endif
#endif /* FBA_POMDP_B200_H */
