/* TEST INFRASTRUCTURE — the CPU oracle. Not shipped, never on the product path.
 *
 * Plain-C restatement of the reference's (samkatt/fba-pomdp) particle-belief / rollout hot path:
 * the algorithm only, over flat arrays, with the random draws taken from an explicit stream of
 * 32-bit words (what the reference's global std::mt19937 produced, src/utils/random.cpp:11).
 * Every function cites the reference file:line it follows.
 *
 * Parity status: PINNED. tests/test_oracle_vs_golden.py checks this file against fixtures under
 * tests/golden/ that oracle/gen_golden.py produced by running the unmodified reference
 * (oracle/_ref/libfba_ref.so, built by oracle/Makefile from /root/reference) under seed "42".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#ifndef FBA_ORACLE_H
#define FBA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXF 16 /* max state / observation features */

/* domain functors: BADomainExtension::{reward,terminal} + generateRandomAction */
enum {
    ORC_DOM_TABLE = 0,          /* reward = rew_sa[s*A+a] + rew_as2[a*S+s'], terminal likewise (OR) */
    ORC_DOM_TIGER = 1,          /* ip[0] = episodic                      TigerBAExtension.cpp:21-44 */
    ORC_DOM_FACTORED_TIGER = 2, /* ip[0] = episodic              FactoredTigerBAExtension.cpp:27-56 */
    ORC_DOM_SYSADMIN = 3,       /* ip[0] = #computers, dp[0] = reboot cost SysAdminBAExtension.cpp:27-48 */
    ORC_DOM_GRIDWORLD = 4,      /* ip[0] = size, ip[1] = #goals, ip[2+2g],ip[3+2g] = goal g (x,y);
                                   dp[0] = goal reward, dp[1] = step reward GridWorldBAExtension.cpp:74-100 */
    ORC_DOM_COLLISION_AVOIDANCE = 5 /* ip[0] = width, ip[1] = height, ip[2] = #obstacles;
                                   dp[0] = move penalty, dp[1] = collide penalty
                                   CollisionAvoidanceBAExtension.cpp:59-89 */
};

/* how generateRandomAction draws */
enum { ORC_ACT_UNIFORM_INT = 0 /* 1 word + Lemire */, ORC_ACT_SLOW_INT = 1 /* floor(u*A) */ };

/* how the domain's sampleStartState draws (SURVEY.md §8 a') */
enum {
    ORC_START_CONST = 0,       /* ip[0]                                   SysAdmin.cpp:102-105 */
    ORC_START_BOOL = 1,        /* boolean() ? ip[0] : ip[1]               Tiger.cpp:16-19 */
    ORC_START_UNIFORM_INT = 2, /* uniform_int over ip[0] states           FactoredTiger.cpp:71-75 */
    ORC_START_SLOW2 = 3,       /* table[floor(u*ip[0]) * ip[1] + floor(u*ip[1])] GridWorld.cpp:260-270 */
    ORC_START_CATEGORICAL = 4  /* sampleFromMult(values, ip[0], total)    distributions.cpp:47-51 */
};

typedef struct {
    int32_t S, A, O, FS, FO;
    int32_t feat_s[ORC_MAXF], feat_o[ORC_MAXF];
    int32_t tabular; /* 1: BAFlatModel (psi keyed by NEW state); 0: BABNModel (increment quirk) */
    int32_t domain;
    int32_t dom_ip[32];
    double dom_dp[8];
    const double* rew_sa;
    const double* rew_as2;
    const uint8_t* term_sa;
    const uint8_t* term_as2;
    int32_t action_draw;
    int32_t start_kind;
    int32_t start_ip[4];
    const float* start_values;
    double start_total;
    const int32_t* start_table;
} orc_model;

/* word stream + the three libstdc++ (GCC 13) distributions the reference uses on it */
typedef struct {
    const uint32_t* words;
    int64_t n;
    int64_t cur;
    int32_t overrun; /* set when a draw ran past n */
} orc_rng;

double orc_uniform01(orc_rng* g);            /* uniform_real_distribution<double>(0,1): 2 words */
int orc_boolean(orc_rng* g);                 /* bernoulli_distribution(0.5): 2 words */
int32_t orc_uniform_int(orc_rng* g, uint32_t range); /* uniform_int_distribution<int>(0,range-1) */

/* A structure table: n_structs structures, each A*FS transition-node parent masks followed by
 * A*FO observation-node parent masks (bit f set = state feature f is a parent). */
typedef struct {
    int32_t n_structs;
    int32_t cap;
    uint32_t* t_par; /* [cap][A*FS] */
    uint32_t* o_par; /* [cap][A*FO] */
} orc_structs;

/* a belief: N particles, count block i at counts + i*stride (layout: for a: T nodes f, O nodes g;
 * each CPT row-major [parent configuration][output]) */
typedef struct {
    int64_t N;
    int64_t stride;
    float* counts;
    int32_t* state;
    int32_t* struct_id;
    double* w;          /* NULL for flat (unweighted) filters */
    double total_weight; /* WeightedFilter::_total_weight */
} orc_belief;

int64_t orc_struct_size(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par);
/* node offsets: off[a*(FS+FO) + j], j < FS transition node, else observation node; returns size */
int64_t orc_struct_offsets(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, int64_t* off);

double orc_reward(const orc_model* m, int s, int a, int s2, int* terminal);
int orc_sample_start_state(const orc_model* m, orc_rng* g);

/* BAPOMDP::step (BAPOMDP.cpp:111-143). Mutates *state (and counts iff update_counts). */
double orc_step(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                int32_t* state, int a, int update_counts, orc_rng* g, int* o_out, int* terminal);
/* BA{Flat,BN}Model::computeObservationProbability in expected mode */
double orc_obs_prob(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par,
                    const float* counts, int state, int a, int o);

/* BABNModel::LogBDScore(prior) (BABNModel.cpp:451-478, DBNNode.cpp:82-117, logGamma random.cpp:127-135):
 * log Bayesian-Dirichlet score of a particle's counts against prior counts of the SAME structure — the
 * quantity the reference's MCMC structure beliefs (MHNIPS2018.cpp:237-238, MHwithinGibbs.cpp:352,365)
 * compare. Same accumulation order as the reference (per row, per node, nodes in block order). */
double orc_log_bd_score(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                        const float* prior_counts);

/* importance_sampling::update (ImportanceSampler.hpp:31-62); returns the un-normalised total */
double orc_is_update(const orc_model* m, const orc_structs* st, orc_belief* b, int a, int o, orc_rng* g);
/* the same in pieces, for beliefs too large to hold at once: the per-particle loop over one block of
 * particles (the running total carried from block to block in the reference's order), then
 * WeightedFilter::normalize (WeightedFilter.cpp:130-143) over all weights */
double orc_is_propose(const orc_model* m, const orc_structs* st, orc_belief* b, int a, int o, orc_rng* g,
                      double running_total);
double orc_normalize(double* w, int64_t n, double total);
/* n_draws x WeightedFilter::sample with identical results in O(n + n_draws log n) */
void orc_weighted_sample_many(const double* w, int64_t n, double total_weight, orc_rng* g, int64_t n_draws,
                              int64_t* out);
/* test helper: position-sensitive 64-bit checksum of every particle's count block */
void orc_block_checksums(const float* counts, int64_t n, int64_t stride, uint64_t* out);
/* WeightedFilter::sample (WeightedFilter.cpp:163-191): index of the drawn particle */
int64_t orc_weighted_sample(const orc_belief* b, orc_rng* g);
/* importance_sampling::resample (ImportanceSampler.hpp:71-94) into dst (same N/stride);
 * ancestors (may be NULL) receives the drawn indices */
void orc_is_resample(const orc_belief* src, orc_belief* dst, orc_rng* g, int64_t* ancestors);
/* BAImportanceSampling::resetDomainStateDistribution (BAImportanceSampling.cpp:90-111) */
void orc_is_reset_domain_states(const orc_model* m, const orc_belief* src, orc_belief* dst, orc_rng* g,
                                int64_t* ancestors);
/* rejectSample (RejectionSampling.hpp:26-72) into dst; returns the number of attempts.
 * ancestors (may be NULL) receives the accepted particles' source indices. */
int64_t orc_reject_sample(const orc_model* m, const orc_structs* st, const orc_belief* src,
                          orc_belief* dst, int a, int o, orc_rng* g, int64_t* ancestors);
/* resetDomainState on every particle of a flat filter (BARejectionSampling.cpp:47-58) */
void orc_flat_reset_domain_states(const orc_model* m, orc_belief* b, orc_rng* g);

/* domain `mutate` kinds for reinvigoration */
enum {
    ORC_MUT_FACTORED_TIGER = 0,     /* flip one parent of O[listen][0]  FactoredTigerPriors.cpp:351-378 */
    ORC_MUT_COLLISION_AVOIDANCE = 1, /* CollisionAvoidancePriors.cpp:455-488 */
    ORC_MUT_SYSADMIN = 2,            /* SysAdminFactoredPrior.cpp:47-55 */
    ORC_MUT_GRIDWORLD = 3            /* GridWorldBAPriors.cpp:200-225 */
};
/* FBAPOMDP::mutate (the domain prior's mutate) on the parent bitmasks of one structure, in place */
void orc_mutate_structure(const orc_model* m, uint32_t* t_par, uint32_t* o_par, int mutate_kind, orc_rng* g);
/* incrementCountsOf(s, a, o, s', amount) (BAFlatModel.cpp:126-141, BABNModel.cpp:354-382) */
void orc_increment_counts(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts, int s,
                          int a, int o, int s2, float amount);
/* computePosterior of the MH structure beliefs (MHNIPS2018.cpp:41-109): replays a history of n_episodes
 * episodes (episode e has episode_len[e] steps; actions / observations concatenated) on `counts`;
 * returns the number of episode attempts, -1 if more than max_attempts were needed */
int64_t orc_mh_replay_history(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                              int n_episodes, const int32_t* episode_len, const int32_t* actions,
                              const int32_t* observations, orc_rng* g, int64_t max_attempts, int32_t* last_state);
/* DBNNode::marginalizeOut (DBNNode.cpp:40-80): src node must be fully connected (parents = all
 * state features) or equal to dst. */
void orc_marginalize_node(const orc_model* m, uint32_t src_par, const float* src, uint32_t dst_par,
                          int range, float* dst);
/* ReinvigoratingRejectionSampling::reinvigorateParticles (…RejectionSampling.cpp:121-131) */
int orc_reinvigorate(const orc_model* m, orc_structs* st, orc_belief* belief, const orc_belief* fc,
                     int64_t amount, int mutate_kind, orc_rng* g);

/* `amount` x breed (factored/ReinvigoratingRejectionSampling.cpp:24-35): structure donor from `belief`,
 * counts donor from `fc`; the k-th bred particle goes to dst_slots[k] of `dst` (a weighted dst follows
 * WeightedFilter::replace), or with dst_slots == NULL to a uniformly drawn slot */
int orc_breed_into(const orc_model* m, orc_structs* st, orc_belief* dst, const int64_t* dst_slots,
                   orc_belief* belief, const orc_belief* fc, int64_t amount, int mutate_kind, orc_rng* g);
void orc_replace_weight(orc_belief* b, int64_t slot);
/* CheatingReinvigoration::cheat (prototypes/CheatingReinvigoration.cpp:136-147) */
void orc_cheat(orc_belief* belief, const orc_belief* correct, int64_t amount, orc_rng* g);
/* WeightedFilter::leastLikely (WeightedFilter.cpp:206-243), ties in libstdc++'s priority_queue order */
void orc_least_likely(const double* w, int64_t n_particles, int64_t n, int64_t* out);
/* StructureIncubatorSampling::reinvigorateBelief (factored/StructureIncubatorSampling.cpp:155-187) */
int64_t orc_promote(orc_belief* shadow, orc_belief* belief, double threshold, orc_rng* g);

/* MHwithinGibbs (factored/MHwithinGibbs.cpp): a state history (episode_len[e] + 1 states per episode)
 * conditioned on one model and the (action, observation) history — by rejection sampling (:38-94; returns the
 * episode attempts, -1 beyond max_attempts) or by backward messages + forward sampling (:96-213;
 * state_prior = FBAPOMDP::domainStatePrior()->prob(s), S floats) — and computePosteriorCounts (:397-436) */
int64_t orc_state_history_rs(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                             int n_episodes, const int32_t* episode_len, const int32_t* actions,
                             const int32_t* observations, orc_rng* g, int64_t max_attempts, int32_t* states_out);
int orc_state_history_msg(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                          const float* state_prior, int n_episodes, const int32_t* episode_len,
                          const int32_t* actions, const int32_t* observations, orc_rng* g, int32_t* states_out);
/* BABNModel::flattenT / flattenO (BABNModel.cpp:89-178): T[s][a][s'], O[a][s'][o] as floats */
void orc_flatten_model(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                       float* T, float* O);
void orc_add_history_counts(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                            int n_episodes, const int32_t* episode_len, const int32_t* actions,
                            const int32_t* observations, const int32_t* states);
/* NestedBelief::updateEstimation for ONE top particle (NestedBelief.cpp:142-187); returns its attempts */
int64_t orc_nested_update_particle(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                                   const int32_t* states_in, int32_t* states_out, int64_t n_bottom, int a, int o,
                                   orc_rng* g, int64_t max_attempts);

/* RBAPOUCT::rollout (RBAPOUCT.cpp:295-323) on a read-only particle (KeepCounts) */
double orc_rollout(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par,
                   const float* counts, int start_state, int depth, double discount, orc_rng* g);

#ifdef __cplusplus
}
#endif
#endif
