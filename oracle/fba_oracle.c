/* TEST INFRASTRUCTURE — the CPU oracle (see fba_oracle.h). Not shipped, never on the product path.
 *
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64 (no FMA), so every
 * multiply and add below is separately rounded, as there.
 */
#include "fba_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* RNG: libstdc++ (GCC 13) distributions over a stream of mt19937 words                        */
/* ------------------------------------------------------------------------------------------ */

static uint32_t next_word(orc_rng* g)
{
    if (g->cur >= g->n)
    {
        g->overrun = 1;
        g->cur++;
        return 0u;
    }
    return g->words[g->cur++];
}

/* rnd::uniform_rand01 (random.cpp:100-103) = uniform_real_distribution<double>(0,1)
 * = generate_canonical<double,53> (bits/random.tcc:3349-3381): two 32-bit words,
 * (w0 + w1 * 2^32) / 2^64 in double arithmetic, clamped below 1. */
double orc_uniform01(orc_rng* g)
{
    double sum = 0.0, tmp = 1.0;
    sum += (double)next_word(g) * tmp;
    tmp *= 4294967296.0;
    sum += (double)next_word(g) * tmp;
    tmp *= 4294967296.0;
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

/* rnd::boolean (random.cpp:90-93) = bernoulli_distribution(0.5) (bits/random.h:3741-3751):
 * generate_canonical<double> < 0.5 */
int orc_boolean(orc_rng* g)
{
    return orc_uniform01(g) < 0.5;
}

/* uniform_int_distribution<int>(0, range-1) on a 32-bit engine
 * (bits/uniform_int_dist.h:257-281, Lemire's nearly-divisionless method) */
int32_t orc_uniform_int(orc_rng* g, uint32_t range)
{
    uint64_t product = (uint64_t)next_word(g) * (uint64_t)range;
    uint32_t low     = (uint32_t)product;
    if (low < range)
    {
        uint32_t threshold = (uint32_t)(-range) % range;
        while (low < threshold)
        {
            product = (uint64_t)next_word(g) * (uint64_t)range;
            low     = (uint32_t)product;
            if (g->overrun) break;
        }
    }
    return (int32_t)(product >> 32);
}

/* rnd::slowRandomInt(0, max) (random.cpp:111-115) */
static int slow_random_int(orc_rng* g, int max)
{
    return (int)floor(orc_uniform01(g) * (double)max);
}

/* ------------------------------------------------------------------------------------------ */
/* Dirichlet rows in expected mode                                                             */
/* ------------------------------------------------------------------------------------------ */

/* rnd::sample::Dir::sampleFromExpectedMult (random.cpp:244-255) feeding
 * sampleFromMult<float const> (random.hpp:93-115): DOUBLE total, FLOAT running prefix. */
static int sample_expected_mult(const float* dir, int n, orc_rng* g)
{
    double total = dir[0];
    for (int i = 1; i < n; ++i) total += dir[i];

    double const p = orc_uniform01(g) * total;
    float sum      = dir[0];
    for (int i = 1; i < n; ++i)
    {
        if (p < sum) return i - 1;
        sum += dir[i];
    }
    return n - 1;
}

/* rnd::sample::Dir::expectedMult(dir, n)[k] (random.cpp:257-279): FLOAT sum, FLOAT divide */
static float expected_mult_at(const float* dir, int n, int k)
{
    float sum = dir[0];
    for (int i = 1; i < n; ++i) sum += dir[i];
    if (sum <= 1e-300) return 0.0f;
    return dir[k] / sum;
}

/* utils::categoricalDistr::sample (distributions.cpp:47-51): sampleFromMult<float const> with the
 * stored double total */
static int sample_from_mult_f(const float* mult, int n, double total, orc_rng* g)
{
    double const p = orc_uniform01(g) * total;
    float sum      = mult[0];
    for (int i = 1; i < n; ++i)
    {
        if (p < sum) return i - 1;
        sum += mult[i];
    }
    return n - 1;
}

/* ------------------------------------------------------------------------------------------ */
/* indexing (utils/index.cpp)                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* indexing::projectUsingStepSize (index.cpp:98-119): feature 0 is the most significant digit */
static void features_of(int v, const int32_t* sizes, int n, int* out)
{
    if (n == 1)
    {
        out[0] = v;
        return;
    }
    int step[ORC_MAXF];
    step[n - 1] = 1;
    for (int i = n - 2; i >= 0; --i) step[i] = step[i + 1] * sizes[i + 1]; /* index.cpp:18-49 */
    for (int i = 0; i < n; ++i)
    {
        out[i] = v / step[i];
        v      = v % step[i];
    }
}

/* indexing::project (index.cpp:51-83) */
static int project(const int* vals, const int32_t* sizes, int n)
{
    int r = 0;
    for (int i = 0; i < n; ++i) r = r * sizes[i] + vals[i];
    return r;
}

/* DBNNode::cptIndex(graph input, 0) / output size (DBNNode.cpp:171-205): mixed radix over the
 * node's parents in ascending feature order */
static int parent_config(const orc_model* m, uint32_t par, const int* x)
{
    int r = 0;
    for (int f = 0; f < m->FS; ++f)
        if (par & (1u << f)) r = r * m->feat_s[f] + x[f];
    return r;
}

static int64_t num_parent_configs(const orc_model* m, uint32_t par)
{
    int64_t r = 1;
    for (int f = 0; f < m->FS; ++f)
        if (par & (1u << f)) r *= m->feat_s[f];
    return r;
}

int64_t orc_struct_offsets(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, int64_t* off)
{
    int64_t k  = 0;
    int const J = m->FS + m->FO;
    for (int a = 0; a < m->A; ++a)
    {
        for (int f = 0; f < m->FS; ++f)
        {
            if (off) off[a * J + f] = k;
            k += num_parent_configs(m, t_par[a * m->FS + f]) * m->feat_s[f];
        }
        for (int q = 0; q < m->FO; ++q)
        {
            if (off) off[a * J + m->FS + q] = k;
            k += num_parent_configs(m, o_par[a * m->FO + q]) * m->feat_o[q];
        }
    }
    return k;
}

int64_t orc_struct_size(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par)
{
    return orc_struct_offsets(m, t_par, o_par, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* domain functors                                                                             */
/* ------------------------------------------------------------------------------------------ */

double orc_reward(const orc_model* m, int s, int a, int s2, int* terminal)
{
    switch (m->domain)
    {
        case ORC_DOM_TIGER: /* TigerBAExtension.cpp:21-44; OBSERVE = 2 (Tiger.hpp:30) */
            *terminal = m->dom_ip[0] && a != 2;
            if (a == 2) return -1.0;
            return (a == s) ? 10.0 : -100.0;
        case ORC_DOM_FACTORED_TIGER: { /* FactoredTigerBAExtension.cpp:27-56 */
            *terminal = m->dom_ip[0] && a != 2;
            if (a == 2) return -1.0;
            int const loc = (s < m->S / 2) ? 0 : 1;
            return (a == loc) ? 10.0 : -100.0;
        }
        case ORC_DOM_SYSADMIN: { /* SysAdminBAExtension.cpp:27-48: float arithmetic */
            *terminal    = 0;
            int up       = __builtin_popcount((unsigned)s2);
            unsigned reb = (unsigned)(a >= m->dom_ip[0]);
            return (double)((float)up - (float)m->dom_dp[0] * (float)reb);
        }
        case ORC_DOM_GRIDWORLD: { /* GridWorldBAExtension.cpp:74-100: depends on s only */
            int const size = m->dom_ip[0], G = m->dom_ip[1];
            int const g = s % G, y = (s / G) % size, x = s / (G * size);
            int const at_goal = (x == m->dom_ip[2 + 2 * g]) && (y == m->dom_ip[3 + 2 * g]);
            *terminal         = at_goal;
            return at_goal ? m->dom_dp[0] : m->dom_dp[1];
        }
        case ORC_DOM_COLLISION_AVOIDANCE: { /* CollisionAvoidanceBAExtension.cpp:59-89 */
            int const H = m->dom_ip[1], nobs = m->dom_ip[2];
            int obst_space = 1;
            for (int i = 0; i < nobs; ++i) obst_space *= H;
            int const x = s2 / (H * obst_space), y = (s2 / obst_space) % H;
            int obst = s2 % obst_space;
            int crashed = 0;
            if (x < nobs)
            { /* obstacle x's position: digit x of the mixed-radix obstacle index (0 = most sig.) */
                int pos = obst;
                for (int i = nobs - 1; i > x; --i) pos /= H;
                crashed = (y == pos % H);
            }
            *terminal = crashed || x == 0;
            if (crashed) return -m->dom_dp[1];
            return (a == 1) ? 0.0 : -m->dom_dp[0]; /* STAY = 1 (CollisionAvoidance.hpp:91) */
        }
        default: { /* ORC_DOM_TABLE */
            int t = 0;
            double r = 0;
            if (m->rew_sa) r += m->rew_sa[(int64_t)s * m->A + a];
            if (m->rew_as2) r += m->rew_as2[(int64_t)a * m->S + s2];
            if (m->term_sa) t |= m->term_sa[(int64_t)s * m->A + a];
            if (m->term_as2) t |= m->term_as2[(int64_t)a * m->S + s2];
            *terminal = t;
            return r;
        }
    }
}

int orc_sample_start_state(const orc_model* m, orc_rng* g)
{
    switch (m->start_kind)
    {
        case ORC_START_CONST: return m->start_ip[0];
        case ORC_START_BOOL: return orc_boolean(g) ? m->start_ip[0] : m->start_ip[1];
        case ORC_START_UNIFORM_INT: return orc_uniform_int(g, (uint32_t)m->start_ip[0]);
        case ORC_START_SLOW2: {
            int const i = slow_random_int(g, m->start_ip[0]);
            int const j = slow_random_int(g, m->start_ip[1]);
            return m->start_table[i * m->start_ip[1] + j];
        }
        default: return sample_from_mult_f(m->start_values, m->start_ip[0], m->start_total, g);
    }
}

static int random_action(const orc_model* m, orc_rng* g)
{
    if (m->action_draw == ORC_ACT_SLOW_INT) return slow_random_int(g, m->A);
    return orc_uniform_int(g, (uint32_t)m->A);
}

/* ------------------------------------------------------------------------------------------ */
/* the hyper-state step                                                                        */
/* ------------------------------------------------------------------------------------------ */

static int64_t node_offsets_small(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par,
                                  int a, int64_t* t_off, int64_t* o_off)
{ /* offsets of action a's nodes only (walks the structure up to a) */
    int64_t k = 0;
    for (int aa = 0; aa <= a; ++aa)
    {
        for (int f = 0; f < m->FS; ++f)
        {
            if (aa == a) t_off[f] = k;
            k += num_parent_configs(m, t_par[aa * m->FS + f]) * m->feat_s[f];
        }
        for (int q = 0; q < m->FO; ++q)
        {
            if (aa == a) o_off[q] = k;
            k += num_parent_configs(m, o_par[aa * m->FO + q]) * m->feat_o[q];
        }
    }
    return k;
}

double orc_step(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                int32_t* state, int a, int update_counts, orc_rng* g, int* o_out, int* terminal)
{
    int x[ORC_MAXF], x2[ORC_MAXF], of[ORC_MAXF];
    int64_t t_off[ORC_MAXF], o_off[ORC_MAXF];
    int const s = *state;
    node_offsets_small(m, t_par, o_par, a, t_off, o_off);
    features_of(s, m->feat_s, m->FS, x);

    /* BA{Flat,BN}Model::sampleStateIndex (BAFlatModel.cpp:83-91, BABNModel.cpp:292-307):
     * one draw per state feature, in feature order, parents taken from the OLD state */
    for (int f = 0; f < m->FS; ++f)
    {
        const float* row =
            counts + t_off[f] + (int64_t)parent_config(m, t_par[a * m->FS + f], x) * m->feat_s[f];
        x2[f] = sample_expected_mult(row, m->feat_s[f], g);
    }
    int const s2 = project(x2, m->feat_s, m->FS);

    /* sampleObservationIndex (BAFlatModel.cpp:93-103, BABNModel.cpp:309-326): parents = NEW state */
    for (int q = 0; q < m->FO; ++q)
    {
        const float* row =
            counts + o_off[q] + (int64_t)parent_config(m, o_par[a * m->FO + q], x2) * m->feat_o[q];
        of[q] = sample_expected_mult(row, m->feat_o[q], g);
    }
    int const o = project(of, m->feat_o, m->FO);

    double const r = orc_reward(m, s, a, s2, terminal); /* BAPOMDP.cpp:131-132 */

    if (update_counts)
    { /* incrementCountsOf(s, a, o, s') (BAPOMDP.cpp:134-137) */
        for (int f = 0; f < m->FS; ++f)
            counts[t_off[f] + (int64_t)parent_config(m, t_par[a * m->FS + f], x) * m->feat_s[f] + x2[f]] +=
                1.0f;
        /* tabular: psi[a][s'][o] (BAFlatModel.cpp:126-141). Factored: the observation CPTs are
         * indexed with the OLD state's features (BABNModel.cpp:366,380) — reproduced, not fixed. */
        const int* xo = m->tabular ? x2 : x;
        for (int q = 0; q < m->FO; ++q)
            counts[o_off[q] + (int64_t)parent_config(m, o_par[a * m->FO + q], xo) * m->feat_o[q] + of[q]] +=
                1.0f;
    }

    *state = s2; /* BAPOMDP.cpp:139-140 */
    *o_out = o;
    return r;
}

void orc_increment_counts(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts, int s,
                          int a, int o, int s2, float amount)
{
    /* BA{Flat,BN}Model::incrementCountsOf(s, a, o, s', amount) (BAFlatModel.cpp:126-141,
     * BABNModel.cpp:354-382): one cell per node; the factored observation CPTs are indexed with the
     * OLD state's features (BABNModel.cpp:366,380) */
    int x[ORC_MAXF], x2[ORC_MAXF], of[ORC_MAXF];
    int64_t t_off[ORC_MAXF], o_off[ORC_MAXF];
    node_offsets_small(m, t_par, o_par, a, t_off, o_off);
    features_of(s, m->feat_s, m->FS, x);
    features_of(s2, m->feat_s, m->FS, x2);
    features_of(o, m->feat_o, m->FO, of);
    for (int f = 0; f < m->FS; ++f)
        counts[t_off[f] + (int64_t)parent_config(m, t_par[a * m->FS + f], x) * m->feat_s[f] + x2[f]] += amount;
    const int* xo = m->tabular ? x2 : x;
    for (int q = 0; q < m->FO; ++q)
        counts[o_off[q] + (int64_t)parent_config(m, o_par[a * m->FO + q], xo) * m->feat_o[q] + of[q]] += amount;
}

int64_t orc_mh_replay_history(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                              int n_episodes, const int32_t* episode_len, const int32_t* actions,
                              const int32_t* observations, orc_rng* g, int64_t max_attempts, int32_t* last_state)
{
    /* computePosterior (MHNIPS2018.cpp:41-109): replay the whole (action, observation) history on a
     * model, episode by episode. Each attempt of an episode draws a domain start state, then per step
     * samples s' and o from the model (sampleStateIndex / sampleObservationIndex, expected mode);
     * a wrong observation abandons the attempt — the increments made so far in this episode are taken
     * back with amount -1, in step order — and the episode is tried again; a right one increments
     * (s, a, o, s'). Returns the number of episode attempts (-1 if max_attempts was exceeded);
     * *last_state = the state after the last step. */
    int32_t trans_s[4096], trans_s2[4096];
    int64_t attempts = 0;
    int32_t new_s    = 0;
    int64_t first    = 0;
    for (int e = 0; e < n_episodes; ++e)
    {
        int const len = episode_len[e];
        if (len > 4096) return -2;
        for (;;)
        {
            if (++attempts > max_attempts) return -1;
            int32_t s   = orc_sample_start_state(m, g);
            int applied = 0;
            for (int t = 0; t < len; ++t)
            {
                int const a   = actions[first + t];
                int32_t state = s;
                int o, term;
                orc_step(m, t_par, o_par, counts, &state, a, 0, g, &o, &term); /* KeepCounts: draws only */
                new_s = state;
                if (o != observations[first + t]) break;
                orc_increment_counts(m, t_par, o_par, counts, s, a, o, new_s, 1.0f);
                trans_s[applied] = s, trans_s2[applied] = new_s;
                ++applied;
                s = new_s;
            }
            if (applied == len) break;
            for (int t = 0; t < applied; ++t)
                orc_increment_counts(m, t_par, o_par, counts, trans_s[t], actions[first + t],
                                     observations[first + t], trans_s2[t], -1.0f);
        }
        first += len;
    }
    *last_state = new_s;
    return attempts;
}

double orc_obs_prob(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par,
                    const float* counts, int state, int a, int o)
{
    int x[ORC_MAXF], of[ORC_MAXF];
    int64_t t_off[ORC_MAXF], o_off[ORC_MAXF];
    node_offsets_small(m, t_par, o_par, a, t_off, o_off);

    if (m->tabular)
    { /* BAFlatModel::computeObservationProbability (BAFlatModel.cpp:105-124) */
        if (m->O == 1) return 1.0;
        return expected_mult_at(counts + o_off[0] + (int64_t)state * m->O, m->O, o);
    }
    /* BABNModel::computeObservationProbability (BABNModel.cpp:328-352): double product of float
     * factors, parents = features of the (new) state */
    features_of(state, m->feat_s, m->FS, x);
    features_of(o, m->feat_o, m->FO, of);
    double prob = 1;
    for (int q = 0; q < m->FO; ++q)
    {
        const float* row =
            counts + o_off[q] + (int64_t)parent_config(m, o_par[a * m->FO + q], x) * m->feat_o[q];
        prob *= expected_mult_at(row, m->feat_o[q], of[q]);
    }
    return prob;
}

/* rnd::math::logGamma (random.cpp:127-135) */
static double log_gamma_ref(double x)
{
    return (x < 1) ? 0.0 : lgamma(x);
}

double orc_log_bd_score(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                        const float* prior_counts)
{
    /* BABNModel::LogBDScore: for a, for transition nodes, for observation nodes — the block's own order */
    double total = 0;
    int64_t off  = 0;
    for (int a = 0; a < m->A; ++a)
        for (int j = 0; j < m->FS + m->FO; ++j)
        {
            uint32_t const par = (j < m->FS) ? t_par[a * m->FS + j] : o_par[a * m->FO + (j - m->FS)];
            int const range    = (j < m->FS) ? m->feat_s[j] : m->feat_o[j - m->FS];
            int64_t cfgs       = 1;
            for (int f = 0; f < m->FS; ++f)
                if (par & (1u << f)) cfgs *= m->feat_s[f];
            double node = 0; /* DBNNode::LogBDScore */
            for (int64_t r = 0; r < cfgs; ++r)
            {
                double distr_total = 0, prior_total = 0;
                for (int v = 0; v < range; ++v)
                {
                    int64_t const i = off + r * range + v;
                    distr_total += counts[i];
                    prior_total += prior_counts[i];
                    node += log_gamma_ref(counts[i]) - log_gamma_ref(prior_counts[i]);
                }
                node += log_gamma_ref(prior_total) - log_gamma_ref(distr_total);
            }
            total += node;
            off += cfgs * range;
        }
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* importance sampling                                                                         */
/* ------------------------------------------------------------------------------------------ */

static const uint32_t* tpar_of(const orc_model* m, const orc_structs* st, int id)
{
    return st->t_par + (int64_t)id * m->A * m->FS;
}
static const uint32_t* opar_of(const orc_model* m, const orc_structs* st, int id)
{
    return st->o_par + (int64_t)id * m->A * m->FO;
}

double orc_is_propose(const orc_model* m, const orc_structs* st, orc_belief* b, int a, int o, orc_rng* g,
                      double running_total)
{
    /* ImportanceSampler.hpp:36-54, particles in index order. running_total: the sum over the particles
     * BEFORE this block, so that a belief processed block after block accumulates `total_weight +=
     * p->w` in exactly the reference's order */
    double total_weight = running_total;
    for (int64_t i = 0; i < b->N; ++i)
    {
        const uint32_t* tp = tpar_of(m, st, b->struct_id[i]);
        const uint32_t* op = opar_of(m, st, b->struct_id[i]);
        float* c           = b->counts + i * b->stride;
        int sim_o, term;
        orc_step(m, tp, op, c, &b->state[i], a, 1, g, &sim_o, &term);
        b->w[i] *= orc_obs_prob(m, tp, op, c, b->state[i], a, o);
        total_weight += b->w[i];
    }
    return total_weight;
}

double orc_normalize(double* w, int64_t n, double total)
{
    /* WeightedFilter::normalize(total) (WeightedFilter.cpp:130-143): returns the new _total_weight */
    double acc = 0;
    for (int64_t i = 0; i < n; ++i)
    {
        w[i] /= total;
        acc += w[i];
    }
    return acc;
}

double orc_is_update(const orc_model* m, const orc_structs* st, orc_belief* b, int a, int o, orc_rng* g)
{
    double const total_weight = orc_is_propose(m, st, b, a, o, g, 0.0);
    b->total_weight           = orc_normalize(b->w, b->N, total_weight);
    return total_weight;
}

void orc_weighted_sample_many(const double* w, int64_t n, double total_weight, orc_rng* g, int64_t n_draws,
                              int64_t* out)
{
    /* n_draws x WeightedFilter::sample (WeightedFilter.cpp:163-191) in O(n + n_draws log n) instead of
     * O(n n_draws), with IDENTICAL results: the reference walks sample = n-1 .. 1, subtracting
     * w[sample] from `remaining_weight` (starting at _total_weight) and stops at the first index whose
     * remainder R[sample] is below the threshold. The remainders do not depend on the draw, so they
     * are computed once, by the same sequential subtractions; and because every w >= 0, rounding
     * keeps R non-increasing as the index goes down, so {k >= 1 : R[k] < threshold} is a prefix
     * 1..m and "first hit walking down" = m = found by binary search (0 if the set is empty).
     * tests/test_oracle_vs_golden.py pins this against orc_weighted_sample draw by draw. */
    double* R = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double rem = total_weight;
    for (int64_t k = n - 1; k >= 1; --k)
    {
        rem -= w[k];
        R[k] = rem;
    }
    for (int64_t j = 0; j < n_draws; ++j)
    {
        double const threshold = orc_uniform01(g) * total_weight;
        int64_t lo = 1, hi = n; /* number of k in [1, n) with R[k] < threshold, plus one */
        while (lo < hi)
        {
            int64_t const mid = (lo + hi) >> 1;
            if (threshold > R[mid]) lo = mid + 1;
            else
                hi = mid;
        }
        out[j] = lo - 1;
    }
    free(R);
}

void orc_block_checksums(const float* counts, int64_t n, int64_t stride, uint64_t* out)
{
    /* test helper: a position-sensitive 64-bit checksum of every particle's count block (bit
     * patterns, wrapping integer arithmetic — reproducible on any device) */
    for (int64_t i = 0; i < n; ++i)
    {
        const uint32_t* c = (const uint32_t*)(counts + i * stride);
        uint64_t h        = 0;
        for (int64_t k = 0; k < stride; ++k) h += (uint64_t)c[k] * (2 * (uint64_t)k + 1) * 0x9E3779B97F4A7C15ull;
        out[i] = h;
    }
}

int64_t orc_weighted_sample(const orc_belief* b, orc_rng* g)
{
    /* WeightedFilter.cpp:163-191 */
    double const threshold = orc_uniform01(g) * b->total_weight;
    int64_t sample         = b->N - 1;
    double remaining       = b->total_weight;
    for (; sample > 0; --sample)
    {
        remaining -= b->w[sample];
        if (threshold > remaining) break;
    }
    return sample;
}

static void copy_particle(const orc_belief* src, int64_t i, orc_belief* dst, int64_t j)
{
    memcpy(dst->counts + j * dst->stride, src->counts + i * src->stride, sizeof(float) * src->stride);
    dst->state[j]     = src->state[i];
    dst->struct_id[j] = src->struct_id[i];
}

void orc_is_resample(const orc_belief* src, orc_belief* dst, orc_rng* g, int64_t* ancestors)
{
    /* ImportanceSampler.hpp:71-94; WeightedFilter::add (WeightedFilter.cpp:60-66) */
    int64_t const n = src->N;
    double const w  = 1 / (double)n;
    double total    = 0;
    for (int64_t j = 0; j < n; ++j)
    {
        int64_t const i = orc_weighted_sample(src, g);
        copy_particle(src, i, dst, j);
        dst->w[j] = w;
        total += w;
        if (ancestors) ancestors[j] = i;
    }
    dst->total_weight = total;
}

void orc_is_reset_domain_states(const orc_model* m, const orc_belief* src, orc_belief* dst, orc_rng* g,
                                int64_t* ancestors)
{
    /* BAImportanceSampling.cpp:90-111: per new particle one weighted draw, then the domain's
     * start-state draws (BAPOMDP::resetDomainState, BAPOMDP.cpp:69-77) */
    int64_t const n = src->N;
    double const w  = 1.0 / (double)n;
    double total    = 0;
    for (int64_t j = 0; j < n; ++j)
    {
        int64_t const i = orc_weighted_sample(src, g);
        copy_particle(src, i, dst, j);
        dst->state[j] = orc_sample_start_state(m, g);
        dst->w[j]     = w;
        total += w;
        if (ancestors) ancestors[j] = i;
    }
    dst->total_weight = total;
}

/* ------------------------------------------------------------------------------------------ */
/* rejection sampling                                                                          */
/* ------------------------------------------------------------------------------------------ */

int64_t orc_reject_sample(const orc_model* m, const orc_structs* st, const orc_belief* src,
                          orc_belief* dst, int a, int o, orc_rng* g, int64_t* ancestors)
{
    /* RejectionSampling.hpp:26-72; FlatFilter::sample (FlatFilter.cpp:97-102) */
    int64_t accepted = 0, attempts = 0;
    float* scratch = (float*)malloc(sizeof(float) * src->stride);
    while (accepted < src->N)
    {
        int64_t const i = orc_uniform_int(g, (uint32_t)src->N);
        if (g->overrun) break;
        memcpy(scratch, src->counts + i * src->stride, sizeof(float) * src->stride);
        int32_t s = src->state[i];
        int sim_o, term;
        orc_step(m, tpar_of(m, st, src->struct_id[i]), opar_of(m, st, src->struct_id[i]), scratch, &s, a,
                 1, g, &sim_o, &term);
        if (sim_o == o)
        {
            memcpy(dst->counts + accepted * dst->stride, scratch, sizeof(float) * src->stride);
            dst->state[accepted]     = s;
            dst->struct_id[accepted] = src->struct_id[i];
            if (ancestors) ancestors[accepted] = i;
            ++accepted;
        }
        ++attempts;
    }
    free(scratch);
    return attempts;
}

void orc_flat_reset_domain_states(const orc_model* m, orc_belief* b, orc_rng* g)
{
    for (int64_t i = 0; i < b->N; ++i) b->state[i] = orc_sample_start_state(m, g);
}

/* ------------------------------------------------------------------------------------------ */
/* reinvigoration                                                                              */
/* ------------------------------------------------------------------------------------------ */

void orc_marginalize_node(const orc_model* m, uint32_t src_par, const float* src, uint32_t dst_par,
                          int range, float* dst)
{
    /* DBNNode::marginalizeOut (DBNNode.cpp:40-80). Walk the source node's parent configurations in
     * row order (indexing::increment, last parent fastest) and add each row onto the destination
     * row its values project to — float adds in that order. */
    int64_t const n_src = num_parent_configs(m, src_par);
    int64_t const n_dst = num_parent_configs(m, dst_par);
    if (src_par == dst_par)
    {
        memcpy(dst, src, sizeof(float) * n_src * range);
        return;
    }
    memset(dst, 0, sizeof(float) * n_dst * range);
    int x[ORC_MAXF];
    for (int64_t cfg = 0; cfg < n_src; ++cfg)
    {
        /* decode cfg over the source parents (ascending features, first = most significant) */
        int64_t rem = cfg;
        for (int f = m->FS - 1; f >= 0; --f)
        {
            x[f] = 0;
            if (src_par & (1u << f))
            {
                x[f] = (int)(rem % m->feat_s[f]);
                rem /= m->feat_s[f];
            }
        }
        int64_t const d = parent_config(m, dst_par, x);
        for (int v = 0; v < range; ++v) dst[d * range + v] += src[cfg * range + v];
    }
}

static int find_or_add_struct(const orc_model* m, orc_structs* st, const uint32_t* tp, const uint32_t* op)
{
    size_t const tb = sizeof(uint32_t) * m->A * m->FS, ob = sizeof(uint32_t) * m->A * m->FO;
    for (int i = 0; i < st->n_structs; ++i)
        if (!memcmp(tpar_of(m, st, i), tp, tb) && !memcmp(opar_of(m, st, i), op, ob)) return i;
    if (st->n_structs >= st->cap) return -1;
    int const id = st->n_structs++;
    memcpy(st->t_par + (int64_t)id * m->A * m->FS, tp, tb);
    memcpy(st->o_par + (int64_t)id * m->A * m->FO, op, ob);
    return id;
}

/* BABNModel::Structure::flip_random_edge (BABNModel.cpp:16-31) on a parent bitmask */
static uint32_t flip_random_edge(uint32_t par, int edge_range, orc_rng* g)
{
    return par ^ (1u << slow_random_int(g, edge_range));
}

void orc_mutate_structure(const orc_model* m, uint32_t* tp, uint32_t* op, int mutate_kind, orc_rng* g)
{
    /* FBAPOMDP::mutate -> the domain prior's mutate, on parent bitmasks */
    switch (mutate_kind)
    {
        case ORC_MUT_FACTORED_TIGER: /* FactoredTigerPriors.cpp:374-375: O[listen=2][0] */
            op[2 * m->FO + 0] = flip_random_edge(op[2 * m->FO + 0], m->FS, g);
            break;
        case ORC_MUT_COLLISION_AVOIDANCE: { /* CollisionAvoidancePriors.cpp:478-486 */
            int const a    = orc_uniform_int(g, (uint32_t)m->A);
            int const obst = 2 + orc_uniform_int(g, (uint32_t)m->dom_ip[2]);
            tp[a * m->FS + obst] = flip_random_edge(tp[a * m->FS + obst], m->FS, g);
            break;
        }
        case ORC_MUT_SYSADMIN: { /* SysAdminFactoredPrior.cpp:51-52: T[action][comp], the two
                                    subscripts' draws happen right to left (comp first) */
            int const comp = orc_uniform_int(g, (uint32_t)m->FS);
            int const a    = orc_uniform_int(g, (uint32_t)m->A);
            tp[a * m->FS + comp] = flip_random_edge(tp[a * m->FS + comp], m->FS, g);
            break;
        }
        default: { /* ORC_MUT_GRIDWORLD, GridWorldBAPriors.cpp:200-225: toggle the goal feature */
            int const a = slow_random_int(g, m->A);
            int const f = slow_random_int(g, 2);
            tp[a * m->FS + f] ^= (1u << (m->FS - 1));
            break;
        }
    }
}

void orc_replace_weight(orc_belief* b, int64_t slot)
{
    /* WeightedFilter::replace(i, s, dealloc) (WeightedFilter.cpp:71-90): the new particle weighs
     * _total_weight / size and _total_weight moves by the difference */
    if (!b->w) return;
    double const w    = b->total_weight / (double)b->N;
    double const diff = w - b->w[slot];
    b->w[slot]        = w;
    b->total_weight += diff;
}

int orc_breed_into(const orc_model* m, orc_structs* st, orc_belief* dst, const int64_t* dst_slots,
                   orc_belief* belief, const orc_belief* fc, int64_t amount, int mutate_kind, orc_rng* g)
{
    /* ReinvigoratingRejectionSampling.cpp:121-131 + breed (:24-35). Draw order (g++ -std=c++11
     * evaluates the call arguments right to left, SURVEY.md §7 hard part 3d):
     * fully-connected donor, structure donor, the domain's mutate draws, replacement slot. */
    int const nT = m->A * m->FS, nO = m->A * m->FO, J = m->FS + m->FO;
    uint32_t tp[ORC_MAXF * 64], op[ORC_MAXF * 64];
    int64_t off_src[64 * 2 * ORC_MAXF], off_dst[64 * 2 * ORC_MAXF];
    if (m->A > 64) return -2;
    float* fresh = (float*)malloc(sizeof(float) * dst->stride);

    for (int64_t k = 0; k < amount; ++k)
    {
        int64_t const counts_donor = orc_uniform_int(g, (uint32_t)fc->N);
        int64_t const struct_donor = orc_uniform_int(g, (uint32_t)belief->N);

        memcpy(tp, tpar_of(m, st, belief->struct_id[struct_donor]), sizeof(uint32_t) * nT);
        memcpy(op, opar_of(m, st, belief->struct_id[struct_donor]), sizeof(uint32_t) * nO);

        orc_mutate_structure(m, tp, op, mutate_kind, g);

        /* BABNModel::marginalizeOut (BABNModel.cpp:205-229) of the counts donor onto it */
        const uint32_t* stp = tpar_of(m, st, fc->struct_id[counts_donor]);
        const uint32_t* sop = opar_of(m, st, fc->struct_id[counts_donor]);
        orc_struct_offsets(m, stp, sop, off_src);
        int64_t const sz = orc_struct_offsets(m, tp, op, off_dst);
        if (sz > dst->stride)
        {
            free(fresh);
            return -3;
        }
        memset(fresh, 0, sizeof(float) * dst->stride);
        const float* src = fc->counts + counts_donor * fc->stride;
        for (int a = 0; a < m->A; ++a)
        {
            for (int f = 0; f < m->FS; ++f)
                orc_marginalize_node(m, stp[a * m->FS + f], src + off_src[a * J + f], tp[a * m->FS + f],
                                     m->feat_s[f], fresh + off_dst[a * J + f]);
            for (int q = 0; q < m->FO; ++q)
                orc_marginalize_node(m, sop[a * m->FO + q], src + off_src[a * J + m->FS + q],
                                     op[a * m->FO + q], m->feat_o[q], fresh + off_dst[a * J + m->FS + q]);
        }
        int const id = find_or_add_struct(m, st, tp, op);
        if (id < 0)
        {
            free(fresh);
            return -1;
        }
        int32_t const dom_state = belief->state[struct_donor]; /* breed: structure donor's state */

        /* FlatFilter::replace (FlatFilter.cpp:39-46): uniformly random slot; or the caller's slot
         * (StructureIncubatorSampling.cpp:74-80,139-153: WeightedFilter::add / replace) */
        int64_t const slot = dst_slots ? dst_slots[k] : orc_uniform_int(g, (uint32_t)dst->N);
        memcpy(dst->counts + slot * dst->stride, fresh, sizeof(float) * dst->stride);
        dst->state[slot]     = dom_state;
        dst->struct_id[slot] = id;
        orc_replace_weight(dst, slot);
    }
    free(fresh);
    return 0;
}

int orc_reinvigorate(const orc_model* m, orc_structs* st, orc_belief* belief, const orc_belief* fc,
                     int64_t amount, int mutate_kind, orc_rng* g)
{
    /* ReinvigoratingRejectionSampling::reinvigorateParticles (…RejectionSampling.cpp:121-131) */
    return orc_breed_into(m, st, belief, NULL, belief, fc, amount, mutate_kind, g);
}

/* ------------------------------------------------------------------------------------------ */
/* the composite structure beliefs' own steps                                                  */
/* ------------------------------------------------------------------------------------------ */

void orc_cheat(orc_belief* belief, const orc_belief* correct, int64_t amount, orc_rng* g)
{
    /* CheatingReinvigoration::cheat (prototypes/CheatingReinvigoration.cpp:136-147):
     * _belief.replace(rnd::slowRandomInt(0, size), copyState(_correct_structured_belief.sample()), ..)
     * — g++ evaluates the arguments right to left: the source particle is drawn first */
    for (int64_t k = 0; k < amount; ++k)
    {
        int64_t const src  = orc_uniform_int(g, (uint32_t)correct->N);
        int64_t const slot = slow_random_int(g, (int)belief->N);
        copy_particle(correct, src, belief, slot);
        orc_replace_weight(belief, slot);
    }
}

/* std::priority_queue<pair<double,int>, vector, Less-on-first> as libstdc++ implements it
 * (bits/stl_heap.h: __push_heap, __adjust_heap, __pop_heap), because WeightedFilter::leastLikely's
 * result order on equal weights is whatever that container does */
typedef struct { double w; int i; } heap_el;
static void heap_push(heap_el* h, int64_t* n, heap_el v)
{
    int64_t hole = (*n)++;
    int64_t parent = (hole - 1) / 2;
    while (hole > 0 && h[parent].w < v.w)
    {
        h[hole] = h[parent];
        hole    = parent;
        parent  = (hole - 1) / 2;
    }
    h[hole] = v;
}
static heap_el heap_pop(heap_el* h, int64_t* n)
{
    /* __pop_heap: the last element's value is sifted in from the root after the top moved out */
    heap_el const top = h[0];
    int64_t const len = --(*n);
    if (len == 0) return top;
    heap_el const v = h[len];
    /* __adjust_heap(first, hole = 0, len, v) */
    int64_t hole = 0, child = 0;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (h[child].w < h[child - 1].w) child--;
        h[hole] = h[child];
        hole    = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child   = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole    = child - 1;
    }
    /* __push_heap(first, hole, top = 0, v) */
    int64_t parent = (hole - 1) / 2;
    while (hole > 0 && h[parent].w < v.w)
    {
        h[hole] = h[parent];
        hole    = parent;
        parent  = (hole - 1) / 2;
    }
    h[hole] = v;
    return top;
}

void orc_least_likely(const double* w, int64_t n_particles, int64_t n, int64_t* out)
{
    /* WeightedFilter::leastLikely (WeightedFilter.cpp:206-243): the queue is seeded with the first n
     * particles; then EVERY particle (the first n again) replaces the current largest if it is
     * strictly lighter; the n survivors are popped largest first */
    heap_el* h = (heap_el*)malloc(sizeof(heap_el) * (size_t)(n + 1));
    int64_t len = 0;
    for (int64_t i = 0; i < n; ++i) heap_push(h, &len, (heap_el){w[i], (int)i});
    for (int64_t i = 0; i < n_particles; ++i)
        if (w[i] < h[0].w)
        {
            heap_pop(h, &len);
            heap_push(h, &len, (heap_el){w[i], (int)i});
        }
    for (int64_t i = 0; i < n; ++i) out[i] = heap_pop(h, &len).i;
    free(h);
}

int64_t orc_promote(orc_belief* shadow, orc_belief* belief, double threshold, orc_rng* g)
{
    /* StructureIncubatorSampling::reinvigorateBelief (factored/StructureIncubatorSampling.cpp:155-187) */
    int64_t moved = 0;
    for (int64_t i = 0; i < shadow->N; ++i)
        if (shadow->w[i] / shadow->total_weight > threshold) /* WeightedFilter::normalizedWeight */
        {
            int64_t const slot = orc_uniform_int(g, (uint32_t)belief->N); /* FlatFilter::replace */
            copy_particle(shadow, i, belief, slot);
            shadow->w[i] = 0;
            ++moved;
        }
    if (moved)
    { /* WeightedFilter::normalize() (WeightedFilter.cpp:118-143) */
        double total = 0;
        for (int64_t i = 0; i < shadow->N; ++i) total += shadow->w[i];
        shadow->total_weight = orc_normalize(shadow->w, shadow->N, total);
    }
    return moved;
}

/* ------------------------------------------------------------------------------------------ */
/* MHwithinGibbs: state histories conditioned on a model, posterior counts                     */
/* ------------------------------------------------------------------------------------------ */

int64_t orc_state_history_rs(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                             int n_episodes, const int32_t* episode_len, const int32_t* actions,
                             const int32_t* observations, orc_rng* g, int64_t max_attempts, int32_t* states_out)
{
    /* rejectionSampleStateHistory (MHwithinGibbs.cpp:38-94): per episode a domain start state, then
     * s' and o sampled from the model (expected Dirichlets, counts untouched) step after step; the first
     * wrong observation abandons the attempt and the episode starts over. states_out receives
     * episode_len[e] + 1 states per episode. Returns the episode attempts, -1 beyond max_attempts. */
    int64_t attempts = 0, first = 0, pos = 0;
    for (int e = 0; e < n_episodes; ++e)
    {
        int const len = episode_len[e];
        for (;;)
        {
            if (++attempts > max_attempts || g->overrun) return -1;
            int32_t s       = orc_sample_start_state(m, g);
            states_out[pos] = s;
            int t           = 0;
            for (; t < len; ++t)
            {
                int o, term;
                orc_step(m, t_par, o_par, (float*)counts, &s, actions[first + t], 0, g, &o, &term);
                if (o != observations[first + t]) break;
                states_out[pos + 1 + t] = s;
            }
            if (t == len) break;
        }
        first += len;
        pos += len + 1;
    }
    return attempts;
}

static int sample_from_mult_d(const double* mult, int64_t n, double total, orc_rng* g)
{
    /* rnd::sample::Dir::sampleFromMult<double> (random.hpp:93-115) */
    double const p = orc_uniform01(g) * total;
    double sum     = mult[0];
    for (int64_t i = 1; i < n; ++i)
    {
        if (p < sum) return (int)(i - 1);
        sum += mult[i];
    }
    return (int)(n - 1);
}

void orc_flatten_model(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                       float* T /* [S][A][S] */, float* O /* [A][S][O] */)
{
    /* BABNModel::flattenT / flattenO (BABNModel.cpp:89-178): every entry starts at 1.0f and is multiplied,
     * feature after feature, by that node's expected multinomial (DBNNode::expectation = expectedMult:
     * float sum, float divide) */
    int x[ORC_MAXF], x2[ORC_MAXF], of[ORC_MAXF];
    int64_t t_off[ORC_MAXF], o_off[ORC_MAXF];
    for (int a = 0; a < m->A; ++a)
    {
        node_offsets_small(m, t_par, o_par, a, t_off, o_off);
        for (int s = 0; s < m->S; ++s)
        {
            features_of(s, m->feat_s, m->FS, x);
            for (int s2 = 0; s2 < m->S; ++s2)
            {
                features_of(s2, m->feat_s, m->FS, x2);
                float p = 1.0f;
                for (int f = 0; f < m->FS; ++f)
                    p *= expected_mult_at(counts + t_off[f] + (int64_t)parent_config(m, t_par[a * m->FS + f], x) * m->feat_s[f],
                                          m->feat_s[f], x2[f]);
                T[((int64_t)s * m->A + a) * m->S + s2] = p;
            }
            /* flattenO: parents = the features of the state the observation is made IN */
            for (int o = 0; o < m->O; ++o)
            {
                features_of(o, m->feat_o, m->FO, of);
                float p = 1.0f;
                for (int q = 0; q < m->FO; ++q)
                    p *= expected_mult_at(counts + o_off[q] + (int64_t)parent_config(m, o_par[a * m->FO + q], x) * m->feat_o[q],
                                          m->feat_o[q], of[q]);
                O[((int64_t)a * m->S + s) * m->O + o] = p;
            }
        }
    }
}

int orc_state_history_msg(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, const float* counts,
                          const float* state_prior, int n_episodes, const int32_t* episode_len,
                          const int32_t* actions, const int32_t* observations, orc_rng* g, int32_t* states_out)
{
    /* msgSampleStateHistory (MHwithinGibbs.cpp:96-213): per episode a backward pass of normalised
     * messages p(o_t.. | s_t) in double over the float tables, then forward sampling of s_0 .. s_T */
    int const S = m->S, A = m->A;
    float* T = (float*)malloc(sizeof(float) * (size_t)S * A * S);
    float* O = (float*)malloc(sizeof(float) * (size_t)A * S * m->O);
    orc_flatten_model(m, t_par, o_par, counts, T, O);
    int max_len = 0;
    for (int e = 0; e < n_episodes; ++e) max_len = episode_len[e] > max_len ? episode_len[e] : max_len;
    double* msg   = (double*)malloc(sizeof(double) * (size_t)(max_len + 1) * S);
    double* probs = (double*)malloc(sizeof(double) * (size_t)S);
    int64_t first = 0, pos = 0;
    for (int e = 0; e < n_episodes; ++e)
    {
        int const len      = episode_len[e];
        const int32_t* act = actions + first;
        const int32_t* obs = observations + first;
        for (int s = 0; s < S; ++s) msg[(int64_t)len * S + s] = O[((int64_t)act[len - 1] * S + s) * m->O + obs[len - 1]];
        for (int step = len - 1; step >= 0; --step)
        {
            int const a = act[step];
            double tot  = 0;
            for (int s = 0; s < S; ++s)
            {
                const float* row = T + ((int64_t)s * A + a) * S;
                double acc       = 0.0; /* std::inner_product(row, row + S, message[step + 1], 0.0) */
                for (int s2 = 0; s2 < S; ++s2) acc = acc + row[s2] * msg[(int64_t)(step + 1) * S + s2];
                if (step != 0) acc *= O[((int64_t)act[step - 1] * S + s) * m->O + obs[step - 1]];
                else
                    acc *= state_prior[s];
                msg[(int64_t)step * S + s] = acc;
                tot += acc;
            }
            for (int s = 0; s < S; ++s) msg[(int64_t)step * S + s] = msg[(int64_t)step * S + s] / tot;
        }
        int state          = sample_from_mult_d(msg, S, 1, g);
        states_out[pos++]  = state;
        for (int step = 0; step < len; ++step)
        {
            double tot = 0;
            for (int s2 = 0; s2 < S; ++s2)
            {
                probs[s2] = T[((int64_t)state * A + act[step]) * S + s2] * msg[(int64_t)(step + 1) * S + s2];
                tot += probs[s2];
            }
            state             = sample_from_mult_d(probs, S, tot, g);
            states_out[pos++] = state;
        }
        first += len;
    }
    free(T), free(O), free(msg), free(probs);
    return g->overrun ? -1 : 0;
}

void orc_add_history_counts(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                            int n_episodes, const int32_t* episode_len, const int32_t* actions,
                            const int32_t* observations, const int32_t* states)
{
    /* MHwithinGibbs::computePosteriorCounts (MHwithinGibbs.cpp:397-436): incrementCountsOf(s_t, a_t, o_t, s_t+1)
     * for every step of every episode; `states` holds episode_len[e] + 1 states per episode */
    int64_t first = 0, pos = 0;
    for (int e = 0; e < n_episodes; ++e)
    {
        for (int t = 0; t < episode_len[e]; ++t)
            orc_increment_counts(m, t_par, o_par, counts, states[pos + t], actions[first + t], observations[first + t],
                                 states[pos + t + 1], 1.0f);
        first += episode_len[e];
        pos += episode_len[e] + 1;
    }
}

int64_t orc_nested_update_particle(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par, float* counts,
                                   const int32_t* states_in, int32_t* states_out, int64_t n_bottom, int a, int o,
                                   orc_rng* g, int64_t max_attempts)
{
    /* NestedBelief::updateEstimation, the body of its loop over top particles (NestedBelief.cpp:142-187):
     * rejection sampling of the particle's bottom filter on the particle's own counts, each acceptance
     * raising the counts it went through by 1 / n_bottom before the next attempt. Returns the attempts
     * (the particle's weight is multiplied by 1.0 / attempts), -1 if max_attempts / the stream ran out. */
    float const update_step = (float)(1.0 / (float)n_bottom); /* :133 */
    int64_t count = 0, accepted = 0;
    while (accepted < n_bottom)
    {
        if (count >= max_attempts || g->overrun) return -1;
        int32_t const old = states_in[orc_uniform_int(g, (uint32_t)n_bottom)]; /* belief.sample(), FlatFilter.cpp:97-102 */
        int32_t s         = old;
        int sim_o, term;
        orc_step(m, t_par, o_par, counts, &s, a, 0, g, &sim_o, &term); /* KeepCounts, :160 */
        if (sim_o == o)
        {
            states_out[accepted++] = s;
            orc_increment_counts(m, t_par, o_par, counts, old, a, o, s, update_step); /* :168 */
        }
        ++count;
    }
    return count;
}

/* ------------------------------------------------------------------------------------------ */
/* rollouts                                                                                    */
/* ------------------------------------------------------------------------------------------ */

double orc_rollout(const orc_model* m, const uint32_t* t_par, const uint32_t* o_par,
                   const float* counts, int start_state, int depth, double discount, orc_rng* g)
{
    /* RBAPOUCT::rollout (RBAPOUCT.cpp:295-323): Discount(_discount.toDouble()) starts at 1
     * (Discount.cpp:3-6); Return::add (Return.cpp:6-9); Discount::increment (Discount.cpp:8-11) */
    double ret = 0, disc = 1;
    int32_t s = start_state;
    int term  = 0;
    while (depth > 0 && !term)
    {
        int const a = random_action(m, g);
        int o;
        double const r = orc_step(m, t_par, o_par, (float*)counts, &s, a, 0, g, &o, &term);
        ret += r * disc;
        disc *= discount;
        --depth;
    }
    return ret;
}
