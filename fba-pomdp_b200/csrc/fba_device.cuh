// Device-side building blocks of the belief / rollout hot path (sm_100a).
//
// Everything here restates arithmetic of the reference (samkatt/fba-pomdp) that replay parity
// depends on; each function cites the file:line it mirrors. The translation unit is compiled with
// -fmad=false and the mixed-precision steps use explicit round-to-nearest intrinsics: the reference
// runs on baseline x86-64 (no FMA), so every multiply and add is separately rounded there.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fba_pomdp_b200.h"

namespace fba {

struct Node
{
    uint32_t par; // parent bitmask over state features
    int32_t off;  // offset of this node's CPT inside the particle's count block (floats)
};

// Model description as the kernels see it (passed by value as a kernel parameter).
struct DevModel
{
    int S, A, O, FS, FO, J; // J = FS + FO nodes per action
    int feat_s[FBA_MAX_FEATURES], feat_o[FBA_MAX_FEATURES];
    int step_s[FBA_MAX_FEATURES], step_o[FBA_MAX_FEATURES]; // indexing::stepSize (index.cpp:18-49)
    // every feature size a power of two (sysadmin, factored tiger): indices decode with shifts
    int pow2_s, pow2_o;
    int shift_s[FBA_MAX_FEATURES], shift_o[FBA_MAX_FEATURES]; // log2(step)
    int tabular, domain, action_draw;
    int sampled; // 1: --dirichlet_sampling_method regular (sample the multinomial from the Dirichlet)
    int dom_ip[32];
    double dom_dp[8];
    const double* rew_sa;
    const double* rew_as2;
    const uint8_t* term_sa;
    const uint8_t* term_as2;
    int start_kind;
    int start_ip[4];
    const float* start_values;
    double start_total;
    const int* start_table;
    const Node* nodes;      // [max_structs][A][J]
    const int* struct_size; // [max_structs] floats
};

// ------------------------------------------------------------------------------------------------
// random sources
// ------------------------------------------------------------------------------------------------

// uniform_real_distribution<double>(0,1) over two 32-bit words = generate_canonical<double,53>
// (libstdc++ bits/random.tcc:3349-3381): (w0 + w1*2^32) / 2^64, clamped below 1.
__host__ __device__ inline double canonical_from_words(uint32_t w0, uint32_t w1)
{
#ifdef __CUDA_ARCH__
    double sum = __dadd_rn((double)w0, __dmul_rn((double)w1, 4294967296.0));
    double ret = __dmul_rn(sum, 5.421010862427522170037264004349708557128906250e-20); // 2^-64, exact
#else
    double sum = (double)w0 + (double)w1 * 4294967296.0;
    double ret = sum * 5.421010862427522170037264004349708557128906250e-20;
#endif
    if (ret >= 1.0) ret = 0.99999999999999988897769753748434595763683319091796875; // nextafter(1,0)
    return ret;
}

// REPLAY: a cursor into the reference's mt19937 word stream (device copy).
struct ReplayRng
{
    const uint32_t* w;
    long long pos, end;
    int overrun;

    __host__ __device__ ReplayRng(const uint32_t* words, long long p, long long e) :
            w(words), pos(p), end(e), overrun(0)
    {
    }
    __host__ __device__ uint32_t next()
    {
        if (pos >= end)
        {
            overrun = 1;
            ++pos;
            return 0u;
        }
        return w[pos++];
    }
};

// PHILOX: Philox4x32-10 (Salmon et al., SC'11), key = seed, counter = (stream id, op offset, block).
struct PhiloxRng
{
    uint32_t k0, k1;
    uint32_t c0, c1, c2, c3;
    uint32_t b0, b1, b2, b3; // buffered words, handed out b0 first (registers: no indexed array)
    int have;
    int overrun; // never set; keeps the two sources interchangeable

    __host__ __device__ PhiloxRng(uint64_t seed, uint64_t stream, uint64_t offset) :
            k0((uint32_t)seed), k1((uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32)), c0(0),
            c1((uint32_t)stream), c2((uint32_t)(stream >> 32)), c3((uint32_t)offset), have(0),
            overrun(0)
    {
    }
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
    {
        unsigned long long p = (unsigned long long)a * b;
        hi                   = (uint32_t)(p >> 32);
        lo                   = (uint32_t)p;
    }
    __host__ __device__ void refill()
    {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r)
        {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, x0, hi0, lo0);
            mulhilo(0xCD9E8D57u, x2, hi1, lo1);
            uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
            x0 = y0, x1 = y1, x2 = y2, x3 = y3;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        b0 = x3, b1 = x2, b2 = x1, b3 = x0;
        ++c0;
        have = 4;
    }
    __host__ __device__ uint32_t next()
    {
        if (!have) refill();
        uint32_t const r = b0;
        b0 = b1, b1 = b2, b2 = b3;
        --have;
        return r;
    }
};

// the three distributions the reference draws with (src/utils/random.cpp:90-115), over either source
template<class R>
__host__ __device__ inline double draw_u(R& g) // rnd::uniform_rand01, random.cpp:100-103
{
    uint32_t w0 = g.next();
    uint32_t w1 = g.next();
    return canonical_from_words(w0, w1);
}

template<class R>
__host__ __device__ inline bool draw_b(R& g) // rnd::boolean, random.cpp:90-93 (bernoulli(0.5))
{
    return draw_u(g) < 0.5;
}

// uniform_int_distribution<int>(0, range-1): Lemire's method as libstdc++ implements it
// (bits/uniform_int_dist.h:257-281) — 1 word, rarely more
template<class R>
__host__ __device__ inline int draw_k(R& g, uint32_t range)
{
    unsigned long long product = (unsigned long long)g.next() * range;
    uint32_t low               = (uint32_t)product;
    if (low < range)
    {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold && !g.overrun)
        {
            product = (unsigned long long)g.next() * range;
            low     = (uint32_t)product;
        }
    }
    return (int)(product >> 32);
}

template<class R>
__host__ __device__ inline int draw_slow_int(R& g, int max) // rnd::slowRandomInt, random.cpp:111-115
{
#ifdef __CUDA_ARCH__
    return (int)floor(__dmul_rn(draw_u(g), (double)max));
#else
    return (int)floor(draw_u(g) * (double)max);
#endif
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Dirichlet rows, expected mode
// ------------------------------------------------------------------------------------------------

// sampleFromExpectedMult (random.cpp:244-255) -> sampleFromMult<float const> (random.hpp:93-115):
// DOUBLE total of the float counts, FLOAT running prefix compared against a double threshold.
// LONG = false: plain loop (factored models: feature ranges of 2..5).
// LONG = true: rows of tabular models (S or O values): 16 independent loads in flight per chunk, so a
// row costs ceil(n/16) memory latencies instead of n/4; the sums stay strictly sequential. Kernels
// are instantiated for both so that the short-row variants keep their small register footprint.
template<bool LONG>
__device__ __forceinline__ int sample_expected_mult(const float* row, int n, double u)
{
    if (!LONG || n <= 4)
    {
        double total = (double)row[0];
        for (int i = 1; i < n; ++i) total = __dadd_rn(total, (double)row[i]);
        double const p = __dmul_rn(u, total);
        float sum      = row[0];
        for (int i = 1; i < n; ++i)
        {
            if (p < (double)sum) return i - 1;
            sum = __fadd_rn(sum, row[i]);
        }
        return n - 1;
    }
    constexpr int CH = 16;
    double total = 0.0;
    for (int base = 0; base < n; base += CH)
    {
        float v[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) v[k] = (base + k < n) ? row[base + k] : 0.0f;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (base + k < n) total = (base + k == 0) ? (double)v[k] : __dadd_rn(total, (double)v[k]);
    }
    double const p = __dmul_rn(u, total);
    float sum      = 0.0f;
    for (int base = 0; base < n; base += CH)
    {
        float v[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) v[k] = (base + k < n) ? row[base + k] : 0.0f; // L1 hits
#pragma unroll
        for (int k = 0; k < CH; ++k)
        {
            int const i = base + k;
            if (i >= n) break;
            if (i == 0)
            {
                sum = v[k];
                continue;
            }
            if (p < (double)sum) return i - 1;
            sum = __fadd_rn(sum, v[k]);
        }
    }
    return n - 1;
}

// The same draw with the row loaded cooperatively by a warp: lane l holds row[base + l] of each
// 32-element chunk (one coalesced request instead of n dependent ones); every lane then replays the
// reference's SEQUENTIAL sums through shuffles, so all lanes return the same, bit-identical index.
// Must be called by all 32 lanes with identical arguments.
__device__ __forceinline__ int sample_expected_mult_warp(const float* row, int n, double u)
{
    int const lane = threadIdx.x & 31;
    double total   = 0.0;
    for (int base = 0; base < n; base += 32)
    {
        float const v = (base + lane < n) ? row[base + lane] : 0.0f;
        int const m   = min(32, n - base);
        for (int k = 0; k < m; ++k)
        {
            double const x = (double)__shfl_sync(0xffffffffu, v, k);
            total          = (base + k == 0) ? x : __dadd_rn(total, x);
        }
    }
    double const p = __dmul_rn(u, total);
    float sum      = 0.0f;
    int result     = n - 1;
    bool found     = false;
    for (int base = 0; base < n && !found; base += 32)
    {
        float const v = (base + lane < n) ? row[base + lane] : 0.0f; // second pass hits L1
        int const m   = min(32, n - base);
        for (int k = 0; k < m; ++k)
        {
            float const x = __shfl_sync(0xffffffffu, v, k);
            int const i   = base + k;
            if (i == 0)
            {
                sum = x;
                continue;
            }
            if (!found && p < (double)sum)
            {
                result = i - 1;
                found  = true;
            }
            sum = __fadd_rn(sum, x);
        }
    }
    return result;
}

// expectedMult(dir, n)[k] (random.cpp:257-279): FLOAT sum and FLOAT divide
__device__ __forceinline__ float expected_mult_at(const float* row, int n, int k)
{
    float sum = row[0];
    for (int i = 1; i < n; ++i) sum = __fadd_rn(sum, row[i]);
    if ((double)sum <= 1e-300) return 0.0f;
    return __fdiv_rn(row[k], sum);
}

// ------------------------------------------------------------------------------------------------
// Dirichlet rows, SAMPLED mode (--dirichlet_sampling_method regular):
// the multinomial is itself drawn from the Dirichlet — p ~ Dir(counts), via one Gamma(count_i, 1)
// per cell — before the categorical draw (sampleFromSampledMult, random.cpp:217-242) or before the
// likelihood is read off (sampleMult, random.cpp:281-304). The reference's gamma sampler is
// Marsaglia–Tsang fed by a 128-strip ziggurat normal (random.cpp:146-213) on libm's log/exp/pow;
// device libm differs in the last bits, so this mode has STATISTICAL parity only (PHILOX mode):
// the same Marsaglia–Tsang construction with a Box–Muller normal.
// ------------------------------------------------------------------------------------------------
template<class R>
__device__ __forceinline__ double draw_normal(R& g)
{
    double const u1 = fmax(draw_u(g), 1e-300), u2 = draw_u(g);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

template<class R>
__device__ double draw_gamma(R& g, double shape) // rnd::sample::gamma, random.cpp:189-213
{
    if (shape <= 0.0) return 0.0; // gamma(shape + 1) * pow(u, 1 / 0) = 0
    double boost = 1.0;
    if (shape < 1.0)
    {
        boost = pow(draw_u(g), 1.0 / shape);
        shape += 1.0;
    }
    double const d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int guard = 0; guard < 1000; ++guard)
    {
        double x, v;
        do {
            x = draw_normal(g);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v               = v * v * v;
        double const u  = draw_u(g), x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;
}

// category ~ Gamma(row_i) / sum: one pass, no storage (weighted reservoir: cell i replaces the
// current pick with probability g_i / (g_0 + … + g_i))
template<class R>
__device__ int sample_sampled_mult(const float* row, int n, R& g)
{
    double sum = 0.0;
    int pick   = n - 1;
    for (int i = 0; i < n; ++i)
    {
        double const gi = draw_gamma(g, (double)row[i]);
        sum += gi;
        if (gi > 0.0 && draw_u(g) * sum < gi) pick = i;
    }
    return pick;
}

// sampleMult(row, n)[k] (random.cpp:281-304): float gammas, float sum, float divide
template<class R>
__device__ float sampled_mult_at(const float* row, int n, int k, R& g)
{
    float sum = 0.0f, gk = 0.0f;
    for (int i = 0; i < n; ++i)
    {
        float const gi = (float)draw_gamma(g, (double)row[i]);
        sum += gi;
        if (i == k) gk = gi;
    }
    return (sum > 0.0f) ? gk / sum : 0.0f;
}

// the categorical draw / the likelihood factor in whichever mode the model is in
// (SAMPLED is a template parameter, not a run-time branch: the expected-mode kernels keep their
// register footprint — k_propose stays at 40 registers without spills)
template<bool LONG, bool SAMPLED, class R>
__device__ __forceinline__ int sample_row(const float* row, int n, R& g)
{
    if (SAMPLED) return sample_sampled_mult(row, n, g);
    return sample_expected_mult<LONG>(row, n, draw_u(g));
}
template<bool SAMPLED, class R>
__device__ __forceinline__ float likelihood_at(const float* row, int n, int k, R& g)
{
    if (SAMPLED) return sampled_mult_at(row, n, k, g);
    return expected_mult_at(row, n, k);
}

// ------------------------------------------------------------------------------------------------
// feature vectors: up to 16 features of <= 256 values packed in two 64-bit words, so that no
// per-thread array (and no local memory) is needed. A single feature (tabular) keeps its full value.
// ------------------------------------------------------------------------------------------------
struct Feat
{
    unsigned long long lo, hi;
    __device__ __forceinline__ int get(int f, bool single) const
    {
        if (single) return (int)lo;
        return (int)(((f < 8) ? (lo >> (8 * f)) : (hi >> (8 * (f - 8)))) & 0xFFull);
    }
    __device__ __forceinline__ void set(int f, int v, bool single)
    {
        if (single)
            lo = (unsigned long long)(unsigned)v;
        else if (f < 8)
            lo |= (unsigned long long)v << (8 * f);
        else
            hi |= (unsigned long long)v << (8 * (f - 8));
    }
};

// indexing::projectUsingStepSize (index.cpp:98-119): feature 0 most significant
__device__ __forceinline__ Feat decode(int v, const int* step, int n, int pow2 = 0, const int* shift = nullptr)
{
    Feat x{0ull, 0ull};
    if (n == 1)
    {
        x.lo = (unsigned long long)(unsigned)v;
        return x;
    }
    if (pow2)
    { // no integer divisions: feature f = bits [shift_f, shift_{f-1}) of the index
        int hi = 31;
        for (int f = 0; f < n; ++f)
        {
            int const sh = shift[f];
            x.set(f, (int)(((unsigned)v & ((2u << hi) - 1u)) >> sh), false);
            hi = sh - 1;
        }
        return x;
    }
    for (int f = 0; f < n; ++f)
    {
        int const q = v / step[f];
        v -= q * step[f];
        x.set(f, q, false);
    }
    return x;
}

// DBNNode::cptIndex (DBNNode.cpp:171-205): mixed radix over the node's parents, ascending features
__device__ __forceinline__ int parent_config(const DevModel& M, uint32_t par, const Feat& x)
{
    bool const single = (M.FS == 1);
    int cfg           = 0;
    while (par)
    {
        int const f = __ffs(par) - 1;
        par &= par - 1;
        cfg = cfg * M.feat_s[f] + x.get(f, single);
    }
    return cfg;
}

// ------------------------------------------------------------------------------------------------
// domain functors: BADomainExtension::{reward, terminal}
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double domain_reward(const DevModel& M, int s, int a, int s2, bool& terminal)
{
    switch (M.domain)
    {
        case FBA_DOM_TIGER: // TigerBAExtension.cpp:21-44 (OBSERVE = 2, Tiger.hpp:30)
            terminal = M.dom_ip[0] && a != 2;
            if (a == 2) return -1.0;
            return (a == s) ? 10.0 : -100.0;
        case FBA_DOM_FACTORED_TIGER: { // FactoredTigerBAExtension.cpp:27-56
            terminal = M.dom_ip[0] && a != 2;
            if (a == 2) return -1.0;
            int const loc = (s < M.S / 2) ? 0 : 1;
            return (a == loc) ? 10.0 : -100.0;
        }
        case FBA_DOM_SYSADMIN: { // SysAdminBAExtension.cpp:27-48 (float arithmetic, exact integers)
            terminal        = false;
            float const up  = (float)__popc((unsigned)s2);
            float const reb = (a >= M.dom_ip[0]) ? 1.0f : 0.0f;
            return (double)__fsub_rn(up, __fmul_rn((float)M.dom_dp[0], reb));
        }
        case FBA_DOM_GRIDWORLD: { // GridWorldBAExtension.cpp:74-100: a function of s alone
            int const size = M.dom_ip[0], G = M.dom_ip[1];
            int const g = s % G, y = (s / G) % size, x = s / (G * size);
            bool const at_goal = (x == M.dom_ip[2 + 2 * g]) && (y == M.dom_ip[3 + 2 * g]);
            terminal           = at_goal;
            return at_goal ? M.dom_dp[0] : M.dom_dp[1];
        }
        case FBA_DOM_COLLISION_AVOIDANCE: { // CollisionAvoidanceBAExtension.cpp:59-89
            int const H = M.dom_ip[1], nobs = M.dom_ip[2];
            int obst_space = 1;
            for (int i = 0; i < nobs; ++i) obst_space *= H;
            int const x = s2 / (H * obst_space), y = (s2 / obst_space) % H;
            bool crashed = false;
            if (x < nobs)
            {
                int pos = s2 % obst_space;
                for (int i = nobs - 1; i > x; --i) pos /= H;
                crashed = (y == pos % H);
            }
            terminal = crashed || x == 0;
            if (crashed) return -M.dom_dp[1];
            return (a == 1) ? 0.0 : -M.dom_dp[0]; // STAY = 1 (CollisionAvoidance.hpp:91)
        }
        default: { // FBA_DOM_TABLE
            double r = 0.0;
            int t    = 0;
            if (M.rew_sa) r += M.rew_sa[(long long)s * M.A + a];
            if (M.rew_as2) r += M.rew_as2[(long long)a * M.S + s2];
            if (M.term_sa) t |= M.term_sa[(long long)s * M.A + a];
            if (M.term_as2) t |= M.term_as2[(long long)a * M.S + s2];
            terminal = t != 0;
            return r;
        }
    }
}

// the domain's sampleStartState draws (SURVEY.md §8 a')
template<class R>
__device__ __forceinline__ int sample_start_state(const DevModel& M, R& g)
{
    switch (M.start_kind)
    {
        case FBA_START_CONST: return M.start_ip[0]; // SysAdmin.cpp:102-105
        case FBA_START_BOOL: return draw_b(g) ? M.start_ip[0] : M.start_ip[1]; // Tiger.cpp:16-19
        case FBA_START_UNIFORM_INT: return draw_k(g, (uint32_t)M.start_ip[0]); // FactoredTiger.cpp:71-75
        case FBA_START_SLOW2: { // GridWorld.cpp:260-270
            int const i = draw_slow_int(g, M.start_ip[0]);
            int const j = draw_slow_int(g, M.start_ip[1]);
            return M.start_table[i * M.start_ip[1] + j];
        }
        default: { // categoricalDistr::sample, distributions.cpp:47-51: sampleFromMult<float const>
            double const p = __dmul_rn(draw_u(g), M.start_total);
            int const n    = M.start_ip[0];
            float sum      = M.start_values[0];
            for (int i = 1; i < n; ++i)
            {
                if (p < (double)sum) return i - 1;
                sum = __fadd_rn(sum, M.start_values[i]);
            }
            return n - 1;
        }
    }
}

template<class R>
__device__ __forceinline__ int random_action(const DevModel& M, R& g)
{
    if (M.action_draw == FBA_ACT_SLOW_INT) return draw_slow_int(g, M.A); // GridWorld.cpp:223
    return draw_k(g, (uint32_t)M.A); // Tiger.cpp:24, SysAdmin.cpp:170, ...
}

// ------------------------------------------------------------------------------------------------
// BAPOMDP::step (BAPOMDP.cpp:111-143) on one particle
// ------------------------------------------------------------------------------------------------
enum StepMode {
    STEP_UPDATE = 0, // UpdateCounts: +1 written into the count block
    STEP_KEEP   = 1, // KeepCounts: counts read-only (rollouts, planning)
    STEP_RECORD = 2  // counts read-only, the J incremented cell offsets are recorded (rejection
                     // sampling applies them to the accepted copy)
};

// nodes: this particle's structure, action a: J entries. counts: the particle's block.
// Returns s'; o_out = simulated observation. x_new returns the new state's features.
// COOP: the 32 lanes of a warp run ONE step together (identical arguments and random source in
// every lane) and load each row cooperatively — for latency-bound small batches of rollouts.
template<int MODE, class R, bool COOP = false, bool LONG = false, bool SAMPLED = false>
__device__ __forceinline__ int hyper_step(const DevModel& M, const Node* __restrict__ nodes,
                                          float* counts, int s, R& g, int& o_out, Feat& x_new,
                                          int* rec)
{
    bool const single_s = (M.FS == 1), single_o = (M.FO == 1);
    Feat const x = decode(s, M.step_s, M.FS, M.pow2_s, M.shift_s);

    // sampleStateIndex (BAFlatModel.cpp:83-91 / BABNModel.cpp:292-307): one draw per state
    // feature in feature order, every node conditioned on the OLD state
    Feat x2{0ull, 0ull};
    int s2 = 0;
    for (int f = 0; f < M.FS; ++f)
    {
        Node const nd   = nodes[f];
        int const range = M.feat_s[f];
        int const cell  = nd.off + parent_config(M, nd.par, x) * range;
        int v;
        if (MODE == STEP_UPDATE && !COOP && !LONG && !SAMPLED && range == 2)
        { // binary feature, counts updated: ONE 16-byte load and ONE 16-byte store of the aligned chunk
          // that holds the row (and a neighbouring row of the same particle) instead of an 8-byte load
          // and a 4-byte store — the write-back of isolated dirty sectors is what bounds this kernel,
          // and 16-byte stores are the cheapest form of it (tools/exp_granule.cu). Same arithmetic as
          // sample_expected_mult(row, 2, u).
            float4* const chunk = reinterpret_cast<float4*>(counts + (cell & ~3));
            float4 ch           = *chunk;
            bool const upper    = (cell & 2) != 0;
            float const r0 = upper ? ch.z : ch.x, r1 = upper ? ch.w : ch.y;
            double const p = __dmul_rn(draw_u(g), __dadd_rn((double)r0, (double)r1));
            v              = (p < (double)r0) ? 0 : 1;
            float const inc = __fadd_rn(v ? r1 : r0, 1.0f);
            if (upper) (v ? ch.w : ch.z) = inc;
            else
                (v ? ch.y : ch.x) = inc;
            *chunk = ch;
        } else
        {
            v = COOP ? sample_expected_mult_warp(counts + cell, range, draw_u(g))
                     : sample_row<LONG, SAMPLED>(counts + cell, range, g);
            if (MODE == STEP_UPDATE) counts[cell + v] = __fadd_rn(counts[cell + v], 1.0f);
        }
        x2.set(f, v, single_s);
        s2 += v * M.step_s[f];
        if (MODE == STEP_RECORD) rec[f] = cell + v;
    }
    if (single_s) s2 = (int)x2.lo;

    // sampleObservationIndex (BAFlatModel.cpp:93-103 / BABNModel.cpp:309-326): parents = NEW state.
    // NOTE the transition increments above touch only transition CPTs, so doing them before the
    // observation draws (instead of after, BAPOMDP.cpp:134-137) changes nothing.
    int o = 0;
    Feat of{0ull, 0ull};
    for (int q = 0; q < M.FO; ++q)
    {
        Node const nd   = nodes[M.FS + q];
        int const range = M.feat_o[q];
        int const cell  = nd.off + parent_config(M, nd.par, x2) * range;
        int const v     = COOP ? sample_expected_mult_warp(counts + cell, range, draw_u(g))
                               : sample_row<LONG, SAMPLED>(counts + cell, range, g);
        of.set(q, v, single_o);
        o += v * M.step_o[q];
    }
    if (single_o) o = (int)of.lo;

    if (MODE != STEP_KEEP)
    {
        // incrementCountsOf, observation part. Tabular: psi[a][s'][o] (BAFlatModel.cpp:126-141).
        // Factored: the observation CPTs are indexed with the OLD state's features
        // (BABNModel.cpp:366,380) — a reference quirk that replay reproduces.
        Feat const& xo = M.tabular ? x2 : x;
        for (int q = 0; q < M.FO; ++q)
        {
            Node const nd   = nodes[M.FS + q];
            int const range = M.feat_o[q];
            int const cell  = nd.off + parent_config(M, nd.par, xo) * range + of.get(q, single_o);
            if (MODE == STEP_UPDATE)
            {
                if (!COOP && !LONG && !SAMPLED && range == 2)
                { // as above: the aligned 16-byte chunk around the cell, read and written whole
                    float4* const chunk = reinterpret_cast<float4*>(counts + (cell & ~3));
                    float4 ch           = *chunk;
                    int const k         = cell & 3;
                    float& t            = (k == 0) ? ch.x : (k == 1) ? ch.y : (k == 2) ? ch.z : ch.w;
                    t                   = __fadd_rn(t, 1.0f);
                    *chunk              = ch;
                } else
                    counts[cell] = __fadd_rn(counts[cell], 1.0f);
            }
            if (MODE == STEP_RECORD) rec[M.FS + q] = cell;
        }
    }
    o_out = o;
    x_new = x2;
    return s2;
}

// BA{Flat,BN}Model::computeObservationProbability, expected mode
// (BAFlatModel.cpp:105-124, BABNModel.cpp:328-352), for the state whose features are x
template<bool SAMPLED, class R>
__device__ __forceinline__ double obs_probability(const DevModel& M, const Node* __restrict__ nodes,
                                                  const float* counts, const Feat& x, int o, R& g)
{
    if (M.tabular)
    {
        if (M.O == 1) return 1.0;
        Node const nd = nodes[M.FS];
        return (double)likelihood_at<SAMPLED>(counts + nd.off + (int)x.lo * M.O, M.O, o, g);
    }
    Feat const of       = decode(o, M.step_o, M.FO, M.pow2_o, M.shift_o);
    bool const single_o = (M.FO == 1);
    double prob         = 1.0;
    for (int q = 0; q < M.FO; ++q)
    {
        Node const nd   = nodes[M.FS + q];
        int const range = M.feat_o[q];
        int const cell  = nd.off + parent_config(M, nd.par, x) * range;
        prob = __dmul_rn(prob, (double)likelihood_at<SAMPLED>(counts + cell, range, of.get(q, single_o), g));
    }
    return prob;
}
// ------------------------------------------------------------------------------------------------
// Tabular particles stored as SHARED BASE + PRIVATE DELTA (SURVEY.md §7 hard part 4): the dense
// phi/psi tables of a large tabular model (gridworld 5: 720 KB, gridworld 7: 7.7 MB per particle)
// cannot be replicated per particle. The reference survives through copy-on-write rows
// (BAFlatModel.cpp:185-252,284-354); here a particle owns only the list of cells it has incremented:
//   block[0] = n, block[1..n] = cell index of each +1, in time order
// and reads a row as base row + its own increments. Every increment is exactly +1.0f
// (BAFlatModel.cpp:126-141), so applying them one by one reproduces the reference's float adds
// bit for bit whatever the order.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxDeltaRow = 512; // longest row (S or O) a delta model may have

__device__ __forceinline__ void delta_row(const float* __restrict__ base, const int* block, int cell0, int n,
                                          float* buf)
{
    for (int i = 0; i < n; ++i) buf[i] = base[cell0 + i];
    int const ne = block[0];
    for (int e = 1; e <= ne; ++e)
    {
        unsigned const c = (unsigned)(block[e] - cell0);
        if (c < (unsigned)n) buf[c] = __fadd_rn(buf[c], 1.0f);
    }
}

// BAPOMDP::step on a delta particle (tabular: one transition node and one observation node per
// action). overflow is set when the increment list is full (the increments are then dropped).
template<int MODE, bool SAMPLED, class R>
__device__ __forceinline__ int hyper_step_delta(const DevModel& M, const Node* __restrict__ nodes,
                                                const float* __restrict__ base, int* block, int cap, int s,
                                                R& g, int& o_out, int* rec, int* overflow)
{
    float buf[kMaxDeltaRow];
    int const cell_t = nodes[0].off + s * M.S; // phi[s][a][.]
    delta_row(base, block, cell_t, M.S, buf);
    int const s2     = sample_row<true, SAMPLED>(buf, M.S, g);
    int const cell_o = nodes[1].off + s2 * M.O; // psi[a][s'][.]
    delta_row(base, block, cell_o, M.O, buf);
    int const o = sample_row<true, SAMPLED>(buf, M.O, g);
    if (MODE == STEP_UPDATE)
    { // incrementCountsOf(s, a, o, s') (BAFlatModel.cpp:126-141)
        int const ne = block[0];
        if (ne + 2 <= cap)
        {
            block[1 + ne] = cell_t + s2;
            block[2 + ne] = cell_o + o;
            block[0]      = ne + 2;
        } else
            *overflow = 2;
    }
    if (MODE == STEP_RECORD)
    {
        rec[0] = cell_t + s2;
        rec[1] = cell_o + o;
    }
    o_out = o;
    return s2;
}

template<bool SAMPLED, class R>
__device__ __forceinline__ double obs_probability_delta(const DevModel& M, const Node* __restrict__ nodes,
                                                        const float* __restrict__ base, const int* block,
                                                        int s2, int o, R& g)
{
    if (M.O == 1) return 1.0; // BAFlatModel.cpp:117-120
    float buf[kMaxDeltaRow];
    delta_row(base, block, nodes[1].off + s2 * M.O, M.O, buf);
    return (double)likelihood_at<SAMPLED>(buf, M.O, o, g);
}

// ------------------------------------------------------------------------------------------------
// FACTORED particles stored as SHARED BASE + PRIVATE JOURNAL: every particle of a
// factored belief descends from one of a few prior prototypes (sysadmin: ONE, SysAdminFactoredPrior.cpp:
// 34-37,138-267), and one update increments exactly J = FS + FO cells of it — so a young particle is its
// prototype plus a short list of cells, J per update, in node order. Layout (4-byte words, everything
// 16-byte aligned so that an update is read and written as whole int4 vectors):
//   block[0] = 3 + nu Jp   (words that follow the first; the copy kernels move (block[0] + 1 + 3) / 4 vectors)
//   block[4 + u Jp + j]    = the cell node j's +1 of update u went to, j < J; -1 for j in [J, Jp - 1);
//   block[4 + u Jp + Jp-1] = -2 - (the update's action)
// with Jp = J + 1 rounded up to a multiple of 4. Against the dense private block (11.8 KB per sysadmin
// particle) an update then appends 48 CONTIGUOUS bytes instead of dirtying 11 isolated 32-byte sectors,
// and a resampling copy moves 16 + 48 t bytes instead of 23.7 KB. A row is the prototype's row plus the
// particle's own +1s; every increment is exactly +1.0f (DBNNode.cpp:124-127), applied one by one, so sums
// are bit-identical to the dense block's (tests/test_cuda_journal.py replays the same fixtures).
// All FS transition rows of a step depend only on the OLD state, so ONE pass over the journal fills them
// all (entry j of an update can only belong to node j); the observation rows need the new state.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxJournalCells = 96; // sum of the feature ranges the gathered rows may have
constexpr int kJournalHeader   = 4;  // words before the first update

// J entries + at least one tag slot, rounded up to whole int4 vectors; the LAST slot of an update holds
// -2 - action, so that a reader skips the updates of other actions after one comparison
__device__ __forceinline__ int journal_padded(int J) { return (J + 1 + 3) & ~3; }
__device__ __forceinline__ int journal_tag(int a) { return -2 - a; }

// likelihood != nullptr: also returns P(o_real | a, s') evaluated AFTER the increment, as
// importance_sampling::update does (ImportanceSampler.hpp:45-46): the observation row of the new state
// is the row the simulated observation was drawn from, plus the +1 just made if it landed in that row
template<int MODE, bool SAMPLED, class R>
__device__ __forceinline__ int hyper_step_journal(const DevModel& M, const Node* __restrict__ nodes,
                                                  const float* __restrict__ base, int* block, int cap, int s,
                                                  R& g, int& o_out, int* rec, int* overflow, int o_real,
                                                  double* likelihood, int a)
{
    bool const single_s = (M.FS == 1), single_o = (M.FO == 1);
    int const J = M.J, Jp = journal_padded(J), nu = (block[0] - (kJournalHeader - 1)) / Jp;
    const int* upd = block + kJournalHeader;
    Feat const x = decode(s, M.step_s, M.FS, M.pow2_s, M.shift_s);
    float row[kMaxJournalCells];
    int cell0[FBA_MAX_FEATURES], roff[FBA_MAX_FEATURES];
    int tot = 0;
    for (int f = 0; f < M.FS; ++f)
    {
        Node const nd   = nodes[f];
        int const range = M.feat_s[f];
        cell0[f]        = nd.off + parent_config(M, nd.par, x) * range;
        roff[f]         = tot;
        for (int v = 0; v < range; ++v) row[tot + v] = base[cell0[f] + v];
        tot += range;
    }
    int const tag = journal_tag(a);
    for (int u = 0; u < nu; ++u) // entry f of an update can only hit node f's row
        for (int f = 0; f < M.FS && upd[u * Jp + Jp - 1] == tag; ++f)
        {
            unsigned const c = (unsigned)(upd[u * Jp + f] - cell0[f]);
            if (c < (unsigned)M.feat_s[f]) row[roff[f] + c] = __fadd_rn(row[roff[f] + c], 1.0f);
        }
    Feat x2{0ull, 0ull};
    int s2 = 0;
    int inc[FBA_MAX_FEATURES * 2];
    for (int f = 0; f < M.FS; ++f)
    {
        int const v = sample_row<true, SAMPLED>(row + roff[f], M.feat_s[f], g);
        x2.set(f, v, single_s);
        s2 += v * M.step_s[f];
        inc[f] = cell0[f] + v;
    }
    if (single_s) s2 = (int)x2.lo;

    // observation rows of the NEW state (BABNModel.cpp:309-326)
    tot = 0;
    for (int q = 0; q < M.FO; ++q)
    {
        Node const nd   = nodes[M.FS + q];
        int const range = M.feat_o[q];
        cell0[q]        = nd.off + parent_config(M, nd.par, x2) * range;
        roff[q]         = tot;
        for (int v = 0; v < range; ++v) row[tot + v] = base[cell0[q] + v];
        tot += range;
    }
    for (int u = 0; u < nu; ++u)
        for (int q = 0; q < M.FO && upd[u * Jp + Jp - 1] == tag; ++q)
        {
            unsigned const c = (unsigned)(upd[u * Jp + M.FS + q] - cell0[q]);
            if (c < (unsigned)M.feat_o[q]) row[roff[q] + c] = __fadd_rn(row[roff[q] + c], 1.0f);
        }
    int o = 0;
    Feat of{0ull, 0ull};
    for (int q = 0; q < M.FO; ++q)
    {
        int const v = sample_row<true, SAMPLED>(row + roff[q], M.feat_o[q], g);
        of.set(q, v, single_o);
        o += v * M.step_o[q];
    }
    if (single_o) o = (int)of.lo;
    // incrementCountsOf, observation part: the OLD state's parent configuration (BABNModel.cpp:366,380)
    for (int q = 0; q < M.FO; ++q)
    {
        Node const nd = nodes[M.FS + q];
        inc[M.FS + q] = nd.off + parent_config(M, nd.par, x) * M.feat_o[q] + of.get(q, single_o);
    }
    if (MODE == STEP_UPDATE)
    {
        if ((nu + 1) * J <= cap)
        {
            int* dst = block + kJournalHeader + nu * Jp;
            for (int j = 0; j < Jp; ++j) dst[j] = (j < J) ? inc[j] : (j == Jp - 1) ? tag : -1;
            block[0] = (kJournalHeader - 1) + (nu + 1) * Jp;
        } else
            *overflow = 2;
    }
    if (MODE == STEP_RECORD)
        for (int j = 0; j < J; ++j) rec[j] = inc[j];
    if (likelihood)
    { // BABNModel::computeObservationProbability (BABNModel.cpp:328-352) on the updated counts
        Feat const fr = decode(o_real, M.step_o, M.FO, M.pow2_o, M.shift_o);
        double prob   = 1.0;
        for (int q = 0; q < M.FO; ++q)
        {
            int const range = M.feat_o[q];
            if (MODE == STEP_UPDATE)
            {
                unsigned const c = (unsigned)(inc[M.FS + q] - cell0[q]);
                if (c < (unsigned)range) row[roff[q] + c] = __fadd_rn(row[roff[q] + c], 1.0f);
            }
            prob = __dmul_rn(prob, (double)likelihood_at<SAMPLED>(row + roff[q], range, fr.get(q, single_o), g));
        }
        *likelihood = prob;
    }
    o_out = o;
    return s2;
}

// The same step when EVERY feature is binary (sysadmin, factored tiger), expected mode, UpdateCounts,
// observation CPTs of at most 8 cells, as three phases so that the kernel can feed the journal in
// chunks STAGED THROUGH SHARED MEMORY (k_propose_journal_staged): begin() fixes the cells of interest —
// the parent configuration of every transition node is known from the old state — add() counts, in 8-bit
// fields of four 64-bit registers, how many +1s each of the 2 FS cells of interest received (f is a
// compile-time constant of the unrolled loop, so the field shifts are too) and histograms the WHOLE CPT of
// each observation node (<= 8 cells, one 64-bit register each) because their row is only known once the
// new state is drawn; finish() rebuilds the rows as base, +1.0f, +1.0f, … — the same sequence of rounded
// adds the dense block went through — draws, and returns the J cells to append. No array in local
// memory in the hot loop. At most 255 updates per particle.
struct JournalBinaryStep
{
    int cell0[FBA_MAX_FEATURES];
    int ooff0, ooff1, tag;
    unsigned long long h0a, h0b, h1a, h1b; // hits of value 0 / 1, features 0-7 / 8-15
    unsigned long long ho0, ho1;           // per-cell hits of observation node 0 / 1
    Feat x;

    __device__ __forceinline__ void begin(const DevModel& M, const Node* __restrict__ nodes, int s, int a)
    {
        tag = journal_tag(a);
        x   = decode(s, M.step_s, M.FS, M.pow2_s, M.shift_s);
#pragma unroll
        for (int f = 0; f < FBA_MAX_FEATURES; ++f)
            cell0[f] = (f < M.FS) ? nodes[f].off + parent_config(M, nodes[f].par, x) * 2 : 0x40000000;
        ooff0 = nodes[M.FS].off, ooff1 = (M.FO > 1) ? nodes[M.FS + 1].off : 0x40000000;
        h0a = h0b = h1a = h1b = ho0 = ho1 = 0ull;
    }

    // n_updates updates of nvec int4 vectors each, lying at upd[0 .. n_updates nvec)
    __device__ __forceinline__ void add(const DevModel& M, const int4* upd, int n_updates, int nvec)
    {
        for (int u = 0; u < n_updates; ++u)
        {
            if (upd[u * nvec + nvec - 1].w != tag) continue; // another action's update: none of its cells is ours
#pragma unroll
            for (int k = 0; k < (FBA_MAX_FEATURES + 4 + 3) / 4; ++k)
            {
                if (k >= nvec) break;
                int4 const v4 = upd[u * nvec + k];
#pragma unroll
                for (int l = 0; l < 4; ++l)
                {
                    int const j = 4 * k + l; // compile-time
                    int const e = (l == 0) ? v4.x : (l == 1) ? v4.y : (l == 2) ? v4.z : v4.w;
                    if (j < FBA_MAX_FEATURES && j < M.FS)
                    {
                        int const c = e - cell0[j < FBA_MAX_FEATURES ? j : 0];
                        unsigned long long const one = 1ull << (8 * (j & 7));
                        if (j < 8)
                        {
                            h0a += (c == 0) ? one : 0ull;
                            h1a += (c == 1) ? one : 0ull;
                        } else
                        {
                            h0b += (c == 0) ? one : 0ull;
                            h1b += (c == 1) ? one : 0ull;
                        }
                    } else if (j == M.FS)
                    { // cells of other actions' nodes fall outside [0, 8)
                        unsigned const d = (unsigned)(e - ooff0);
                        ho0 += (d < 8u) ? 1ull << (8 * d) : 0ull;
                    } else if (j == M.FS + 1 && j < M.J)
                    {
                        unsigned const d = (unsigned)(e - ooff1);
                        ho1 += (d < 8u) ? 1ull << (8 * d) : 0ull;
                    }
                }
            }
        }
    }

    // draws s' and the simulated observation, fills inc[0..J) with the cells to increment, returns s';
    // *likelihood = P(o_real | a, s') on the counts AFTER the increment (ImportanceSampler.hpp:45-46)
    template<class R>
    __device__ __forceinline__ int finish(const DevModel& M, const Node* __restrict__ nodes,
                                          const float* __restrict__ base, R& g, int o_real, int* inc,
                                          double* likelihood) const
    {
        Feat x2{0ull, 0ull};
        int s2 = 0;
#pragma unroll
        for (int f = 0; f < FBA_MAX_FEATURES; ++f)
        {
            if (f >= M.FS) break;
            int const k0 = (int)(((f < 8 ? h0a : h0b) >> (8 * (f & 7))) & 0xff);
            int const k1 = (int)(((f < 8 ? h1a : h1b) >> (8 * (f & 7))) & 0xff);
            float r0 = base[cell0[f]], r1 = base[cell0[f] + 1];
            for (int k = 0; k < k0; ++k) r0 = __fadd_rn(r0, 1.0f);
            for (int k = 0; k < k1; ++k) r1 = __fadd_rn(r1, 1.0f);
            // sample_expected_mult(row, 2, u): double total, float prefix
            double const p = __dmul_rn(draw_u(g), __dadd_rn((double)r0, (double)r1));
            int const v    = (p < (double)r0) ? 0 : 1;
            x2.set(f, v, M.FS == 1);
            s2 += v * M.step_s[f];
            inc[f] = cell0[f] + v;
        }
        if (M.FS == 1) s2 = (int)x2.lo;
        // observation nodes: rows of the NEW state out of the CPT histograms
        Feat const fr = decode(o_real, M.step_o, M.FO, M.pow2_o, M.shift_o);
        double prob   = 1.0;
        for (int q = 0; q < M.FO; ++q)
        {
            Node const nd = nodes[M.FS + q];
            int const r   = parent_config(M, nd.par, x2) * 2; // the row's first cell inside the CPT
            unsigned long long const h = q ? ho1 : ho0;
            int const k0 = (int)((h >> (8 * r)) & 0xff), k1 = (int)((h >> (8 * (r + 1))) & 0xff);
            float r0 = base[nd.off + r], r1 = base[nd.off + r + 1];
            for (int k = 0; k < k0; ++k) r0 = __fadd_rn(r0, 1.0f);
            for (int k = 0; k < k1; ++k) r1 = __fadd_rn(r1, 1.0f);
            double const p = __dmul_rn(draw_u(g), __dadd_rn((double)r0, (double)r1));
            int const v    = (p < (double)r0) ? 0 : 1;
            // incrementCountsOf: the OLD state's parent configuration (BABNModel.cpp:366,380)
            int const ci  = parent_config(M, nd.par, x) * 2 + v;
            inc[M.FS + q] = nd.off + ci;
            if (ci == r) r0 = __fadd_rn(r0, 1.0f);
            if (ci == r + 1) r1 = __fadd_rn(r1, 1.0f);
            // expectedMult(row)[o_real_q] on the updated row: float sum, float divide (random.cpp:257-279)
            float const sum = __fadd_rn(r0, r1);
            float const num = fr.get(q, M.FO == 1) ? r1 : r0;
            prob = __dmul_rn(prob, ((double)sum <= 1e-300) ? 0.0 : (double)__fdiv_rn(num, sum));
        }
        *likelihood = prob;
        return s2;
    }
};

// BAPOMDP::step on a base+delta particle of either kind: tabular (row scans over long rows) or factored
// (journal). proto = the particle's prior prototype (its base table); proto_sid[proto] = that
// prototype's structure id (tabular: the only structure).
template<int MODE, bool SAMPLED, class R>
__device__ __forceinline__ int step_delta_particle(const DevModel& M, const int* __restrict__ proto_sid, int proto, int a,
                                                   const float* __restrict__ tb, int* block, int cap, int s, R& g,
                                                   int& o_out, int* rec, int* overflow, int o_real, double* likelihood)
{
    if (M.tabular)
    {
        const Node* nodes = M.nodes + (long long)a * M.J;
        int const s2      = hyper_step_delta<MODE, SAMPLED>(M, nodes, tb, block, cap, s, g, o_out, rec, overflow);
        if (likelihood) *likelihood = obs_probability_delta<SAMPLED>(M, nodes, tb, block, s2, o_real, g);
        return s2;
    }
    const Node* nodes = M.nodes + ((long long)proto_sid[proto] * M.A + a) * M.J;
    return hyper_step_journal<MODE, SAMPLED>(M, nodes, tb, block, cap, s, g, o_out, rec, overflow, o_real, likelihood, a);
}

#endif // __CUDACC__

} // namespace fba
