// Kernels of the belief / rollout hot path (sm_100a). See DESIGN.md §Kernels for the roofline of
// each; the host-side C ABI that launches them is in fba_capi.cu.
#pragma once
#include <cooperative_groups.h>

#include <cuda_pipeline.h>

#include "fba_device.cuh"

namespace fba {

constexpr int kThreads = 256;

// How a kernel obtains its random source for work item `idx`.
struct RngArgs
{
    // REPLAY
    const uint32_t* words; // device copy of the slice this operation consumes
    long long n_words;
    const long long* offsets; // per-item word offset (NULL: idx * words_per_item)
    long long words_per_item;
    // PHILOX
    unsigned long long seed, offset, stream_base;
};

template<bool REPLAY>
struct RngOf;
template<>
struct RngOf<true>
{
    using type = ReplayRng;
    __device__ static ReplayRng make(const RngArgs& r, long long idx)
    {
        long long const p = r.offsets ? r.offsets[idx] : idx * r.words_per_item;
        return ReplayRng(r.words, p, r.n_words);
    }
};
template<>
struct RngOf<false>
{
    using type = PhiloxRng;
    __device__ static PhiloxRng make(const RngArgs& r, long long idx)
    {
        return PhiloxRng(r.seed, r.stream_base + (unsigned long long)idx, r.offset);
    }
};

// words a replay source has handed out so far (0 for the counter-based source)
__device__ __forceinline__ long long rng_words_used(const ReplayRng& g, long long start) { return g.pos - start; }
__device__ __forceinline__ long long rng_words_used(const PhiloxRng&, long long) { return 0; }

// ------------------------------------------------------------------------------------------------
// particle block copies: one warp per destination particle, 16-byte vectors, streaming hints
// (every byte is touched once per update, so nothing should be kept in L1)
// ------------------------------------------------------------------------------------------------
// Cache hint of the block copies. Measured on B200 (tools/exp_bulk.py, sysadmin, 1.25e6 particles,
// k_gather): .nc.L1::no_allocate 5.73 TB/s, plain 5.84 TB/s, .cs (evict-first streaming) 5.99 TB/s
// with 4 and 6.15 TB/s with 8 loads in flight per lane; the TMA bulk-copy version (k_gather_bulk)
// 5.37 TB/s. The in-place copy (scattered sources and destinations) is best with .cs and 4.
#ifndef FBA_COPY_HINT
#define FBA_COPY_HINT 1 // 0: .nc.L1::no_allocate / .L1::no_allocate, 1: .cs (streaming), 2: plain
#endif

__device__ __forceinline__ float4 ld_stream(const float4* p)
{
    float4 v;
#if FBA_COPY_HINT == 0
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
#elif FBA_COPY_HINT == 1
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
#else
    v = *p;
#endif
    return v;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v)
{
#if FBA_COPY_HINT == 0
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
#elif FBA_COPY_HINT == 1
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
#else
    *p = v;
#endif
}

// U independent 16-byte loads in flight per lane before the first store
template<int U = 4>
__device__ __forceinline__ void warp_copy_block(const float* __restrict__ src, float* __restrict__ dst,
                                                int n_vec, int lane)
{
    const float4* s = reinterpret_cast<const float4*>(src);
    float4* d       = reinterpret_cast<float4*>(dst);
    int i           = lane;
    for (; i + 32 * (U - 1) < n_vec; i += 32 * U)
    {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = ld_stream(s + i + 32 * k);
#pragma unroll
        for (int k = 0; k < U; ++k) st_stream(d + i + 32 * k, v[k]);
    }
    for (; i < n_vec; i += 32) st_stream(d + i, ld_stream(s + i));
}

// Belief::initiate: particle i clones prototype particle_proto[i]
__global__ void __launch_bounds__(kThreads)
    k_init_from_protos(float* __restrict__ counts, long long stride, int* __restrict__ state,
                       int* __restrict__ sid, double* __restrict__ w, long long N,
                       const float* __restrict__ protos, const int* __restrict__ proto_sid,
                       const int* __restrict__ particle_proto, const int* __restrict__ particle_state)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    double const w0       = 1.0 / (double)N;
    for (long long i = warp0; i < N; i += nwarp)
    {
        int const p = particle_proto ? particle_proto[i] : 0;
        warp_copy_block(protos + (long long)p * stride, counts + i * stride, (int)(stride >> 2), lane);
        if (lane == 0)
        {
            sid[i] = proto_sid[p];
            if (particle_state) state[i] = particle_state[i];
            if (w) w[i] = w0;
        }
    }
}

// init_sampled: domain start state (and prototype) drawn on device
template<bool REPLAY>
__global__ void __launch_bounds__(kThreads)
    k_draw_init(DevModel M, long long N, int n_protos, const double* __restrict__ proto_cdf,
                int* __restrict__ particle_proto, int* __restrict__ particle_state, RngArgs ra)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto g = RngOf<REPLAY>::make(ra, i);
    int p  = 0;
    if (proto_cdf)
    {
        double const u = draw_u(g);
        while (p < n_protos - 1 && u >= proto_cdf[p]) ++p;
    }
    particle_proto[i] = p;
    particle_state[i] = sample_start_state(M, g);
}

// importance_sampling::resample's copy (ImportanceSampler.hpp:79-82): new particle j = copy of
// ancestor anc[j]; uniform weight. Also used by resetDomainStateDistribution (states overwritten).
// DOMINANT KERNEL: reads and writes every count cell once — HBM-bound.
__global__ void __launch_bounds__(kThreads)
    k_gather(const float* __restrict__ src, float* __restrict__ dst, long long stride,
             const int* __restrict__ src_state, int* __restrict__ dst_state,
             const int* __restrict__ src_sid, int* __restrict__ dst_sid,
             const int* __restrict__ struct_size, double* __restrict__ w, double w_new,
             const int* __restrict__ anc, long long n_out, int delta)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long j = warp0; j < n_out; j += nwarp)
    {
        long long const i = anc[j];
        int const id      = src_sid[i];
        // copy only the cells this particle owns (rounded to 16 bytes): its structure's CPTs, or —
        // base+delta storage — the header and the increments recorded so far
        int const n_vec = delta ? (reinterpret_cast<const int*>(src + i * stride)[0] + 1 + 3) >> 2
                                : (struct_size[id] + 3) >> 2;
        warp_copy_block<8>(src + i * stride, dst + j * stride, n_vec, lane);
        if (lane == 0)
        {
            dst_sid[j] = id;
            if (dst_state) dst_state[j] = src_state[i];
            if (w) w[j] = w_new;
        }
    }
}

// Scattered particle copies between two beliefs: job j copies src[jobs[j].x] into dst[jobs[j].y] (count
// block, domain state, structure id), one warp per job. How single particles move between the filters of
// the composite beliefs (CheatingReinvigoration.cpp:136-147, StructureIncubatorSampling.cpp:155-187);
// the host has already reduced jobs with the same destination to the last one.
__global__ void __launch_bounds__(kThreads)
    k_replace_from(const float* __restrict__ src, float* __restrict__ dst, long long stride,
                   const int* __restrict__ src_state, int* __restrict__ dst_state,
                   const int* __restrict__ src_sid, int* __restrict__ dst_sid,
                   const int* __restrict__ struct_size, const int2* __restrict__ jobs, long long n)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long j = warp0; j < n; j += nwarp)
    {
        int2 const job  = jobs[j];
        int const id    = src_sid[job.x];
        int const n_vec = (int)((stride + 3) >> 2); // the whole block: padding beyond the structure stays zero
        warp_copy_block<4>(src + (long long)job.x * stride, dst + (long long)job.y * stride, n_vec, lane);
        if (lane == 0)
        {
            dst_sid[job.y]   = id;
            dst_state[job.y] = src_state[job.x];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same gather through the TMA engine (cp.async.bulk, SASS: UBLKCP): one CTA per SM; its thread 0
// drives a ring of shared-memory stages — bulk load global -> shared (completion on an mbarrier),
// bulk store shared -> global (bulk async-group) — so whole count blocks move as single DMA
// transfers and no register or LSU bandwidth is spent on the payload. The other lanes write the
// per-particle scalars (state, structure id, weight). Loads run STAGES-1 blocks ahead of the stores.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes,
                                          unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

constexpr int kBulkStages = 8;

// dynamic shared memory: kBulkStages * stage_bytes, then kBulkStages mbarriers
__global__ void __launch_bounds__(32)
    k_gather_bulk(const float* __restrict__ src, float* __restrict__ dst, long long stride,
                  const int* __restrict__ src_state, int* __restrict__ dst_state,
                  const int* __restrict__ src_sid, int* __restrict__ dst_sid,
                  const int* __restrict__ struct_size, double* __restrict__ w, double w_new,
                  const int* __restrict__ anc, long long n_out, int stage_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)kBulkStages * stage_bytes);
    int const tid            = threadIdx.x;
    long long const first = blockIdx.x, step = gridDim.x;
    long long const L     = (n_out > first) ? (n_out - first + step - 1) / step : 0; // this CTA's jobs

    if (tid == 0)
    {
        for (int s = 0; s < kBulkStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid == 0)
    {
        auto issue_load = [&](long long n) {
            long long const j  = first + n * step;
            long long const i  = anc[j];
            unsigned const bytes = (unsigned)(((struct_size[src_sid[i]] + 3) >> 2) << 4);
            int const st       = (int)(n % kBulkStages);
            mbar_expect_tx(&bars[st], bytes);
            bulk_load(smem + (size_t)st * stage_bytes, src + i * stride, bytes, &bars[st]);
        };
        long long const ahead = kBulkStages - 1;
        for (long long n = 0; n < L && n < ahead; ++n) issue_load(n);
        for (long long n = 0; n < L; ++n)
        {
            int const st = (int)(n % kBulkStages);
            mbar_wait(&bars[st], (unsigned)((n / kBulkStages) & 1));
            long long const j    = first + n * step;
            long long const i    = anc[j];
            unsigned const bytes = (unsigned)(((struct_size[src_sid[i]] + 3) >> 2) << 4);
            bulk_store(dst + j * stride, smem + (size_t)st * stage_bytes, bytes);
            // the stage the next load wants was last used by job n + ahead - kBulkStages = n - 1:
            // its store must have finished reading shared memory (at most this iteration's may pend)
            if (n + ahead < L)
            {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                issue_load(n + ahead);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else
    {
        for (long long n = tid - 1; n < L; n += 31)
        {
            long long const j = first + n * step;
            long long const i = anc[j];
            dst_sid[j]        = src_sid[i];
            if (dst_state) dst_state[j] = src_state[i];
            if (w) w[j] = w_new;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// importance_sampling::update, per-particle part (ImportanceSampler.hpp:37-54): step in place,
// weight *= P(o | a, particle). One thread per particle.
// ------------------------------------------------------------------------------------------------
template<bool REPLAY, bool LONG, bool SAMPLED>
#ifndef FBA_PROPOSE_MIN_BLOCKS
#define FBA_PROPOSE_MIN_BLOCKS 6 // 40 registers, no spills: measured 4 % faster than 48 (tools/exp_propose.py)
#endif
__global__ void __launch_bounds__(kThreads, (LONG || SAMPLED) ? 1 : FBA_PROPOSE_MIN_BLOCKS)
    k_propose(DevModel M, float* counts, long long stride, int* __restrict__ state,
              const int* __restrict__ sid, double* __restrict__ w, long long N, int a, int o, RngArgs ra,
              int* __restrict__ overrun)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto g            = RngOf<REPLAY>::make(ra, i);
    const Node* nodes = M.nodes + ((long long)sid[i] * M.A + a) * M.J;
    float* c          = counts + i * stride;
    int sim_o;
    Feat x2;
    int const s2 = hyper_step<STEP_UPDATE, decltype(g), false, LONG, SAMPLED>(M, nodes, c, state[i], g, sim_o, x2, nullptr);
    double const prob = obs_probability<SAMPLED>(M, nodes, c, x2, o, g);
    state[i]          = s2;
    w[i]              = __dmul_rn(w[i], prob);
    if (g.overrun) *overrun = 1;
}

// ------------------------------------------------------------------------------------------------
// weights: replay (sequential, bit-exact) and native (tree) reductions
// ------------------------------------------------------------------------------------------------

// REPLAY normalisation. One block; thread 0 carries the three sequential floating-point chains the
// reference runs, everyone else only moves data:
//   A  total        = sum_i w_i in index order                 (ImportanceSampler.hpp:51)
//   B  w_i /= total; total_weight = sum_i w_i in index order   (WeightedFilter.cpp:130-143)
//   C  R_k = total_weight - w_{N-1} - ... - w_k, k = N-1..1    (WeightedFilter.cpp:168-183)
// scal[0] = total, scal[1] = total_weight. If !do_normalise only C runs (scal[1] given).
constexpr int kChunk = 2048;
__global__ void __launch_bounds__(kThreads)
    k_seq_normalize(double* __restrict__ w, long long N, double* __restrict__ scal,
                    double* __restrict__ R, int do_normalise)
{
    __shared__ double buf[kChunk];
    __shared__ double s_total;
    int const tid = threadIdx.x;

    if (do_normalise)
    {
        double acc = 0.0;
        for (long long base = 0; base < N; base += kChunk)
        {
            int const n = (int)min((long long)kChunk, N - base);
            for (int k = tid; k < n; k += kThreads) buf[k] = w[base + k];
            __syncthreads();
            if (tid == 0)
                for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, buf[k]);
            __syncthreads();
        }
        if (tid == 0) s_total = acc;
        __syncthreads();
        double const total = s_total;
        acc                = 0.0;
        for (long long base = 0; base < N; base += kChunk)
        {
            int const n = (int)min((long long)kChunk, N - base);
            for (int k = tid; k < n; k += kThreads)
            {
                double const v = __ddiv_rn(w[base + k], total);
                buf[k]         = v;
                w[base + k]    = v;
            }
            __syncthreads();
            if (tid == 0)
                for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, buf[k]);
            __syncthreads();
        }
        if (tid == 0)
        {
            scal[0] = total;
            scal[1] = acc;
            s_total = acc;
        }
        __syncthreads();
    } else
    {
        if (tid == 0) s_total = scal[1];
        __syncthreads();
    }

    double rem = s_total;
    for (long long top = N; top > 0; top -= kChunk)
    {
        long long const base = max(0ll, top - kChunk);
        int const n          = (int)(top - base);
        for (int k = tid; k < n; k += kThreads) buf[k] = w[base + k];
        __syncthreads();
        if (tid == 0)
            for (int k = n - 1; k >= 0; --k)
            {
                rem    = __dsub_rn(rem, buf[k]);
                buf[k] = rem;
            }
        __syncthreads();
        for (int k = tid; k < n; k += kThreads) R[base + k] = buf[k];
        __syncthreads();
    }
}

// REPLAY: WeightedFilter::sample for draw j (WeightedFilter.cpp:163-191): the largest k >= 1 with
// threshold > R_k, else 0. R is non-decreasing in k, so this is a binary search.
__device__ __forceinline__ int pick_from_suffix(const double* __restrict__ R, long long N, double thr)
{
    long long lo = 1, hi = N; // count k in [1,N) with R[k] < thr; they form a prefix
    while (lo < hi)
    {
        long long const mid = (lo + hi) >> 1;
        if (R[mid] < thr) lo = mid + 1;
        else
            hi = mid;
    }
    return (int)(lo - 1);
}

__global__ void __launch_bounds__(kThreads)
    k_pick_replay(const double* __restrict__ R, long long N, const double* __restrict__ scal,
                  RngArgs ra, int* __restrict__ anc, long long n_out, int* __restrict__ overrun)
{
    long long const j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    auto g           = RngOf<true>::make(ra, j);
    double const thr = __dmul_rn(draw_u(g), scal[1]);
    anc[j]           = pick_from_suffix(R, N, thr);
    if (g.overrun) *overrun = 1;
}

// NATIVE reductions: fixed tile -> deterministic result for a given N.
constexpr int kTile = 1024; // doubles per block (4 per thread)

__device__ __forceinline__ double block_reduce_sum(double v, double* sh)
{
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0)
    {
        v = (lane < kThreads / 32) ? sh[lane] : 0.0;
        for (int o = 4; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v; // valid in thread 0
}

// The native reductions / scans below are written as *_body device functions over one tile (or one
// block-wide pass) so that the one-launch-per-phase kernels and the fused one-CTA-per-run kernel
// (k_runs_update) execute literally the same arithmetic in the same order: a batched run is bit-identical
// to a stand-alone belief.
__device__ __forceinline__ void tile_sums_body(int tile, const double* __restrict__ w, long long N,
                                               double* __restrict__ tile_sum, double* sh)
{
    long long const base = (long long)tile * kTile;
    double v             = 0.0;
#pragma unroll
    for (int k = 0; k < kTile / kThreads; ++k)
    {
        long long const i = base + threadIdx.x + k * kThreads;
        if (i < N) v += w[i];
    }
    v = block_reduce_sum(v, sh);
    if (threadIdx.x == 0) tile_sum[tile] = v;
}

__global__ void __launch_bounds__(kThreads)
    k_tile_sums(const double* __restrict__ w, long long N, double* __restrict__ tile_sum)
{
    __shared__ double sh[kThreads / 32];
    tile_sums_body(blockIdx.x, w, N, tile_sum, sh);
}

// exclusive scan of the tile sums in one block; scal[0] = grand total
__device__ __forceinline__ void scan_tile_sums_body(double* __restrict__ tile_sum, int n_tiles,
                                                    double* __restrict__ scal, double* sh, double* carry)
{
    if (threadIdx.x == 0) *carry = 0.0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += kThreads)
    {
        int const i    = base + threadIdx.x;
        double const v = (i < n_tiles) ? tile_sum[i] : 0.0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kThreads; o <<= 1)
        {
            double const t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0.0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        double const incl = sh[threadIdx.x];
        if (i < n_tiles) tile_sum[i] = *carry + incl - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) *carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) scal[0] = *carry;
}

__global__ void __launch_bounds__(kThreads)
    k_scan_tile_sums(double* __restrict__ tile_sum, int n_tiles, double* __restrict__ scal)
{
    __shared__ double sh[kThreads];
    __shared__ double carry;
    scan_tile_sums_body(tile_sum, n_tiles, scal, sh, &carry);
}

// w_i /= scal[0] (optionally), inclusive prefix sums into cdf
__device__ __forceinline__ void scale_and_scan_body(int tile, double* __restrict__ w, long long N,
                                                    const double* __restrict__ tile_off, double total,
                                                    double* __restrict__ cdf, double* sh)
{
    long long const base = (long long)tile * kTile + (long long)threadIdx.x * (kTile / kThreads);
    double v[kTile / kThreads];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < kTile / kThreads; ++k)
    {
        long long const i = base + k;
        double x          = (i < N) ? w[i] / total : 0.0;
        if (i < N) w[i] = x;
        run += x;
        v[k] = run;
    }
    sh[threadIdx.x] = run;
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1)
    {
        double const t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0.0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    double const off = tile_off[tile] / total + sh[threadIdx.x] - run;
#pragma unroll
    for (int k = 0; k < kTile / kThreads; ++k)
    {
        long long const i = base + k;
        if (i < N) cdf[i] = off + v[k];
    }
}

__global__ void __launch_bounds__(kThreads)
    k_scale_and_scan(double* __restrict__ w, long long N, const double* __restrict__ tile_off,
                     const double* __restrict__ total_ptr, double divide_by, double* __restrict__ cdf)
{
    __shared__ double sh[kThreads];
    scale_and_scan_body(blockIdx.x, w, N, tile_off, total_ptr ? *total_ptr : divide_by, cdf, sh);
}

// ------------------------------------------------------------------------------------------------
// REPLAY normalisation for LARGE beliefs: the reference's three SEQUENTIAL floating-point chains
// (k_seq_normalize: A total, B total_weight, C remainders) evaluated in parallel with bit-identical
// results. One thread of k_seq_normalize needs ~8 ns per element and chain (30 ms for 1.25e6
// particles); here the chains are cut into segments of kTile elements, every segment is walked by its
// own thread from a SPECULATED start value (a tree-order prefix sum, off by a few ulps from the
// sequential one), and a short serial pass turns the speculated segments into the exact ones:
//
//   Inside one binade all doubles are multiples of u = ulp. If the speculated value s' and the true
//   value s of the running sum lie in the same binade, s' - s = m u, and the next rounded operation
//   fl(s + w) = u round(s / u + w / u) commutes with the shift: fl(s' + w) = fl(s + w) + m u — unless
//   the fractional part of w / u is exactly 1/2 (a tie, resolved by the parity of the neighbour) or
//   the result leaves the binade (the grid changes). Weights are >= 0, so a chain is monotone: it stays
//   in one binade iff its first and last value do.
//
// So a segment whose speculated walk met no tie, and whose speculated AND true end points all lie in
// one binade, is EXACTLY the true walk shifted by delta = s'_start - s_start (all differences exact:
// same binade, Sterbenz). Any other segment — the first one, the dozen in which the running sum
// crosses a power of two, the one or two per million elements with a tie — is recomputed sequentially
// from its true start by the serial pass. tests/test_cuda_parity.py compares this path with
// k_seq_normalize bit for bit (weights, totals, every remainder) on adversarial weight vectors.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int binade_of(double x) // sign + biased exponent
{
    return __double2hiint(x) >> 20;
}

// walks segment `c` (elements [c kTile, min(N, (c+1) kTile)) ascending for the sums, descending for the
// remainders) from start value s; returns the end value; *tie = the walk met an exact tie or a value
// too small to test; SUB writes every intermediate value to out[]
template<bool SUB>
__device__ __forceinline__ double chain_walk(const double* v, long long lo, long long hi, double s, double* out,
                                             bool* tie)
{
    bool t = false;
    if (!SUB)
    {
        for (long long k = lo; k < hi; ++k)
        {
            double const b = v[k];
            double const x = __dadd_rn(s, b);
            double const e = __dsub_rn(b, __dsub_rn(x, s)); // exact rounding error when b <= s (Fast2Sum)
            int const ex   = (__double2hiint(x) >> 20) & 0x7ff;
            t |= (ex <= 54) | (fabs(e) == __hiloint2double((ex - 53) << 20, 0));
            s = x;
        }
    } else
    {
        for (long long k = hi - 1; k >= lo; --k)
        {
            double const b = v[k];
            double const x = __dsub_rn(s, b);
            double const e = __dsub_rn(__dsub_rn(s, x), b);
            int const ex   = (__double2hiint(x) >> 20) & 0x7ff;
            t |= (ex <= 54) | (fabs(e) == __hiloint2double((ex - 53) << 20, 0));
            s      = x;
            out[k] = x;
        }
    }
    *tie = t;
    return s;
}

// speculative pass, one thread per segment. spec_prefix[c] = tree-order sum of the elements before
// segment c (k_tile_sums + k_scan_tile_sums), spec[0] = tree-order total. The start of the remainder
// chain's segment c is top - (elements at and above (c+1) kTile).
template<bool SUB>
__global__ void __launch_bounds__(64)
    k_chain_segments(const double* __restrict__ v, long long N, const double* __restrict__ spec_prefix,
                     const double* __restrict__ spec_total, const double* __restrict__ top,
                     double* __restrict__ seg_start, double* __restrict__ seg_end,
                     unsigned char* __restrict__ seg_tie, double* __restrict__ out)
{
    long long const c      = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long const n_seg  = (N + kTile - 1) / kTile;
    if (c >= n_seg) return;
    long long const lo = c * kTile, hi = min(N, lo + kTile);
    double s;
    if (!SUB) s = spec_prefix[c];
    else
        s = (c + 1 < n_seg) ? __dsub_rn(*top, __dsub_rn(*spec_total, spec_prefix[c + 1])) : *top;
    seg_start[c] = s;
    bool tie;
    seg_end[c] = chain_walk<SUB>(v, lo, hi, s, out, &tie);
    seg_tie[c] = tie ? 1 : 0;
}

// serial pass, ONE block: thread 0 carries the chain from segment to segment (true start of the first
// segment = *start0 for the remainders, 0 for the sums), shifting the speculated segments and
// recomputing the others; the other threads only stage data through shared memory (segment
// descriptors 256 at a time; the elements of a segment that has to be recomputed). result[0] = the
// chain's exact end value; SUB: delta[c] = what to subtract from the speculated remainders of segment
// c (0 where they were recomputed in place).
template<bool SUB>
__global__ void __launch_bounds__(kThreads)
    k_chain_fix(const double* __restrict__ v, long long N, const double* __restrict__ seg_start,
                const double* __restrict__ seg_end, const unsigned char* __restrict__ seg_tie,
                const double* __restrict__ start0, double* __restrict__ result, double* __restrict__ delta,
                double* __restrict__ out, long long* __restrict__ n_recomputed)
{
    __shared__ double sh_start[kThreads], sh_end[kThreads], buf[kTile];
    __shared__ unsigned char sh_tie[kThreads];
    __shared__ double s_sh;
    __shared__ int pos_sh, need_sh;
    int const tid         = threadIdx.x;
    long long const n_seg = (N + kTile - 1) / kTile;
    if (tid == 0) s_sh = start0 ? *start0 : 0.0;
    long long redo = 0;
    for (long long first = 0; first < n_seg; first += kThreads)
    { // batch of segments first .. first + n_b - 1 in PROCESSING order (remainders: from the top down)
        int const n_b = (int)min((long long)kThreads, n_seg - first);
        if (tid < n_b)
        {
            long long const c = SUB ? n_seg - 1 - (first + tid) : first + tid;
            sh_start[tid] = seg_start[c], sh_end[tid] = seg_end[c], sh_tie[tid] = seg_tie[c];
        }
        if (tid == 0) pos_sh = 0;
        __syncthreads();
        while (true)
        {
            if (tid == 0)
            {
                double s = s_sh;
                int j    = pos_sh;
                int need = -1;
                for (; j < n_b; ++j)
                {
                    double const sp = sh_start[j], ep = sh_end[j];
                    int const id    = binade_of(sp);
                    bool ok         = !sh_tie[j] && binade_of(ep) == id && binade_of(s) == id;
                    double d = 0.0, e = 0.0;
                    if (ok)
                    {
                        d  = __dsub_rn(sp, s); // exact: same binade
                        e  = __dsub_rn(ep, d);
                        ok = binade_of(e) == id;
                    }
                    if (!ok)
                    {
                        need = j;
                        break;
                    }
                    if (SUB) delta[n_seg - 1 - (first + j)] = d;
                    s = e;
                }
                s_sh = s, pos_sh = j, need_sh = need;
            }
            __syncthreads();
            if (need_sh < 0) break;
            long long const c  = SUB ? n_seg - 1 - (first + need_sh) : first + need_sh;
            long long const lo = c * kTile;
            int const n        = (int)(min(N, lo + kTile) - lo);
            for (int k = tid; k < n; k += kThreads) buf[k] = v[lo + k];
            __syncthreads();
            if (tid == 0)
            {
                bool tie;
                s_sh = chain_walk<SUB>(buf, 0, n, s_sh, buf, &tie); // SUB: buf[k] becomes the remainder
                if (SUB) delta[c] = 0.0;
                pos_sh = need_sh + 1;
                ++redo;
            }
            __syncthreads();
            if (SUB)
                for (int k = tid; k < n; k += kThreads) out[lo + k] = buf[k];
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0)
    {
        result[0] = s_sh;
        if (n_recomputed) *n_recomputed += redo;
    }
}

// remainders of the shifted segments: R_k = R'_k - delta (exact: multiples of one ulp, same binade)
__global__ void __launch_bounds__(kThreads)
    k_chain_apply_shift(double* __restrict__ R, long long N, const double* __restrict__ delta)
{
    long long const k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double const d = delta[k / kTile];
    if (d != 0.0) R[k] = __dsub_rn(R[k], d);
}

// chain B's element-wise half: w_i /= total (one correctly rounded division each) + tile sums of the result
__global__ void __launch_bounds__(kThreads)
    k_divide_tile_sums(double* __restrict__ w, long long N, const double* __restrict__ total,
                       double* __restrict__ tile_sum)
{
    __shared__ double sh[kThreads / 32];
    long long const base = (long long)blockIdx.x * kTile;
    double const t       = *total;
    double v             = 0.0;
#pragma unroll
    for (int k = 0; k < kTile / kThreads; ++k)
    {
        long long const i = base + threadIdx.x + k * kThreads;
        if (i < N)
        {
            double const x = __ddiv_rn(w[i], t);
            w[i]           = x;
            v += x;
        }
    }
    v = block_reduce_sum(v, sh);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = v;
}

// NATIVE ancestor selection over the inclusive cdf (cdf[N-1] ~ 1).
// systematic: threshold_j = (j + u0) / n_out; multinomial: threshold_j = u_j.
__global__ void __launch_bounds__(kThreads)
    k_pick_native(const double* __restrict__ cdf, long long N, long long n_out, int systematic,
                  RngArgs ra, int* __restrict__ anc)
{
    long long const j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    double thr;
    if (systematic)
    {
        auto g = RngOf<false>::make(ra, 0);
        thr    = ((double)j + draw_u(g)) / (double)n_out;
    } else
    {
        auto g = RngOf<false>::make(ra, j);
        thr    = draw_u(g);
    }
    thr *= cdf[N - 1];
    long long lo = 0, hi = N - 1; // first i with cdf[i] > thr
    while (lo < hi)
    {
        long long const mid = (lo + hi) >> 1;
        if (cdf[mid] > thr) hi = mid;
        else
            lo = mid + 1;
    }
    anc[j] = (int)lo;
}

// ------------------------------------------------------------------------------------------------
// NATIVE in-place systematic resampling. Survivors (>= 1 offspring) keep their slot and are never
// moved; each dead slot receives one of the extra offspring of a multiply-drawn particle. Only the
// duplicated blocks cross HBM (2 x 4C bytes each) instead of every block.
//   offspring: n_i = F(cdf_i) - F(cdf_{i-1}),  F(c) = clamp(ceil(c / s - u), 0, n_out),
//              s = cdf_{N-1} / n_out  (systematic positions (j + u) * s, j = 0..n_out-1)
//   dead_i = (n_i == 0), extra_i = max(n_i - 1, 0); exclusive scans give the k-th dead slot and,
//   by binary search over the extra scan, the source of the k-th extra copy.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long sys_count_below(double c, double inv_s, double u, long long n_out)
{
    double const v = ceil(c * inv_s - u);
    if (!(v > 0.0)) return 0;
    if (v >= (double)n_out) return n_out;
    return (long long)v;
}

// multi-GPU: this rank's offspring quota from the all-gathered shard totals, on device, so the
// host need not synchronise before the resampling kernels are enqueued. Systematic allocation over
// ranks: quota_g = #{j : (j + u)/n_total in (C_{g-1}, C_g]}. out_total[0] = W, out_quota[0] = quota.
__global__ void k_shard_quota(const double* __restrict__ totals, int n_ranks, int rank, double u,
                              long long n_local, double* __restrict__ out_total,
                              long long* __restrict__ out_quota)
{
    double W = 0.0;
    for (int g = 0; g < n_ranks; ++g) W += totals[g];
    long long const n_total = n_local * n_ranks;
    long long prev = 0, quota = 0;
    double acc = 0.0;
    for (int g = 0; g <= rank; ++g)
    {
        acc += totals[g];
        long long edge = (g == n_ranks - 1) ? n_total : (long long)floor(acc / W * (double)n_total - u + 1.0);
        edge  = min(max(edge, prev), n_total);
        quota = edge - prev;
        prev  = edge;
    }
    out_total[0] = W;
    out_quota[0] = (W > 0.0) ? quota : 0;
}

// multi-GPU, peer-to-peer exchange: the whole exchange plan on device (n_ranks <= 16), identical on
// every rank because it is a pure function of the all-gathered totals and u.
//   plan[g*G + h] = records rank g ships to rank h (greedy matching in rank order)
// out_quota[0] = this rank's quota, out_total[0] = W.
constexpr int kMaxRanks = 16;
__global__ void k_shard_plan(const double* __restrict__ totals, int n_ranks, int rank, double u,
                             long long n_local, double* __restrict__ out_total,
                             long long* __restrict__ out_quota, long long* __restrict__ plan)
{
    double W = 0.0;
    for (int g = 0; g < n_ranks; ++g) W += totals[g];
    long long const n_total = n_local * n_ranks;
    long long quota[kMaxRanks], surplus[kMaxRanks], deficit[kMaxRanks];
    long long prev = 0;
    double acc     = 0.0;
    for (int g = 0; g < n_ranks; ++g)
    {
        acc += totals[g];
        long long edge = (g == n_ranks - 1) ? n_total : (long long)floor(acc / W * (double)n_total - u + 1.0);
        edge     = min(max(edge, prev), n_total);
        quota[g] = (W > 0.0) ? edge - prev : n_local;
        prev     = edge;
        surplus[g] = max(0ll, quota[g] - n_local);
        deficit[g] = max(0ll, n_local - quota[g]);
    }
    for (int k = 0; k < n_ranks * n_ranks; ++k) plan[k] = 0;
    int h = 0;
    for (int g = 0; g < n_ranks; ++g)
        while (surplus[g] > 0)
        {
            while (h < n_ranks && deficit[h] == 0) ++h;
            if (h >= n_ranks) break;
            long long const k = min(surplus[g], deficit[h]);
            plan[g * n_ranks + h] += k;
            surplus[g] -= k, deficit[h] -= k;
        }
    out_total[0] = W;
    out_quota[0] = quota[rank];
}

// where surplus record r of rank `rank` goes: destination rank and record offset inside that rank's
// import buffer (incoming records are ordered by sender rank)
__device__ __forceinline__ void p2p_destination(const long long* __restrict__ plan, int n_ranks, int rank,
                                                long long r, int& dst_rank, long long& dst_off)
{
    long long before = 0;
    for (int h = 0; h < n_ranks; ++h)
    {
        long long const n = plan[rank * n_ranks + h];
        if (r < before + n)
        {
            long long off = r - before;
            for (int g = 0; g < rank; ++g) off += plan[g * n_ranks + h];
            dst_rank = h, dst_off = off;
            return;
        }
        before += n;
    }
    dst_rank = -1, dst_off = 0;
}

// ---- cross-GPU synchronisation through peer-mapped memory (NVLink / NVSwitch) ----
// Every rank owns one control block in its HBM, mapped by every peer (CUDA IPC). Ranks signal each
// other by storing a monotonically increasing STEP STAMP with release semantics at system scope
// into the waiter's own block, and wait by polling their OWN memory with acquire loads — no host,
// no NCCL, no cross-GPU polling traffic.
struct P2PCtrl
{
    double box[2][kMaxRanks];                  // [step parity][sender]: shard weight totals
    unsigned long long total_flag[kMaxRanks];  // [sender]: its total for that step is in box
    unsigned long long dead_flag[kMaxRanks];   // [owner]: that rank's dead-slot list of that step is complete
    unsigned long long landed_flag[kMaxRanks]; // [sender]: its peer stores of that step are done
    int totals[2];                             // this rank's (#dead, #extra) of the current resample
    int pad[2];
};
constexpr long long kP2PDeadOffset = 4096; // the rank's dead-slot list (N ints) follows the header
static_assert(sizeof(P2PCtrl) <= kP2PDeadOffset, "control block header");

struct PeerTable
{
    P2PCtrl* ctrl[kMaxRanks]; // own entry = local
    float* counts[kMaxRanks];
    int* state[kMaxRanks];
    int* sid[kMaxRanks];
    long long timeout_ns; // a wait longer than this is reported in stats[3] and abandoned (never hang the GPU)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_sys_s32(const int* p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// polls OWN memory until the flag carries at least `step`; false after timeout_ns
__device__ __forceinline__ bool p2p_wait(const unsigned long long* flag, unsigned long long step, long long timeout_ns)
{
    if (ld_acquire_sys(flag) >= step) return true;
    unsigned long long const t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < step)
    {
        if ((long long)(global_timer_ns() - t0) > timeout_ns) return false;
        __nanosleep(64);
    }
    return true;
}

// One warp. Replaces the all-gather of the shard totals AND k_shard_plan: lane g stores this rank's
// total into rank g's box (value, then stamp with release), polls its own box for rank g's total,
// and lane 0 computes quotas and the exchange plan — identical on every rank because it is a pure
// function of the G totals and u. Advances the belief's step stamp.
__global__ void k_p2p_plan(P2PCtrl* me, PeerTable peers, int n_ranks, int rank, const double* __restrict__ local_total,
                           double u, long long n_local, double* __restrict__ out_total,
                           long long* __restrict__ out_quota, long long* __restrict__ plan,
                           unsigned long long* __restrict__ d_step, long long* __restrict__ stats)
{
    __shared__ double tot[kMaxRanks];
    int const lane                = threadIdx.x;
    unsigned long long const step = *d_step + 1;
    __syncwarp();
    if (lane == 0) *d_step = step;
    if (lane < n_ranks)
    {
        P2PCtrl* peer = peers.ctrl[lane];
        st_relaxed_sys_f64(&peer->box[step & 1][rank], *local_total);
        st_release_sys(&peer->total_flag[rank], step);
        if (!p2p_wait(&me->total_flag[lane], step, peers.timeout_ns)) atomicAdd((unsigned long long*)&stats[3], 1ull);
        tot[lane] = ld_relaxed_sys_f64(&me->box[step & 1][lane]);
    }
    __syncwarp();
    if (lane != 0) return;
    double W = 0.0;
    for (int g = 0; g < n_ranks; ++g) W += tot[g];
    long long const n_total = n_local * n_ranks;
    long long quota[kMaxRanks], surplus[kMaxRanks], deficit[kMaxRanks];
    long long prev = 0;
    double acc     = 0.0;
    for (int g = 0; g < n_ranks; ++g)
    {
        acc += tot[g];
        long long edge = (g == n_ranks - 1) ? n_total : (long long)floor(acc / W * (double)n_total - u + 1.0);
        edge     = min(max(edge, prev), n_total);
        quota[g] = (W > 0.0) ? edge - prev : n_local;
        prev     = edge;
        surplus[g] = max(0ll, quota[g] - n_local);
        deficit[g] = max(0ll, n_local - quota[g]);
    }
    for (int k = 0; k < n_ranks * n_ranks; ++k) plan[k] = 0;
    int h = 0;
    for (int g = 0; g < n_ranks; ++g)
        while (surplus[g] > 0)
        {
            while (h < n_ranks && deficit[h] == 0) ++h;
            if (h >= n_ranks) break;
            long long const k = min(surplus[g], deficit[h]);
            plan[g * n_ranks + h] += k;
            surplus[g] -= k, deficit[h] -= k;
        }
    out_total[0] = W;
    out_quota[0] = quota[rank];
}

// One warp, after a kernel whose results peers are waiting for: lane g stamps this rank's flag in
// rank g's control block. which = 0: dead-slot list complete, 1: peer stores done. The fence makes
// everything the preceding kernels of this stream wrote (peer stores included) visible system-wide
// before the stamp.
__global__ void k_p2p_signal(PeerTable peers, int n_ranks, int rank, int which,
                             const unsigned long long* __restrict__ d_step)
{
    int const lane = threadIdx.x;
    if (lane >= n_ranks) return;
    __threadfence_system();
    P2PCtrl* peer = peers.ctrl[lane];
    st_release_sys(which == 0 ? &peer->dead_flag[rank] : &peer->landed_flag[rank], *d_step);
}

// One warp: returns once every rank that ships records to this one has stamped "landed" for this
// step — after this kernel the imported particles are in place.
__global__ void k_p2p_wait_landed(const P2PCtrl* me, int n_ranks, int rank, const long long* __restrict__ plan,
                                  const unsigned long long* __restrict__ d_step, long long timeout_ns,
                                  long long* __restrict__ stats)
{
    int const lane = threadIdx.x;
    if (lane >= n_ranks || plan[lane * n_ranks + rank] == 0) return;
    if (!p2p_wait(&me->landed_flag[lane], *d_step, timeout_ns)) atomicAdd((unsigned long long*)&stats[3], 1ull);
}

__device__ __forceinline__ void offspring_body(int tile, const double* __restrict__ cdf, long long N,
                                               long long n_out, const RngArgs& ra, int* __restrict__ noff,
                                               int2* __restrict__ tile_sum, int* shd, int* she)
{
    auto g             = RngOf<false>::make(ra, 0);
    double const u     = draw_u(g);
    // a shard whose weights all collapsed gets no offspring: every slot is dead
    double const inv_s = (n_out > 0 && cdf[N - 1] > 0.0) ? (double)n_out / cdf[N - 1] : 0.0;
    if (inv_s == 0.0) n_out = 0;
    long long const base = (long long)tile * kTile;
    int d = 0, e = 0;
#pragma unroll
    for (int k = 0; k < kTile / kThreads; ++k)
    {
        long long const i = base + threadIdx.x + k * kThreads;
        if (i < N)
        {
            long long const hi = (i == N - 1) ? n_out : sys_count_below(cdf[i], inv_s, u, n_out);
            long long const lo = (i == 0) ? 0 : sys_count_below(cdf[i - 1], inv_s, u, n_out);
            int const n        = (int)(hi - lo);
            noff[i]            = n;
            d += (n == 0);
            e += (n > 1) ? n - 1 : 0;
        }
    }
    for (int o = 16; o; o >>= 1)
    {
        d += __shfl_down_sync(0xffffffffu, d, o);
        e += __shfl_down_sync(0xffffffffu, e, o);
    }
    int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) shd[wid] = d, she[wid] = e;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int td = 0, te = 0;
        for (int k = 0; k < kThreads / 32; ++k) td += shd[k], te += she[k];
        tile_sum[tile] = make_int2(td, te);
    }
}

// per tile: offspring counts -> (dead, extra) tile sums
__global__ void __launch_bounds__(kThreads)
    k_offspring(const double* __restrict__ cdf, long long N, long long n_out,
                const long long* __restrict__ n_out_ptr, RngArgs ra, int* __restrict__ noff,
                int2* __restrict__ tile_sum)
{
    __shared__ int shd[kThreads / 32], she[kThreads / 32];
    if (n_out_ptr) n_out = *n_out_ptr;
    offspring_body(blockIdx.x, cdf, N, n_out, ra, noff, tile_sum, shd, she);
}

// exclusive scan of the (dead, extra) tile sums in one block; totals[0] = #dead, totals[1] = #extra
__device__ __forceinline__ void scan_tile_pairs_body(int2* __restrict__ tile_sum, int n_tiles,
                                                     int* __restrict__ totals, long long* __restrict__ stats,
                                                     int* shd, int* she, int* cd, int* ce)
{
    if (threadIdx.x == 0) *cd = 0, *ce = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += kThreads)
    {
        int const i  = base + threadIdx.x;
        int2 const v = (i < n_tiles) ? tile_sum[i] : make_int2(0, 0);
        shd[threadIdx.x] = v.x, she[threadIdx.x] = v.y;
        __syncthreads();
        for (int o = 1; o < kThreads; o <<= 1)
        {
            int const a = (threadIdx.x >= o) ? shd[threadIdx.x - o] : 0;
            int const b = (threadIdx.x >= o) ? she[threadIdx.x - o] : 0;
            __syncthreads();
            shd[threadIdx.x] += a, she[threadIdx.x] += b;
            __syncthreads();
        }
        if (i < n_tiles) tile_sum[i] = make_int2(*cd + shd[threadIdx.x] - v.x, *ce + she[threadIdx.x] - v.y);
        __syncthreads();
        if (threadIdx.x == kThreads - 1) *cd += shd[kThreads - 1], *ce += she[kThreads - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        totals[0] = *cd, totals[1] = *ce;
        if (stats)
        {
            stats[0] += *ce; // block copies this resample performs
            stats[1] += 1;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
    k_scan_tile_pairs(int2* __restrict__ tile_sum, int n_tiles, int* __restrict__ totals,
                      long long* __restrict__ stats)
{
    __shared__ int shd[kThreads], she[kThreads];
    __shared__ int cd, ce;
    scan_tile_pairs_body(tile_sum, n_tiles, totals, stats, shd, she, &cd, &ce);
}

// per tile: dead_slot[k] = index of the k-th dead particle; extra_scan[i] = exclusive scan of extra
__device__ __forceinline__ void offspring_apply_body(int tile, const int* __restrict__ noff, long long N,
                                                     const int2* __restrict__ tile_off,
                                                     int* __restrict__ dead_slot, int* __restrict__ extra_scan,
                                                     int* __restrict__ src_of, long long src_cap, int* shd, int* she)
{
    constexpr int PER    = kTile / kThreads;
    long long const base = (long long)tile * kTile + (long long)threadIdx.x * PER;
    int n[PER];
    int d = 0, e = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k)
    {
        long long const i = base + k;
        n[k]              = (i < N) ? noff[i] : 1;
        d += (n[k] == 0);
        e += (n[k] > 1) ? n[k] - 1 : 0;
    }
    shd[threadIdx.x] = d, she[threadIdx.x] = e;
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1)
    {
        int const a = (threadIdx.x >= o) ? shd[threadIdx.x - o] : 0;
        int const b = (threadIdx.x >= o) ? she[threadIdx.x - o] : 0;
        __syncthreads();
        shd[threadIdx.x] += a, she[threadIdx.x] += b;
        __syncthreads();
    }
    int2 const off = tile_off[tile];
    int pd = off.x + shd[threadIdx.x] - d, pe = off.y + she[threadIdx.x] - e;
#pragma unroll
    for (int k = 0; k < PER; ++k)
    {
        long long const i = base + k;
        if (i < N)
        {
            extra_scan[i] = pe;
            if (n[k] == 0) dead_slot[pd++] = (int)i;
            // source of the copies pe, pe+1 (a particle rarely has more than two extras; further
            // entries keep the -1 they were memset to and are resolved by binary search)
            if (n[k] > 1 && pe < src_cap) src_of[pe] = (int)i;
            if (n[k] > 2 && pe + 1 < src_cap) src_of[pe + 1] = (int)i;
            pe += (n[k] > 1) ? n[k] - 1 : 0;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
    k_offspring_apply(const int* __restrict__ noff, long long N, const int2* __restrict__ tile_off,
                      int* __restrict__ dead_slot, int* __restrict__ extra_scan, int* __restrict__ src_of,
                      long long src_cap)
{
    __shared__ int shd[kThreads], she[kThreads];
    offspring_apply_body(blockIdx.x, noff, N, tile_off, dead_slot, extra_scan, src_of, src_cap, shd, she);
}

// multi-GPU: imported record r fills the (n_extra + r)-th dead slot (those the local extras left)
__global__ void __launch_bounds__(kThreads)
    k_import_inplace(float* __restrict__ dst, long long stride, int* __restrict__ state, int* __restrict__ sid,
                     const int* __restrict__ dead_slot, const int* __restrict__ totals, long long slot_offset,
                     long long n, const char* __restrict__ in, long long rec_bytes)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    long long const first = totals[1] + slot_offset;
    for (long long r = warp0; r < n; r += nwarp)
    {
        const char* rec      = in + r * rec_bytes;
        long long const slot = dead_slot[first + r];
        warp_copy_block(reinterpret_cast<const float*>(rec), dst + slot * stride, (int)(stride >> 2), lane);
        if (lane == 0)
        {
            const int* tail = reinterpret_cast<const int*>(rec + stride * sizeof(float));
            state[slot]     = tail[0];
            sid[slot]       = tail[1];
        }
    }
}

// copy k in [0, n_copies): source = the last i with extra_scan[i] <= k, destination = dead_slot[k]
// (k < n_fill) or export record k - n_fill. One warp per copy, within the SAME buffer: sources are
// survivors (never written), destinations are dead (never read).
__global__ void __launch_bounds__(kThreads)
    k_copy_inplace(float* counts, long long stride, int* state, int* sid, const int* __restrict__ struct_size,
                   const int* __restrict__ extra_scan, const int* __restrict__ src_of, long long N,
                   const int* __restrict__ dead_slot, const int* __restrict__ totals, char* __restrict__ xport,
                   long long rec_bytes, long long xport_cap, long long* __restrict__ stats, long long src_cap,
                   const long long* __restrict__ plan, int n_ranks, int rank, PeerTable peers, int delta,
                   const unsigned long long* __restrict__ d_step)
{
    long long const n_dead = totals[0];
    long long const n_fill = min(n_dead, (long long)totals[1]); // the rest (if any) is this shard's surplus
    long long n_copies     = totals[1];
    if (!plan && n_copies - n_fill > xport_cap)
    { // more surplus than the export buffer holds: drop the excess and report it
        if (blockIdx.x == 0 && threadIdx.x == 0) stats[2] = n_copies - n_fill - xport_cap;
        n_copies = n_fill + xport_cap;
    }
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = warp0; k < n_copies; k += nwarp)
    {
        long long i = (k < src_cap) ? src_of[k] : -1;
        if (i < 0)
        { // third or later extra of one particle: the last i with extra_scan[i] <= k
            long long lo = 0, hi = N;
            while (lo < hi)
            {
                long long const mid = (lo + hi) >> 1;
                if (extra_scan[mid] > (int)k) hi = mid;
                else
                    lo = mid + 1;
            }
            i = lo - 1;
        }
        int const id = sid[i];
        if (k < n_fill)
        {
            long long const j = dead_slot[k];
            int const n_vec   = delta ? (reinterpret_cast<const int*>(counts + i * stride)[0] + 1 + 3) >> 2
                                      : (struct_size[id] + 3) >> 2;
            warp_copy_block(counts + i * stride, counts + j * stride, n_vec, lane);
            if (lane == 0)
            {
                sid[j]   = id;
                state[j] = state[i];
            }
        } else
        {
            if (plan)
            { // peer-to-peer: the surplus block goes STRAIGHT into a dead slot of the destination GPU's
              // particle array over NVLink — no staging buffer, no import pass, no capacity limit.
              // Slot = the destination's dead list at (its own extras + records of lower-ranked
              // senders + r), read from the destination's control block once that rank has stamped
              // its list complete for this step.
                int dst_rank;
                long long dst_off;
                p2p_destination(plan, n_ranks, rank, k - n_fill, dst_rank, dst_off);
                int slot = -1;
                if (lane == 0 && dst_rank >= 0)
                {
                    if (p2p_wait(&peers.ctrl[rank]->dead_flag[dst_rank], *d_step, peers.timeout_ns))
                    {
                        const P2PCtrl* pc = peers.ctrl[dst_rank];
                        int const extras  = ld_relaxed_sys_s32(&pc->totals[1]);
                        const int* dl     = reinterpret_cast<const int*>(reinterpret_cast<const char*>(pc) + kP2PDeadOffset);
                        slot              = ld_relaxed_sys_s32(dl + extras + dst_off);
                    }
                }
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot < 0)
                {
                    if (lane == 0) atomicAdd((unsigned long long*)&stats[3], 1ull);
                    continue;
                }
                int const n_vec = delta ? (reinterpret_cast<const int*>(counts + i * stride)[0] + 1 + 3) >> 2
                                        : (struct_size[id] + 3) >> 2;
                warp_copy_block(counts + i * stride, peers.counts[dst_rank] + (long long)slot * stride, n_vec, lane);
                if (lane == 0)
                {
                    peers.sid[dst_rank][slot]   = id;
                    peers.state[dst_rank][slot] = state[i];
                }
                continue;
            }
            char* rec = xport + (k - n_fill) * rec_bytes;
            warp_copy_block(counts + i * stride, reinterpret_cast<float*>(rec), (int)(stride >> 2), lane);
            if (lane == 0)
            {
                int* tail = reinterpret_cast<int*>(rec + stride * sizeof(float));
                tail[0]   = state[i];
                tail[1]   = id;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// MANY INDEPENDENT RUNS on one GPU. The reference's real workloads are hundreds
// to thousands of independent runs of a SMALL belief (episodic tiger: 1024 particles), each of which
// is launch-latency-bound on its own (nine launches, ~40 us per update). Here the beliefs of R runs
// lie back to back in ONE particle array (run r owns particles [r n, (r+1) n)), and ONE launch with
// ONE CTA PER RUN performs the whole update of every run — propose with the run's own (action,
// observation), normalise, systematic in-place resample — by calling the same *_body functions as the
// one-launch-per-phase kernels, phase after phase, with __syncthreads() in between. Run r draws from
// Philox streams keyed (seed + r, local particle index), i.e. exactly what a stand-alone belief of n
// particles seeded with seed + r draws: a batched run is bit-identical to that belief
// (tests/test_cuda_runs.py).
//   MODE 0: updateEstimation (update + resample); 1: resetDomainStateDistribution (resample + start
//   states); 2: sample (one particle index per run); 3: normalise only (cdf for fba_runs_plan)
// ------------------------------------------------------------------------------------------------
struct RunsArgs
{
    float* counts;
    long long stride;
    int *state, *sid;
    double *w, *cdf;
    int *noff, *escan, *dead, *src_of; // per particle (src_of: n per run)
    double* tile;                      // [R][n_tiles]
    int2* tile_pairs;                  // [R][n_tiles]
    int* totals;                       // [R][2]
    double* scal;                      // [R]: the run's total weight before normalisation (step likelihood)
    long long* picked;                 // [R]: MODE 2 result (local particle index)
    long long n;                       // particles per run
    int n_tiles;
    const int *action, *observation;   // [R] (MODE 0); NULL: every run takes action0 / observation0
    int action0, observation0;
    const unsigned char* active;       // [R] or NULL: runs with 0 are left untouched
    const int* struct_size;
    unsigned long long* copies;        // [1]: block copies made, all runs (may be NULL)
    long long* stats;                  // single belief (R = 1): fba_belief's [copies, resamples] (may be NULL)
};

template<bool LONG, bool SAMPLED, int MODE>
__global__ void __launch_bounds__(kThreads)
    k_runs_step(DevModel M, RunsArgs A, RngArgs ra)
{
    __shared__ double sh_d[kThreads];
    __shared__ int sh_a[kThreads], sh_b[kThreads];
    __shared__ double carry;
    __shared__ int cd, ce;
    int const r = blockIdx.x;
    if (A.active && !A.active[r]) return;
    long long const n  = A.n;
    long long const p0 = (long long)r * n; // first particle of the run
    float* const counts = A.counts + p0 * A.stride;
    int* const state    = A.state + p0;
    int* const sid      = A.sid + p0;
    double* const w     = A.w + p0;
    double* const cdf   = A.cdf + p0;
    int* const noff     = A.noff + p0;
    int* const escan    = A.escan + p0;
    int* const dead     = A.dead + p0;
    int* const src_of   = A.src_of + p0;
    double* const tile  = A.tile + (long long)r * A.n_tiles;
    int2* const pairs   = A.tile_pairs + (long long)r * A.n_tiles;
    int* const totals   = A.totals + 2 * r;
    RngArgs rr          = ra; // the run's own random source: a stand-alone belief seeded seed + r
    rr.seed             = ra.seed + (unsigned long long)r;

    if constexpr (MODE == 2)
    { // Belief::sample of every run. A run's weights are uniform whenever this can be called (init, update +
      // resample and reset all leave 1 / n), so the weighted draw is floor(u n) — the rule the stand-alone belief
      // uses in that state (fba_belief_sample), with the same u
        if (threadIdx.x == 0)
        {
            auto g         = RngOf<false>::make(rr, 0);
            double const u = draw_u(g);
            A.picked[r]    = min(n - 1, (long long)floor(__dmul_rn(u, (double)n)));
        }
        return;
    }

    if (MODE == 0)
    { // importance_sampling::update, per-particle part (k_propose)
        int const a = A.action ? A.action[r] : A.action0, o = A.observation ? A.observation[r] : A.observation0;
        for (long long i = threadIdx.x; i < n; i += kThreads)
        {
            auto g            = RngOf<false>::make(rr, i);
            const Node* nodes = M.nodes + ((long long)sid[i] * M.A + a) * M.J;
            float* c          = counts + i * A.stride;
            int sim_o;
            Feat x2;
            int const s2 = hyper_step<STEP_UPDATE, decltype(g), false, LONG, SAMPLED>(M, nodes, c, state[i], g, sim_o,
                                                                                      x2, nullptr);
            double const prob = obs_probability<SAMPLED>(M, nodes, c, x2, o, g);
            state[i]          = s2;
            w[i]              = __dmul_rn(w[i], prob);
        }
        ++rr.offset;
        __syncthreads();
    }
    // native_normalize: tile sums -> scan -> divide (MODE 0: by the total; otherwise by 1) + cdf
    for (int t = 0; t < A.n_tiles; ++t)
    {
        tile_sums_body(t, w, n, tile, sh_d);
        __syncthreads();
    }
    scan_tile_sums_body(tile, A.n_tiles, A.scal + r, sh_d, &carry);
    __syncthreads();
    double const total = (MODE == 0) ? A.scal[r] : 1.0;
    for (int t = 0; t < A.n_tiles; ++t)
    {
        scale_and_scan_body(t, w, n, tile, total, cdf, sh_d);
        __syncthreads();
    }
    // MODE 3: normalise only — the cdf a planner's root sampling reads
    if (MODE == 2)
    { // WeightedFilter::sample on the native cdf (k_pick_native, multinomial, one draw)
        if (threadIdx.x == 0)
        {
            auto g     = RngOf<false>::make(rr, 0);
            double thr = draw_u(g);
            thr *= cdf[n - 1];
            long long lo = 0, hi = n - 1;
            while (lo < hi)
            {
                long long const mid = (lo + hi) >> 1;
                if (cdf[mid] > thr) hi = mid;
                else
                    lo = mid + 1;
            }
            A.picked[r] = lo;
        }
    }
    if constexpr (MODE < 2)
    {
        // resample_inplace
        for (int t = 0; t < A.n_tiles; ++t)
        {
            offspring_body(t, cdf, n, n, rr, noff, pairs, sh_a, sh_b);
            __syncthreads();
        }
        ++rr.offset;
        scan_tile_pairs_body(pairs, A.n_tiles, totals, nullptr, sh_a, sh_b, &cd, &ce);
        for (long long k = threadIdx.x; k < n; k += kThreads) src_of[k] = -1;
        __syncthreads();
        for (int t = 0; t < A.n_tiles; ++t)
        {
            offspring_apply_body(t, noff, n, pairs, dead, escan, src_of, n, sh_a, sh_b);
            __syncthreads();
        }
        { // k_copy_inplace, the run's own warps: the k-th extra copy fills the k-th dead slot
            long long const n_fill = min((long long)totals[0], (long long)totals[1]);
            if (threadIdx.x == 0 && A.copies) atomicAdd(A.copies, (unsigned long long)n_fill);
            if (threadIdx.x == 0 && A.stats) A.stats[0] += n_fill, A.stats[1] += 1;
            int const lane = threadIdx.x & 31;
            for (long long k = threadIdx.x >> 5; k < n_fill; k += kThreads / 32)
            {
                long long i = src_of[k];
                if (i < 0)
                {
                    long long lo = 0, hi = n;
                    while (lo < hi)
                    {
                        long long const mid = (lo + hi) >> 1;
                        if (escan[mid] > (int)k) hi = mid;
                        else
                            lo = mid + 1;
                    }
                    i = lo - 1;
                }
                int const id      = sid[i];
                long long const j = dead[k];
                warp_copy_block(counts + i * A.stride, counts + j * A.stride, (A.struct_size[id] + 3) >> 2, lane);
                if (lane == 0)
                {
                    sid[j]   = id;
                    state[j] = state[i];
                }
            }
        }
        double const uniform = 1.0 / (double)n;
        for (long long i = threadIdx.x; i < n; i += kThreads) w[i] = uniform;
        if (MODE == 1)
        { // BABelief::resetDomainStateDistribution: fresh start states (k_reset_states)
            __syncthreads();
            for (long long i = threadIdx.x; i < n; i += kThreads)
            {
                auto g   = RngOf<false>::make(rr, i);
                state[i] = sample_start_state(M, g);
            }
        }
    } // MODE < 2
}

// fba_runs_init: prototype + start state per particle, run r drawing like a stand-alone
// fba_belief_init_sampled with seed + r (k_draw_init)
__global__ void __launch_bounds__(kThreads)
    k_runs_draw_init(DevModel M, long long n, long long n_total, int n_protos,
                     const double* __restrict__ proto_cdf, int* __restrict__ particle_proto,
                     int* __restrict__ particle_state, RngArgs ra)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    RngArgs rr = ra;
    rr.seed    = ra.seed + (unsigned long long)(i / n);
    auto g     = RngOf<false>::make(rr, i % n);
    int p      = 0;
    if (proto_cdf)
    {
        double const u = draw_u(g);
        while (p < n_protos - 1 && u >= proto_cdf[p]) ++p;
    }
    particle_proto[i] = p;
    particle_state[i] = sample_start_state(M, g);
}

__global__ void __launch_bounds__(kThreads) k_fill(double* __restrict__ w, long long N, double v)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) w[i] = v;
}

// ------------------------------------------------------------------------------------------------
// resetDomainStateDistribution: fresh domain start state per particle
// (BAPOMDP::resetDomainState, BAPOMDP.cpp:69-77). Weighted beliefs first pick an ancestor
// (k_pick_*) and gather; the start-state draws of item j follow its pick in the stream.
// ------------------------------------------------------------------------------------------------
template<bool REPLAY>
__global__ void __launch_bounds__(kThreads)
    k_reset_states(DevModel M, int* __restrict__ state, long long N, RngArgs ra, int skip_words,
                   int* __restrict__ overrun)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto g = RngOf<REPLAY>::make(ra, i);
    for (int k = 0; k < skip_words; ++k) g.next(); // the weighted pick's uniform (2 words)
    state[i] = sample_start_state(M, g);
    if (g.overrun) *overrun = 1;
}

// ------------------------------------------------------------------------------------------------
// rollouts: RBAPOUCT::rollout (RBAPOUCT.cpp:295-323), one thread per rollout, counts read-only
// ------------------------------------------------------------------------------------------------
// simulated steps actually executed (rollouts end early at terminal states): one atomic per warp
__device__ __forceinline__ void count_steps(unsigned long long* steps_done, int mine)
{
    unsigned const active = __activemask();
    unsigned const total  = __reduce_add_sync(active, (unsigned)mine);
    if ((threadIdx.x & 31) == (__ffs(active) - 1) && total) atomicAdd(steps_done, (unsigned long long)total);
}

// COOP = false: one thread per rollout (large batches: most independent work per SM).
// COOP = true: one warp per rollout, rows loaded cooperatively (small batches are latency-bound:
// one coalesced request per row instead of `range` dependent ones).
template<bool REPLAY, bool COOP, bool LONG, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_rollouts(DevModel M, const float* counts, long long stride, const int* __restrict__ sid,
               long long n, const long long* __restrict__ particle, const int* __restrict__ start,
               const int* __restrict__ depth, double discount, RngArgs ra, double* __restrict__ ret_out,
               int* __restrict__ overrun, unsigned long long* __restrict__ steps_done)
{
    long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long const r = COOP ? (t >> 5) : t;
    if (r >= n) return;
    auto g            = RngOf<REPLAY>::make(ra, r);
    long long const p = particle[r];
    float* c          = const_cast<float*>(counts) + p * stride; // STEP_KEEP never writes
    const Node* base  = M.nodes + (long long)sid[p] * M.A * M.J;
    // Discount(_discount.toDouble()) starts at 1 (Discount.cpp:3-6)
    double ret = 0.0, disc = 1.0;
    int s = start[r], d = depth[r];
    bool terminal = false;
    while (d > 0 && !terminal)
    {
        int const a = random_action(M, g);
        int o;
        Feat x2;
        int const s2 =
            hyper_step<STEP_KEEP, decltype(g), COOP, LONG, SAMPLED>(M, base + (long long)a * M.J, c, s, g, o, x2, nullptr);
        double const rew = domain_reward(M, s, a, s2, terminal);
        ret  = __dadd_rn(ret, __dmul_rn(rew, disc)); // Return::add (Return.cpp:6-9)
        disc = __dmul_rn(disc, discount);            // Discount::increment (Discount.cpp:8-11)
        s    = s2;
        --d;
    }
    if (!COOP || (threadIdx.x & 31) == 0)
    {
        ret_out[r] = ret;
        if (g.overrun) *overrun = 1;
    }
    count_steps(steps_done, COOP ? ((threadIdx.x & 31) == 0 ? depth[r] - d : 0) : depth[r] - d);
}

// ------------------------------------------------------------------------------------------------
// planner support: n independent single steps (BAPOMDP::step in KeepCounts mode, BAPOMDP.cpp:111-143)
// from given (particle, domain state, action) triples — the in-tree steps of a wave of POMCP
// simulations (RBAPOUCT::traverseChanceNode, RBAPOUCT.cpp:249). One thread per request.
// ------------------------------------------------------------------------------------------------
template<bool REPLAY, bool LONG, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_step_batch(DevModel M, const float* counts, long long stride, const int* __restrict__ sid, long long n,
                 const long long* __restrict__ particle, const int* __restrict__ state,
                 const int* __restrict__ action, RngArgs ra, int* __restrict__ new_state,
                 int* __restrict__ obs, double* __restrict__ reward, int* __restrict__ terminal,
                 int* __restrict__ overrun)
{
    long long const r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    auto g            = RngOf<REPLAY>::make(ra, r);
    long long const p = particle[r];
    float* c          = const_cast<float*>(counts) + p * stride; // STEP_KEEP never writes
    int const a       = action[r];
    const Node* nodes = M.nodes + ((long long)sid[p] * M.A + a) * M.J;
    int o;
    Feat x2;
    int const s  = state[r];
    int const s2 = hyper_step<STEP_KEEP, decltype(g), false, LONG, SAMPLED>(M, nodes, c, s, g, o, x2, nullptr);
    bool term;
    reward[r]    = domain_reward(M, s, a, s2, term);
    new_state[r] = s2;
    obs[r]       = o;
    terminal[r]  = term ? 1 : 0;
    if (g.overrun) *overrun = 1;
}

__global__ void __launch_bounds__(kThreads)
    k_gather_states(const int* __restrict__ state, const long long* __restrict__ idx, long long n,
                    int* __restrict__ out)
{
    long long const j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = state[idx[j]];
}

// Belief::sample x n in one launch (native mode): weighted beliefs draw from the cdf, flat ones
// uniformly
__global__ void __launch_bounds__(kThreads)
    k_sample_batch(const double* __restrict__ cdf, long long N, long long n, RngArgs ra,
                   long long* __restrict__ out)
{
    long long const j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    auto g = RngOf<false>::make(ra, j);
    if (!cdf)
    {
        out[j] = draw_k(g, (uint32_t)N);
        return;
    }
    double const thr = draw_u(g) * cdf[N - 1];
    long long lo = 0, hi = N - 1;
    while (lo < hi)
    {
        long long const mid = (lo + hi) >> 1;
        if (cdf[mid] > thr) hi = mid;
        else
            lo = mid + 1;
    }
    out[j] = lo;
}

// ------------------------------------------------------------------------------------------------
// POMCP with the SEARCH TREE ON THE DEVICE: planners::RBAPOUCT::selectAction
// (RBAPOUCT.cpp:67-153) as waves of simulations, one thread per simulation, the whole simulation —
// root particle, UCB descent (traverseActionNode / traverseChanceNode, RBAPOUCT.cpp:197-277), leaf
// expansion, random-policy rollout (RBAPOUCT.cpp:295-323) and back-up — inside one kernel. The
// simulations of a wave share the tree through atomics:
//   * a tree node IS a slot of an open-addressing hash table keyed (parent node, action, observation);
//     claiming the slot with one atomicCAS creates the node, so there is no allocator and no waiting;
//   * UCB reads q = q_sum / n_done and the selection counts n_sel, which are incremented when an action
//     is CHOSEN (virtual visits): concurrent simulations spread over the actions instead of all taking
//     the same path; returns are added to q_sum / n_done when the simulation has finished
//     (ChanceNode::addVisit, MCTSTreeNodes.cpp:8-12: the mean of the returns).
// Root sampling (RBAPOUCT.cpp:86-106): the particle keeps its counts for the whole simulation
// (KeepCounts), only the domain state evolves. The tree differs from the sequential planner's only
// in the order in which simulations see each other's statistics.
// ------------------------------------------------------------------------------------------------
struct __align__(16) TreeStat // one 16-byte load per action in the UCB scan
{
    int n_sel;    // times the action was chosen (incl. simulations in flight)
    int n_done;   // returns backed up
    double q_sum; // sum of those returns
};

struct TreeArgs
{
    unsigned long long* keys; // [table]: (parent + 1) << 32 | (action * O + observation); EMPTY = ~0
    int* visits;              // [table + 1] selections through the node; node `table` is the root
    TreeStat* stat;           // [table + 1][A] per (node, action)
    unsigned int mask;        // table - 1 (table is a power of two)
    int root;                 // = table
    // belief
    const float* counts;
    long long stride;
    const int *sid, *state;
    const double* cdf; // weighted beliefs: inclusive cdf; NULL: flat filter, uniform pick
    long long N;
    const float* base; // base+delta storage
    long long base_stride;
    const int* proto_sid; // base+delta storage: prototype -> structure id
    // search
    int depth;
    double u, discount;
    long long first_sim, n_wave;
    int* path_node;   // [depth][wave]
    int* path_action; // [depth][wave]
    double* path_reward;
    int* overflow;    // set when the table was full (the simulation then ends in a rollout)
    // many runs at once (fba_runs_plan): R > 0. Thread t serves run t / w, its simulation
    // first_sim + t % w; run r owns particles [r run_n, (r+1) run_n), its root is node root + r, it
    // draws from Philox streams keyed seed + r — with w = 1 exactly what fba_tree_search(wave = 1)
    // does on a stand-alone belief seeded seed + r.
    int R, w;
    long long run_n, n_sims;
    const int* depth_r;          // [R] search depth per run (episodes are at different steps)
    const unsigned char* active; // [R] or NULL
};

__device__ __forceinline__ unsigned int tree_hash(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned int)k;
}

// Action choice in a tree node (RBAPOUCT::selectChanceNodeUCB, RBAPOUCT.cpp:162-205, with the table of
// RBAPOUCT.cpp:349-357): argmax_a q(a) + u sqrt(log(m + 1) / n(a)), +inf for n = 0, ties broken uniformly.
// The reference is sequential: visitor number m of a node sees the statistics of visitors 0..m-1. Here the
// simulations of a wave run concurrently, so a visitor first takes an atomic TICKET m = visits[node]++:
//   * tickets 0..A-1 take the A untried actions, one each, in a random order that is a function of the node
//     (sampling without replacement = what "uniformly among the +inf candidates" does sequentially) — no two
//     concurrent visitors can both see the same action as untried, whatever the thread timing;
//   * later tickets read the statistics through L2 (__ldcg: atomics land in L2, an L1 line could hold a
//     stale snapshot for the whole kernel and herd every visitor of one SM onto one action) and maximise
//     the bound with m = their ticket; n(a) counts selections incl. simulations still in flight (virtual
//     visits), q(a) is the mean of the returns backed up so far (0 while there is none, like an
//     unvisited ChanceNode's qValue()).
// With one simulation per wave this is the reference's algorithm exactly. (q_sum, n_done) are two atomics:
// a concurrent reader may see a mean that is off by one in-flight sample — transient, bounded, and absent
// when wave = 1.
__device__ __forceinline__ unsigned long long tree_mix(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// k-th element of a random permutation of 0..A-1 keyed by `key` (every visitor of a node derives the
// same permutation): sampling without replacement over a bitmask for A <= 32, a random affine map otherwise
__device__ __forceinline__ int tree_untried(unsigned long long key, int A, int k)
{
    if (A <= 32)
    {
        unsigned int left = (A == 32) ? 0xffffffffu : ((1u << A) - 1u);
        int pick = 0;
        for (int i = 0; i <= k; ++i)
        {
            key = tree_mix(key);
            unsigned int const r = (unsigned int)(((key >> 32) * (unsigned long long)(A - i)) >> 32);
            pick = (int)__fns(left, 0, (int)r + 1);
            left &= ~(1u << pick);
        }
        return pick;
    }
    key = tree_mix(key);
    int const off = (int)((key >> 33) % (unsigned long long)A);
    int step      = 1 + (int)((tree_mix(key) >> 33) % (unsigned long long)(A - 1));
    for (;; ++step)
    { // smallest step >= the drawn one that is coprime with A
        int x = A, y = step % A;
        while (y) { int const z = x % y; x = y, y = z; }
        if (x == 1) break;
    }
    return (int)(((long long)off + (long long)k * (step % A)) % A);
}

template<class R>
__device__ __forceinline__ int tree_ucb(const TreeArgs& T, int node, int A, R& g, unsigned long long perm_key)
{
    int const m = atomicAdd(&T.visits[node], 1); // ticket = visits before this one (ActionNode::visited())
    int pick    = 0;
    if (m < A) pick = tree_untried(perm_key, A, m);
    else
    {
        double const lg = log1p((double)m);
        double best     = -1.7976931348623157e308;
        int ties        = 0;
        const int4* row = reinterpret_cast<const int4*>(T.stat + (long long)node * A);
        // four actions' statistics per round trip
        for (int a0 = 0; a0 < A; a0 += 4)
        {
            int4 raw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (a0 + k < A) raw[k] = __ldcg(row + a0 + k);
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                int const a = a0 + k;
                if (a >= A) break;
                // every action has been handed out (m >= A); its holder's increment may still be on its way
                int const ns = max(raw[k].x, 1), nd = raw[k].y;
                double const qs = __hiloint2double(raw[k].w, raw[k].z);
                double const v  = ((nd > 0) ? qs / (double)nd : 0.0) + T.u * sqrt(lg / (double)ns);
                if (v > best) best = v, pick = a, ties = 1;
                else if (v == best && draw_k(g, (uint32_t)++ties) == 0)
                    pick = a;
            }
        }
    }
    atomicAdd(&T.stat[(long long)node * A + pick].n_sel, 1);
    return pick;
}

template<bool DELTA, bool LONG, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_pomcp_wave(DevModel M, TreeArgs T, RngArgs ra)
{
    long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T.n_wave) return;
    long long sim = T.first_sim + t, p0 = 0, np = T.N;
    int root = T.root, depth = T.depth;
    if (T.R > 0)
    {
        int const r = (int)(t / T.w);
        sim         = T.first_sim + t % T.w;
        if (sim >= T.n_sims || (T.active && !T.active[r])) return;
        ra.seed += (unsigned long long)r;
        p0 = (long long)r * T.run_n, np = T.run_n;
        root += r;
        if (T.depth_r) depth = T.depth_r[r];
    }
    auto g = RngOf<false>::make(ra, sim);
    // the order in which a node's first visitors try its actions: a function of (seed, search, the
    // action-observation path from the root to the node) — the same for every visitor of the node and
    // independent of where the hash table put it
    unsigned long long perm_key = tree_mix(ra.seed ^ (ra.offset * 0xd1342543de82ef95ull));
    // root particle: Belief::sample()
    long long p;
    if (T.cdf)
    {
        const double* cdf = T.cdf + p0;
        double const thr  = draw_u(g) * cdf[np - 1];
        long long lo = 0, hi = np - 1;
        while (lo < hi)
        {
            long long const mid = (lo + hi) >> 1;
            if (cdf[mid] > thr) hi = mid;
            else
                lo = mid + 1;
        }
        p = p0 + lo;
    } else
        p = p0 + draw_k(g, (uint32_t)np);
    float* c         = const_cast<float*>(T.counts) + p * T.stride; // KeepCounts: never written
    const Node* base = DELTA ? M.nodes : M.nodes + (long long)T.sid[p] * M.A * M.J;
    const float* tb  = DELTA ? T.base + (long long)T.sid[p] * T.base_stride : nullptr;
    int s    = T.state[p];
    int node = root, len = 0, d = depth;
    // ONE loop over simulated steps for both phases (in the tree: UCB action, path record, child
    // lookup; below the new leaf: random action, discounted return — RBAPOUCT::rollout), so that
    // the lanes of a warp, which are at different depths and phases of different simulations, still
    // run the expensive part — BAPOMDP::step — together instead of serialising on divergent code.
    bool in_tree = true;
    double leaf = 0.0, disc = 1.0;
    while (d > 0)
    {
        int a;
        if (in_tree)
        {
            a = tree_ucb(T, node, M.A, g, perm_key);
        } else
            a = random_action(M, g);
        int o, s2;
        if (DELTA)
            s2 = step_delta_particle<STEP_KEEP, SAMPLED>(M, T.proto_sid, T.sid[p], a, tb, reinterpret_cast<int*>(c), 0, s,
                                                         g, o, nullptr, nullptr, 0, nullptr);
        else
        {
            Feat x2;
            s2 = hyper_step<STEP_KEEP, decltype(g), false, LONG, SAMPLED>(M, base + (long long)a * M.J, c, s, g, o, x2,
                                                                          nullptr);
        }
        bool terminal;
        double const rew = domain_reward(M, s, a, s2, terminal);
        if (in_tree)
        {
            T.path_node[len * T.n_wave + t]   = node;
            T.path_action[len * T.n_wave + t] = a;
            T.path_reward[len * T.n_wave + t] = rew;
            ++len;
        } else
        { // Return::add / Discount::increment (Return.cpp:6-9, Discount.cpp:8-11)
            leaf = __dadd_rn(leaf, __dmul_rn(rew, disc));
            disc = __dmul_rn(disc, T.discount);
        }
        s = s2;
        --d;
        if (terminal) break; // delayed return 0 (RBAPOUCT.cpp:252)
        if (in_tree)
        { // the child for (node, a, o): find it, or create it by claiming a slot
            unsigned long long const key =
                ((unsigned long long)(node + 1) << 32) | (unsigned long long)((long long)a * M.O + o);
            unsigned int h = tree_hash(key) & T.mask;
            int child = -1;
            bool created = false;
            for (unsigned int probe = 0; probe <= T.mask; ++probe, h = (h + 1) & T.mask)
            {
                unsigned long long const seen = atomicCAS(&T.keys[h], ~0ull, key);
                if (seen == ~0ull)
                {
                    child   = (int)h;
                    created = true;
                    break;
                }
                if (seen == key)
                {
                    child = (int)h;
                    break;
                }
            }
            if (child < 0) *T.overflow = 1;
            // a new leaf is evaluated by a random-policy rollout (RBAPOUCT.cpp:258-266, 295-323):
            // the remaining steps of this loop
            if (created || child < 0) in_tree = false;
            else
            {
                node     = child;
                perm_key = tree_mix(perm_key ^ (key & 0xffffffffull));
            }
        }
    }
    // back up: ret = r + discount * delayed (RBAPOUCT.cpp:271-272)
    double ret = leaf;
    for (int k = len - 1; k >= 0; --k)
    {
        ret = T.path_reward[k * T.n_wave + t] + T.discount * ret;
        long long const cell = (long long)T.path_node[k * T.n_wave + t] * M.A + T.path_action[k * T.n_wave + t];
        atomicAdd(&T.stat[cell].q_sum, ret);
        atomicAdd(&T.stat[cell].n_done, 1);
    }
}

// ------------------------------------------------------------------------------------------------
// rejection sampling (RejectionSampling.hpp:26-72) as waves of independent attempts:
// attempt t picks a particle uniformly, simulates a step on it WITHOUT touching it and records the
// outcome; the accepted attempts, in attempt order, become the new particles.
// ------------------------------------------------------------------------------------------------
template<bool REPLAY, bool LONG, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_rs_attempt(DevModel M, const float* counts, long long stride, const int* __restrict__ state,
                 const int* __restrict__ sid, long long N, int a, int o, long long n_attempts,
                 RngArgs ra, int* __restrict__ src_out, int* __restrict__ state_out,
                 int* __restrict__ accept_out, int* __restrict__ rec_out, int* __restrict__ overrun)
{
    long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_attempts) return;
    auto g            = RngOf<REPLAY>::make(ra, t);
    int const i       = draw_k(g, (uint32_t)N); // FlatFilter::sample (FlatFilter.cpp:97-102)
    const Node* nodes = M.nodes + ((long long)sid[i] * M.A + a) * M.J;
    float* c          = const_cast<float*>(counts) + (long long)i * stride; // STEP_RECORD: read-only
    int rec[2 * FBA_MAX_FEATURES];
    int sim_o;
    Feat x2;
    int const s2 = hyper_step<STEP_RECORD, decltype(g), false, LONG, SAMPLED>(M, nodes, c, state[i], g, sim_o, x2, rec);
    src_out[t]    = i;
    state_out[t]  = s2;
    accept_out[t] = (sim_o == o) ? 1 : 0;
    for (int k = 0; k < M.J; ++k) rec_out[t * M.J + k] = rec[k];
    if (g.overrun) *overrun = 1;
}

// exclusive scan of the 0/1 accept flags over a wave (up to a few million attempts), three phases:
// per-tile counts, one-block scan of the tile counts (+ the total), per-tile scan with offset
constexpr int kFlagTile = 2048; // flags per CTA (8 per thread)

__global__ void __launch_bounds__(kThreads)
    k_flag_tile_counts(const int* __restrict__ flags, long long n, int* __restrict__ tile_count)
{
    __shared__ int sh[kThreads / 32];
    long long const base = (long long)blockIdx.x * kFlagTile;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kFlagTile / kThreads; ++k)
    {
        long long const i = base + threadIdx.x + k * kThreads;
        if (i < n) c += flags[i];
    }
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int t = 0;
        for (int k = 0; k < kThreads / 32; ++k) t += sh[k];
        tile_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kThreads)
    k_flag_scan_tiles(int* __restrict__ tile_count, int n_tiles, int* __restrict__ total)
{
    __shared__ int sh[kThreads];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += kThreads)
    {
        int const i = base + threadIdx.x;
        int const v = (i < n_tiles) ? tile_count[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kThreads; o <<= 1)
        {
            int const t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_tiles) tile_count[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) carry += sh[kThreads - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kThreads)
    k_flag_scan_apply(const int* __restrict__ flags, long long n, const int* __restrict__ tile_off,
                      int* __restrict__ pos)
{
    __shared__ int sh[kThreads];
    constexpr int PER    = kFlagTile / kThreads;
    long long const base = (long long)blockIdx.x * kFlagTile + (long long)threadIdx.x * PER;
    int v[PER];
    int run = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k)
    {
        long long const i = base + k;
        v[k]              = (i < n) ? flags[i] : 0;
        run += v[k];
    }
    sh[threadIdx.x] = run;
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1)
    {
        int const t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    int off = tile_off[blockIdx.x] + sh[threadIdx.x] - run;
#pragma unroll
    for (int k = 0; k < PER; ++k)
    {
        long long const i = base + k;
        if (i < n) pos[i] = off;
        off += v[k];
    }
}

// the attempt that produced the `need`-th acceptance of this wave: out[0] = its index + 1
__global__ void __launch_bounds__(kThreads)
    k_find_nth_accept(const int* __restrict__ accept, const int* __restrict__ pos, long long n, int need,
                      int* __restrict__ out)
{
    long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && accept[t] && pos[t] == need - 1) out[0] = (int)(t + 1);
}

// accepted attempt t with slot = already + pos[t] < N: copy its source block and apply the
// recorded +1 increments (one warp per attempt)
__global__ void __launch_bounds__(kThreads)
    k_rs_commit(const float* __restrict__ src, float* dst, long long stride,
                const int* __restrict__ src_sid, int* __restrict__ dst_sid, int* __restrict__ dst_state,
                const int* __restrict__ struct_size, long long N, int J, long long n_attempts,
                const int* __restrict__ att_src, const int* __restrict__ att_state,
                const int* __restrict__ accept, const int* __restrict__ pos, const int* __restrict__ rec,
                long long already, int delta, int delta_cap, int* __restrict__ overflow)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp0; t < n_attempts; t += nwarp)
    {
        if (!accept[t]) continue;
        long long const slot = already + pos[t];
        if (slot >= N) continue;
        long long const i = att_src[t];
        int const id      = src_sid[i];
        float* d          = dst + slot * stride;
        if (delta)
        { // base+delta: copy the increment list, append the two recorded cells
            const int* sb = reinterpret_cast<const int*>(src + i * stride);
            int* db       = reinterpret_cast<int*>(d);
            int const ne  = sb[0];
            warp_copy_block(src + i * stride, d, (ne + 1 + 3) >> 2, lane);
            __syncwarp();
            if (lane == 0)
            {
                if (ne + 2 <= delta_cap)
                {
                    db[1 + ne] = rec[t * J], db[2 + ne] = rec[t * J + 1];
                    db[0] = ne + 2;
                } else
                    *overflow = 2;
                dst_sid[slot]   = id;
                dst_state[slot] = att_state[t];
            }
            continue;
        }
        warp_copy_block(src + i * stride, d, (struct_size[id] + 3) >> 2, lane);
        __syncwarp();
        if (lane < J) d[rec[t * J + lane]] = __fadd_rn(src[i * stride + rec[t * J + lane]], 1.0f);
        if (lane == 0)
        {
            dst_sid[slot]   = id;
            dst_state[slot] = att_state[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// BABNModel::LogBDScore(prior) (BABNModel.cpp:451-478, DBNNode.cpp:82-117): log Bayesian-Dirichlet
// score of every particle's counts against prior counts of the same structure — what the reference's
// MCMC structure beliefs compare (MHNIPS2018.cpp:237-238, MHwithinGibbs.cpp:352,365).
//   score = sum over rows [ sum_v (lG(c_v) - lG(p_v)) + lG(sum_v p_v) - lG(sum_v c_v) ],
//   lG(x) = x < 1 ? 0 : lgamma(x)   (rnd::math::logGamma, random.cpp:127-135)
// One warp per particle; lanes take the rows of a node in turn, partial sums in double, one shuffle
// reduction at the end (so the sum order differs from the reference's: agreement to ~1e-12 relative).
// prior_n = 1: every particle is scored against the same prior particle; otherwise particle i against
// prior particle i. Structures must match (mismatch -> *flag = 1).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double log_gamma_ref(double x)
{
    return (x < 1.0) ? 0.0 : lgamma(x);
}

__global__ void __launch_bounds__(kThreads)
    k_log_bd_score(DevModel M, const float* __restrict__ counts, long long stride, const int* __restrict__ sid,
                   long long N, const float* __restrict__ prior, long long prior_stride,
                   const int* __restrict__ prior_sid, long long prior_n, double* __restrict__ score,
                   int* __restrict__ flag)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp0; i < N; i += nwarp)
    {
        long long const j = (prior_n == 1) ? 0 : i;
        int const id      = sid[i];
        if (lane == 0 && prior_sid[j] != id) *flag = 1;
        const float* c    = counts + i * stride;
        const float* p    = prior + j * prior_stride;
        const Node* nodes = M.nodes + (long long)id * M.A * M.J;
        double acc        = 0.0;
        // many small nodes (sysadmin-10: 220 nodes of at most 8 rows): a lane per NODE, walking its rows, keeps the
        // warp busy; few nodes: the lanes share the rows of one node at a time
        bool const by_node = M.A * M.J >= 32;
        for (int an = by_node ? lane : 0; an < M.A * M.J; an += by_node ? 32 : 1)
        {
            int const q     = an % M.J;
            Node const nd   = nodes[an];
            int const range = (q < M.FS) ? M.feat_s[q] : M.feat_o[q - M.FS];
            int cfgs        = 1;
            for (uint32_t m = nd.par; m; m &= m - 1) cfgs *= M.feat_s[__ffs(m) - 1];
            for (int r = by_node ? 0 : lane; r < cfgs; r += by_node ? 1 : 32)
            {
                int const row = nd.off + r * range;
                double ct = 0.0, pt = 0.0;
                for (int v = 0; v < range; ++v)
                {
                    double const cv = (double)c[row + v], pv = (double)p[row + v];
                    ct += cv, pt += pv;
                    acc += log_gamma_ref(cv) - log_gamma_ref(pv);
                }
                acc += log_gamma_ref(pt) - log_gamma_ref(ct);
            }
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) score[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Rejection sampling IN PLACE (PHILOX mode). The new belief is the first N accepted attempts; a flat
// filter has no order, so — as with in-place resampling — a source particle that was accepted at
// least once KEEPS ITS SLOT and takes the increments of its first accepted attempt there, and only the
// further accepted attempts of the same source are copies, into the slots of the sources that were
// never accepted. With sources drawn uniformly that is ~37 % of the blocks instead of all of them,
// and no second buffer. Deterministic: "first" is the smallest accepted index (atomicMin), extras
// and empty slots are matched in index order by two flag scans.
//   acc_*: the accepted attempts, compacted in attempt order, k = 0 .. N-1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    k_rs_collect(long long N, int J, long long n_attempts, const int* __restrict__ att_src,
                 const int* __restrict__ att_state, const int* __restrict__ accept, const int* __restrict__ pos,
                 const int* __restrict__ rec, long long already, int* __restrict__ acc_src,
                 int* __restrict__ acc_state, int* __restrict__ acc_rec)
{
    long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_attempts || !accept[t]) return;
    long long const k = already + pos[t];
    if (k >= N) return;
    acc_src[k]   = att_src[t];
    acc_state[k] = att_state[t];
    for (int j = 0; j < J; ++j) acc_rec[k * J + j] = rec[t * J + j];
}

__global__ void __launch_bounds__(kThreads) k_fill_int(int* __restrict__ p, long long n, int v)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// first[i] = the smallest accepted attempt whose source is particle i (INT_MAX: never accepted)
__global__ void __launch_bounds__(kThreads)
    k_rs_first(const int* __restrict__ acc_src, long long N, int* __restrict__ first)
{
    long long const k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) atomicMin(&first[acc_src[k]], (int)k);
}

// extra[k] = attempt k is NOT the first of its source (it needs a copy); empty[i] = slot i has no
// accepted attempt (it receives one)
__global__ void __launch_bounds__(kThreads)
    k_rs_flags(const int* __restrict__ acc_src, const int* __restrict__ first, long long N,
               int* __restrict__ extra, int* __restrict__ empty)
{
    long long const k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    extra[k] = first[acc_src[k]] != (int)k;
    empty[k] = first[k] == 0x7fffffff;
}

__global__ void __launch_bounds__(kThreads)
    k_rs_empty_list(const int* __restrict__ empty, const int* __restrict__ empty_pos, long long N,
                    int* __restrict__ empty_slot)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && empty[i]) empty_slot[empty_pos[i]] = (int)i;
}

// the e-th extra attempt becomes the e-th empty slot: copy of its (still untouched) source block plus
// its own increments. One warp per accepted attempt.
__global__ void __launch_bounds__(kThreads)
    k_rs_place_extras(float* counts, long long stride, int* sid, int* state,
                      const int* __restrict__ struct_size, long long N, int J, const int* __restrict__ acc_src,
                      const int* __restrict__ acc_state, const int* __restrict__ acc_rec,
                      const int* __restrict__ extra, const int* __restrict__ extra_pos,
                      const int* __restrict__ empty_slot)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = warp0; k < N; k += nwarp)
    {
        if (!extra[k]) continue;
        long long const i = acc_src[k], j = empty_slot[extra_pos[k]];
        int const id      = sid[i];
        const float* sb   = counts + i * stride;
        float* d          = counts + j * stride;
        warp_copy_block(sb, d, (struct_size[id] + 3) >> 2, lane);
        __syncwarp();
        if (lane < J) d[acc_rec[k * J + lane]] = __fadd_rn(sb[acc_rec[k * J + lane]], 1.0f);
        if (lane == 0)
        {
            sid[j]   = id;
            state[j] = acc_state[k];
        }
    }
}

// afterwards: every accepted source takes its FIRST accepted attempt's increments where it is
__global__ void __launch_bounds__(kThreads)
    k_rs_apply_first(float* counts, long long stride, int* __restrict__ state, long long N, int J,
                     const int* __restrict__ first, const int* __restrict__ acc_state,
                     const int* __restrict__ acc_rec)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int const k = first[i];
    if (k == 0x7fffffff) return;
    float* c = counts + i * stride;
    for (int j = 0; j < J; ++j)
    {
        int const cell = acc_rec[(long long)k * J + j];
        c[cell]        = __fadd_rn(c[cell], 1.0f);
    }
    state[i] = acc_state[k];
}

// ------------------------------------------------------------------------------------------------
// reinvigoration: BABNModel::marginalizeOut (BABNModel.cpp:205-229) of a fully connected counts
// donor onto a mutated structure, written into the replaced slot. One block per bred particle.
// Each destination cell sums its source rows in ascending source-configuration order — the order
// DBNNode::marginalizeOut (DBNNode.cpp:40-80) adds them — so the float result is identical.
// ------------------------------------------------------------------------------------------------
struct BreedJob
{
    int fc_index;   // counts donor in the fully connected belief
    int fc_struct;  // its structure id
    int new_struct; // the mutated structure id
    int slot;       // destination slot in the belief
    int state;      // domain state of the structure donor
};

constexpr int kBreedCache = 8192; // source configurations whose projection is cached per node

__global__ void __launch_bounds__(kThreads)
    k_breed(DevModel M, const float* __restrict__ fc_counts, long long fc_stride, float* dst_counts,
            long long dst_stride, int* __restrict__ dst_state, int* __restrict__ dst_sid,
            const BreedJob* __restrict__ jobs)
{
    __shared__ short proj_of[kBreedCache];
    BreedJob const job = jobs[blockIdx.x];
    const float* src   = fc_counts + (long long)job.fc_index * fc_stride;
    float* dst         = dst_counts + (long long)job.slot * dst_stride;
    const Node* sn     = M.nodes + (long long)job.fc_struct * M.A * M.J;
    const Node* dn     = M.nodes + (long long)job.new_struct * M.A * M.J;

    for (int an = 0; an < M.A * M.J; ++an)
    {
        int const j     = an % M.J;
        int const range = (j < M.FS) ? M.feat_s[j] : M.feat_o[j - M.FS];
        Node const s = sn[an], d = dn[an];
        int n_src = 1, n_dst = 1;
        for (int f = 0; f < M.FS; ++f)
        {
            if (s.par & (1u << f)) n_src *= M.feat_s[f];
            if (d.par & (1u << f)) n_dst *= M.feat_s[f];
        }
        if (s.par == d.par)
        {
            for (int k = threadIdx.x; k < n_src * range; k += blockDim.x) dst[d.off + k] = src[s.off + k];
            continue;
        }
        // projection of every source configuration onto the destination parents, once per node
        // (shared memory when it fits, recomputed per use otherwise)
        auto project = [&](int cfg) {
            int rem = cfg, proj = 0, mult = 1;
            for (int f = M.FS - 1; f >= 0; --f)
            {
                if (!(s.par & (1u << f))) continue;
                int const xv = rem % M.feat_s[f];
                rem /= M.feat_s[f];
                if (d.par & (1u << f))
                {
                    proj += xv * mult;
                    mult *= M.feat_s[f];
                }
            }
            return proj;
        };
        bool const cached = n_src <= kBreedCache;
        __syncthreads();
        if (cached)
            for (int cfg = threadIdx.x; cfg < n_src; cfg += blockDim.x) proj_of[cfg] = (short)project(cfg);
        __syncthreads();
        for (int cell = threadIdx.x; cell < n_dst * range; cell += blockDim.x)
        {
            int const dcfg = cell / range, v = cell - dcfg * range;
            float acc = 0.0f;
            for (int cfg = 0; cfg < n_src; ++cfg)
            {
                int const proj = cached ? (int)proj_of[cfg] : project(cfg);
                if (proj == dcfg) acc = __fadd_rn(acc, src[s.off + cfg * range + v]);
            }
            dst[d.off + cell] = acc;
        }
    }
    // zero the padding up to the stride so downloads compare equal
    int const used = M.struct_size[job.new_struct];
    for (long long k = used + threadIdx.x; k < dst_stride; k += blockDim.x) dst[k] = 0.0f;
    if (threadIdx.x == 0)
    {
        dst_state[job.slot] = job.state;
        dst_sid[job.slot]   = job.new_struct;
    }
}

// ------------------------------------------------------------------------------------------------
// base+delta storage (tabular models too large for dense private blocks): the per-particle kernels
// ------------------------------------------------------------------------------------------------

// Belief::initiate: empty increment lists; sid[i] = which base table (prior prototype) particle i uses
__global__ void __launch_bounds__(kThreads)
    k_init_delta(float* __restrict__ blocks, long long stride, int* __restrict__ state, int* __restrict__ sid,
                 double* __restrict__ w, long long N, const int* __restrict__ particle_proto,
                 const int* __restrict__ particle_state, int header)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    reinterpret_cast<int*>(blocks + i * stride)[0] = header; // words following the first: 0 (tabular) / 3 (journal)
    sid[i] = particle_proto ? particle_proto[i] : 0;
    if (particle_state) state[i] = particle_state[i];
    if (w) w[i] = 1.0 / (double)N;
}

// base + delta / base + journal  ->  dense private blocks (fba_belief_compact): one warp per particle copies its
// base table and then replays its own increments onto the copy — every increment is exactly +1.0f, so atomic adds in
// any order give the float sums the sequential replay gives. The particle's sid (a base-table index) becomes the
// structure id the dense kernels expect.
__global__ void __launch_bounds__(kThreads)
    k_compact(const float* __restrict__ base, long long base_stride, const float* __restrict__ blocks, long long stride,
              int* __restrict__ sid, const int* __restrict__ proto_sid, float* __restrict__ dense, long long N,
              int journal_J /* 0: tabular delta lists */)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp0; i < N; i += nwarp)
    {
        int const proto = sid[i];
        float* out      = dense + i * base_stride;
        warp_copy_block<4>(base + (long long)proto * base_stride, out, (int)(base_stride >> 2), lane);
        for (int k = (int)(base_stride & ~3ll) + lane; k < base_stride; k += 32) out[k] = base[(long long)proto * base_stride + k];
        __syncwarp();
        const int* blk = reinterpret_cast<const int*>(blocks + i * stride);
        if (journal_J == 0)
            for (int e = 1 + lane; e <= blk[0]; e += 32) atomicAdd(out + blk[e], 1.0f);
        else
        {
            int const Jp = (journal_J + 1 + 3) & ~3, nu = (blk[0] - (kJournalHeader - 1)) / Jp;
            for (int k = lane; k < nu * journal_J; k += 32)
            {
                int const u = k / journal_J, j = k - u * journal_J;
                atomicAdd(out + blk[kJournalHeader + u * Jp + j], 1.0f);
            }
        }
        __syncwarp();
        if (lane == 0) sid[i] = proto_sid[proto];
    }
}

template<bool REPLAY, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_propose_delta(DevModel M, const float* __restrict__ base, long long base_stride, float* blocks,
                    long long stride, int cap, int* __restrict__ state, const int* __restrict__ sid,
                    const int* __restrict__ proto_sid, double* __restrict__ w, long long N, int a, int o, RngArgs ra,
                    int* __restrict__ overrun)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto g          = RngOf<REPLAY>::make(ra, i);
    int const proto = sid[i];
    const float* tb = base + (long long)proto * base_stride;
    int* block      = reinterpret_cast<int*>(blocks + i * stride);
    int sim_o;
    double prob;
    int const s2 = step_delta_particle<STEP_UPDATE, SAMPLED>(M, proto_sid, proto, a, tb, block, cap, state[i], g, sim_o,
                                                             nullptr, overrun, o, &prob);
    state[i] = s2;
    w[i]     = __dmul_rn(w[i], prob);
    if (g.overrun) *overrun = 1;
}

// importance-sampling update of a FACTORED journal belief (no tabular row buffers on the stack): the
// general step, one thread per particle reading its own journal.
template<bool REPLAY, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_propose_journal(DevModel M, const float* __restrict__ base, long long base_stride, float* blocks,
                      long long stride, int cap, int* __restrict__ state, const int* __restrict__ sid,
                      const int* __restrict__ proto_sid, double* __restrict__ w, long long N, int a, int o, RngArgs ra,
                      int* __restrict__ overrun)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto g            = RngOf<REPLAY>::make(ra, i);
    int const proto   = sid[i];
    const float* tb   = base + (long long)proto * base_stride;
    int* block        = reinterpret_cast<int*>(blocks + i * stride);
    const Node* nodes = M.nodes + ((long long)proto_sid[proto] * M.A + a) * M.J;
    double prob;
    int sim_o;
    int const s2 = hyper_step_journal<STEP_UPDATE, SAMPLED>(M, nodes, tb, block, cap, state[i], g, sim_o, nullptr, overrun,
                                                            o, &prob, a);
    state[i] = s2;
    w[i]     = __dmul_rn(w[i], prob);
    if (g.overrun) *overrun = 1;
}

// The same for models whose features are all binary (JournalBinaryStep), journals STAGED THROUGH SHARED
// MEMORY. A thread walking its own journal touches a different 4.6 KB region than its neighbours: every
// load instruction of the warp costs 32 sectors in 32 DRAM pages and the walk is latency-bound (measured:
// 1.3 ms at 35 updates per particle, 4x the bytes' worth). Here a warp owns 32 consecutive particles and
// reads their journals TOGETHER: for each particle in turn the 32 lanes fetch 32 consecutive int4 vectors
// (one 512-byte coalesced request, 32 of them in flight per lane) into a shared-memory row; then every
// lane walks its own particle's row (rows padded to 33 vectors: conflict-free 128-bit reads). All
// particles of a belief hold the same number of updates nu (one per update, copies keep it), which the
// host knows, so there is no divergence in the staging loop. Appends are three 16-byte stores per particle.
constexpr int kStageWarps = 4;
constexpr int kStageVec   = 16; // int4 vectors of one particle's journal per staged chunk
template<bool REPLAY>
__global__ void __launch_bounds__(kStageWarps * 32)
    k_propose_journal_staged(DevModel M, const float* __restrict__ base, long long base_stride, float* blocks,
                             long long stride, int cap, int* __restrict__ state, const int* __restrict__ sid,
                             const int* __restrict__ proto_sid, double* __restrict__ w, long long N, int a, int o,
                             RngArgs ra, int* __restrict__ overrun, int nu)
{
    __shared__ int4 stage[kStageWarps][32][kStageVec + 1];
    int const lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    long long const first = ((long long)blockIdx.x * kStageWarps + wib) * 32; // the warp's first particle
    if (first >= N) return;
    long long const i = first + lane;
    bool const valid  = i < N;
    int const J = M.J, Jp = journal_padded(J), nvec = Jp >> 2;
    int const per_chunk = kStageVec / nvec; // whole updates per staged chunk
    JournalBinaryStep st;
    const Node* nodes = nullptr;
    const float* tb   = nullptr;
    if (valid)
    {
        int const proto = sid[i];
        tb              = base + (long long)proto * base_stride;
        nodes           = M.nodes + ((long long)proto_sid[proto] * M.A + a) * M.J;
        st.begin(M, nodes, state[i], a);
    }
    int const n_here = (int)min(32ll, N - first);
    for (int u0 = 0; u0 < nu; u0 += per_chunk)
    {
        int const n  = min(per_chunk, nu - u0);
        int const nv = n * nvec;
        if (lane < nv) // asynchronous 16-byte copies global -> shared (LDGSTS): all n_here requests of a lane in flight
            for (int p = 0; p < n_here; ++p)
                __pipeline_memcpy_async(
                    &stage[wib][p][lane],
                    reinterpret_cast<const int4*>(blocks + (first + p) * stride + kJournalHeader) + u0 * nvec + lane,
                    sizeof(int4));
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncwarp();
        if (valid) st.add(M, &stage[wib][lane][0], n, nvec);
        __syncwarp();
    }
    if (!valid) return;
    auto g = RngOf<REPLAY>::make(ra, i);
    int inc[FBA_MAX_FEATURES + 4];
    double prob;
    int const s2 = st.finish(M, nodes, tb, g, o, inc, &prob);
    int* block   = reinterpret_cast<int*>(blocks + i * stride);
    if ((nu + 1) * J <= cap)
    {
        int4* dst = reinterpret_cast<int4*>(block + kJournalHeader + nu * Jp);
#pragma unroll
        for (int k = 0; k < (FBA_MAX_FEATURES + 4 + 3) / 4; ++k)
        {
            if (k >= nvec) break;
            int4 v;
            v.x = (4 * k + 0 < J) ? inc[4 * k + 0] : -1;
            v.y = (4 * k + 1 < J) ? inc[4 * k + 1] : -1;
            v.z = (4 * k + 2 < J) ? inc[4 * k + 2] : -1;
            v.w = (4 * k + 3 < J) ? inc[4 * k + 3] : (4 * k + 3 == Jp - 1) ? journal_tag(a) : -1;
            dst[k] = v;
        }
        block[0] = (kJournalHeader - 1) + (nu + 1) * Jp;
    } else
        *overrun = 2;
    state[i] = s2;
    w[i]     = __dmul_rn(w[i], prob);
    if (g.overrun) *overrun = 1;
}

template<bool REPLAY, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_rollouts_delta(DevModel M, const float* __restrict__ base, long long base_stride, const float* blocks,
                     long long stride, const int* __restrict__ sid, long long n,
                     const long long* __restrict__ particle, const int* __restrict__ start,
                     const int* __restrict__ depth, double discount, RngArgs ra, double* __restrict__ ret_out,
                     int* __restrict__ overrun, unsigned long long* __restrict__ steps_done,
                     const int* __restrict__ proto_sid)
{
    long long const r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    auto g            = RngOf<REPLAY>::make(ra, r);
    long long const p = particle[r];
    const float* tb   = base + (long long)sid[p] * base_stride;
    int* block        = reinterpret_cast<int*>(const_cast<float*>(blocks) + p * stride); // read-only here
    double ret = 0.0, disc = 1.0;
    int s = start[r], d = depth[r];
    bool terminal = false;
    while (d > 0 && !terminal)
    {
        int const a = random_action(M, g);
        int o;
        int const s2 = step_delta_particle<STEP_KEEP, SAMPLED>(M, proto_sid, sid[p], a, tb, block, 0, s, g, o, nullptr,
                                                               nullptr, 0, nullptr);
        double const rew = domain_reward(M, s, a, s2, terminal);
        ret  = __dadd_rn(ret, __dmul_rn(rew, disc));
        disc = __dmul_rn(disc, discount);
        s    = s2;
        --d;
    }
    ret_out[r] = ret;
    if (g.overrun) *overrun = 1;
    count_steps(steps_done, depth[r] - d);
}

template<bool REPLAY, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_step_batch_delta(DevModel M, const float* __restrict__ base, long long base_stride, const float* blocks,
                       long long stride, const int* __restrict__ sid, long long n,
                       const long long* __restrict__ particle, const int* __restrict__ state,
                       const int* __restrict__ action, RngArgs ra, int* __restrict__ new_state,
                       int* __restrict__ obs, double* __restrict__ reward, int* __restrict__ terminal,
                       int* __restrict__ overrun, const int* __restrict__ proto_sid)
{
    long long const r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    auto g            = RngOf<REPLAY>::make(ra, r);
    long long const p = particle[r];
    int const a       = action[r];
    int o;
    int const s  = state[r];
    int const s2 = step_delta_particle<STEP_KEEP, SAMPLED>(
        M, proto_sid, sid[p], a, base + (long long)sid[p] * base_stride,
        reinterpret_cast<int*>(const_cast<float*>(blocks) + p * stride), 0, s, g, o, nullptr, nullptr, 0, nullptr);
    bool term;
    reward[r]    = domain_reward(M, s, a, s2, term);
    new_state[r] = s2;
    obs[r]       = o;
    terminal[r]  = term ? 1 : 0;
    if (g.overrun) *overrun = 1;
}

template<bool REPLAY, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_rs_attempt_delta(DevModel M, const float* __restrict__ base, long long base_stride, const float* blocks,
                       long long stride, const int* __restrict__ state, const int* __restrict__ sid, long long N,
                       int a, int o, long long n_attempts, RngArgs ra, int* __restrict__ src_out,
                       int* __restrict__ state_out, int* __restrict__ accept_out, int* __restrict__ rec_out,
                       int* __restrict__ overrun)
{
    long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_attempts) return;
    auto g      = RngOf<REPLAY>::make(ra, t);
    int const i = draw_k(g, (uint32_t)N);
    int rec[2];
    int sim_o;
    int const s2 = hyper_step_delta<STEP_RECORD, SAMPLED>(
        M, M.nodes + (long long)a * M.J, base + (long long)sid[i] * base_stride,
        reinterpret_cast<int*>(const_cast<float*>(blocks) + (long long)i * stride), 0, state[i], g, sim_o, rec,
        nullptr);
    src_out[t]         = i;
    state_out[t]       = s2;
    accept_out[t]      = (sim_o == o) ? 1 : 0;
    rec_out[t * 2]     = rec[0];
    rec_out[t * 2 + 1] = rec[1];
    if (g.overrun) *overrun = 1;
}


// ------------------------------------------------------------------------------------------------
// MH structure beliefs: computePosterior (MHNIPS2018.cpp:41-109) — the whole
// (action, observation) history replayed on every PROPOSAL particle, one thread per proposal. Per
// episode attempt: a domain start state, then per step s' and o sampled from the particle's own counts
// (expected mode whatever the simulator's, MHNIPS2018.cpp:47); a wrong observation abandons the attempt
// — the cells incremented so far in this episode get -1, in step order — and the episode is tried again;
// a right one gets its +1 on the cells hyper_step<STEP_RECORD> recorded (the simulated observation IS
// the observed one then). state[i] = the state after the last step. rec: per-particle scratch of the
// current attempt's cells, max_len x J ints.
// ------------------------------------------------------------------------------------------------
struct HistoryArgs
{
    int n_episodes, max_len;
    const int* episode_len;
    const int* actions;
    const int* observations;
    long long max_attempts;
};

template<bool REPLAY, bool LONG>
__global__ void __launch_bounds__(kThreads)
    k_mh_replay(DevModel M, float* counts, long long stride, int* __restrict__ state, const int* __restrict__ sid,
                long long N, HistoryArgs H, RngArgs ra, int* __restrict__ rec_all, int* __restrict__ failed,
                int* __restrict__ overrun, long long* __restrict__ words_used)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto g            = RngOf<REPLAY>::make(ra, i);
    float* c          = counts + i * stride;
    const Node* nodes = M.nodes + (long long)sid[i] * M.A * M.J;
    int* rec          = rec_all + i * (long long)H.max_len * M.J;
    long long attempts = 0;
    int new_s = 0, first = 0;
    for (int e = 0; e < H.n_episodes; ++e)
    {
        int const len = H.episode_len[e];
        for (;;)
        {
            if (++attempts > H.max_attempts)
            {
                *failed  = 1;
                state[i] = new_s;
                return;
            }
            int s       = sample_start_state(M, g);
            int applied = 0;
            for (int t = 0; t < len; ++t)
            {
                int const a = H.actions[first + t];
                int o;
                Feat x2;
                int* r = rec + applied * M.J;
                int const s2 =
                    hyper_step<STEP_RECORD, decltype(g), false, LONG, false>(M, nodes + (long long)a * M.J, c, s, g, o, x2, r);
                new_s = s2;
                if (o != H.observations[first + t]) break;
                for (int k = 0; k < M.J; ++k) c[r[k]] = __fadd_rn(c[r[k]], 1.0f);
                ++applied;
                s = s2;
            }
            if (applied == len) break;
            for (int t = 0; t < applied; ++t)
                for (int k = 0; k < M.J; ++k) c[rec[t * M.J + k]] = __fadd_rn(c[rec[t * M.J + k]], -1.0f);
        }
        first += len;
    }
    state[i] = new_s;
    if (g.overrun) *overrun = 1;
    if (words_used) words_used[i] = rng_words_used(g, i * ra.words_per_item);
}

// ------------------------------------------------------------------------------------------------
// beliefs::bayes_adaptive::NestedBelief::updateEstimation (src/beliefs/bayes-adaptive/NestedBelief.cpp:129-193):
// a weighted TOP filter of count blocks, each with its own flat BOTTOM filter of n_bottom domain states.
// One thread per top particle runs the reference's loop as it is written — it is sequential by nature:
// every accepted bottom particle raises the counts by 1 / n_bottom before the next attempt samples from
// them. Per attempt: a uniformly drawn bottom state (FlatFilter::sample), BAPOMDP::step in KeepCounts mode,
// accept iff the simulated observation is the real one, then incrementCountsOf(old, a, o, new, 1 / n_bottom);
// after n_bottom acceptances the particle's weight is multiplied by 1 / attempts. The top filter is
// normalised by the caller (sequential sums, a few thousand doubles at most).
// ------------------------------------------------------------------------------------------------
template<bool REPLAY, bool LONG, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_nested_update(DevModel M, float* counts, long long stride, const int* __restrict__ sid, double* __restrict__ w,
                    long long n_top, int n_bottom, const int* __restrict__ states_in, int* __restrict__ states_out,
                    int action, int observation, float amount, long long max_attempts, RngArgs ra,
                    long long* __restrict__ attempts_out, int* __restrict__ failed, int* __restrict__ overrun)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_top) return;
    auto g            = RngOf<REPLAY>::make(ra, i);
    float* c          = counts + i * stride;
    const Node* nodes = M.nodes + ((long long)sid[i] * M.A + action) * M.J;
    const int* in     = states_in + i * (long long)n_bottom;
    int* out          = states_out + i * (long long)n_bottom;
    int rec[2 * FBA_MAX_FEATURES];
    long long count = 0;
    int accepted    = 0;
    while (accepted < n_bottom)
    {
        if (count >= max_attempts || g.overrun)
        {
            *failed = 1;
            break;
        }
        int const s = in[draw_k(g, (uint32_t)n_bottom)];
        int o;
        Feat x2;
        int const s2 = hyper_step<STEP_RECORD, decltype(g), false, LONG, SAMPLED>(M, nodes, c, s, g, o, x2, rec);
        if (o == observation)
        {
            out[accepted++] = s2;
            for (int k = 0; k < M.J; ++k) c[rec[k]] = __fadd_rn(c[rec[k]], amount);
        }
        ++count;
    }
    // _filter.particle(i)->w *= 1.0 / static_cast<double>(count)  (NestedBelief.cpp:183)
    if (count > 0) w[i] = __dmul_rn(w[i], __ddiv_rn(1.0, (double)count));
    if (attempts_out) attempts_out[i] = count;
    if (g.overrun) *overrun = 1;
}

// PHILOX mode, one WARP per top particle: the 32 lanes make one attempt each per round, all reading the counts as
// they stand at the start of the round; the acceptances of a round — the first ones in lane order, up to what the
// bottom filter still needs — then land together (atomic adds of the same 1 / n_bottom, so the float sums do not
// depend on their order), and the attempt count is the position of the last one taken. Against the loop above the
// only difference is that an attempt does not see the (at most 31) increments of 1 / n_bottom made earlier in its own
// round: a relative change of the sampled rows of order 32 / (n_bottom x row total), invisible at any sample size this
// belief is run with; REPLAY mode and option "nested_exact" keep the thread-per-particle loop.
template<bool LONG, bool SAMPLED>
__global__ void __launch_bounds__(kThreads)
    k_nested_update_warp(DevModel M, float* counts, long long stride, const int* __restrict__ sid, double* __restrict__ w,
                         long long n_top, int n_bottom, const int* __restrict__ states_in, int* __restrict__ states_out,
                         int action, int observation, float amount, long long max_attempts, RngArgs ra,
                         long long* __restrict__ attempts_out, int* __restrict__ failed)
{
    long long const i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int const lane    = threadIdx.x & 31;
    if (i >= n_top) return;
    PhiloxRng g(ra.seed, ra.stream_base + (unsigned long long)i * 32ull + (unsigned long long)lane, ra.offset);
    float* c          = counts + i * stride;
    const Node* nodes = M.nodes + ((long long)sid[i] * M.A + action) * M.J;
    const int* in     = states_in + i * (long long)n_bottom;
    int* out          = states_out + i * (long long)n_bottom;
    int rec[2 * FBA_MAX_FEATURES];
    long long count = 0;
    int accepted    = 0;
    while (accepted < n_bottom)
    {
        if (count >= max_attempts)
        {
            if (lane == 0) *failed = 1;
            break;
        }
        int const s = in[draw_k(g, (uint32_t)n_bottom)];
        int o;
        Feat x2;
        int const s2 = hyper_step<STEP_RECORD, PhiloxRng, false, LONG, SAMPLED>(M, nodes, c, s, g, o, x2, rec);
        unsigned const ok   = __ballot_sync(0xffffffffu, o == observation);
        int const need      = n_bottom - accepted;
        int const my_rank   = __popc(ok & ((1u << lane) - 1u));
        bool const taken    = (o == observation) && my_rank < need;
        int const n_ok      = __popc(ok);
        __syncwarp(); // every lane has read the round's counts before any of them changes
        if (taken)
        {
            out[accepted + my_rank] = s2;
            for (int k = 0; k < M.J; ++k) atomicAdd(c + rec[k], amount);
        }
        if (n_ok >= need)
        { // the attempt that filled the filter is the need-th accepted lane of this round
            unsigned m = ok;
            for (int k = 1; k < need; ++k) m &= m - 1;
            count += __ffs(m);
            accepted = n_bottom;
        } else
        {
            count += 32;
            accepted += n_ok;
        }
        __syncwarp();
        __threadfence_block();
    }
    if (lane == 0)
    {
        if (count > 0) w[i] = __dmul_rn(w[i], __ddiv_rn(1.0, (double)count));
        if (attempts_out) attempts_out[i] = count;
    }
}

// ------------------------------------------------------------------------------------------------
// MHwithinGibbs (src/beliefs/bayes-adaptive/factored/MHwithinGibbs.cpp): state histories conditioned on a
// model and the (action, observation) history, and the posterior counts of a state history.
// ------------------------------------------------------------------------------------------------

// rejectionSampleStateHistory (:38-94), one thread per particle (model): per episode a start state, then s'
// and o from the particle's counts (expected Dirichlets, counts untouched) step after step; the first wrong
// observation abandons the attempt. states_out: per particle (n_steps + n_episodes) states.
template<bool REPLAY, bool LONG>
__global__ void __launch_bounds__(kThreads)
    k_state_history_rs(DevModel M, const float* __restrict__ counts, long long stride, const int* __restrict__ sid,
                       long long N, HistoryArgs H, RngArgs ra, int* __restrict__ states_out, long long out_stride,
                       int* __restrict__ failed, int* __restrict__ overrun, long long* __restrict__ words_used)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto g            = RngOf<REPLAY>::make(ra, i);
    float* c          = const_cast<float*>(counts) + i * stride; // STEP_KEEP never writes
    const Node* nodes = M.nodes + (long long)sid[i] * M.A * M.J;
    int* out          = states_out + i * out_stride;
    long long attempts = 0;
    int first = 0, pos = 0;
    for (int e = 0; e < H.n_episodes; ++e)
    {
        int const len = H.episode_len[e];
        for (;;)
        {
            if (++attempts > H.max_attempts || g.overrun)
            {
                *failed = 1;
                return;
            }
            int s    = sample_start_state(M, g);
            out[pos] = s;
            int t    = 0;
            for (; t < len; ++t)
            {
                int o;
                Feat x2;
                s = hyper_step<STEP_KEEP, decltype(g), false, LONG, false>(
                    M, nodes + (long long)H.actions[first + t] * M.J, c, s, g, o, x2, nullptr);
                if (o != H.observations[first + t]) break;
                out[pos + 1 + t] = s;
            }
            if (t == len) break;
        }
        first += len;
        pos += len + 1;
    }
    if (g.overrun) *overrun = 1;
    if (words_used) words_used[i] = rng_words_used(g, i * ra.words_per_item);
}

// The same in PHILOX mode with ONE WARP per particle: the attempts of an episode are independent draws, so the 32
// lanes make one attempt each per round, each from its own Philox stream, and the lowest lane that reproduced the
// observations supplies the episode's states — the same distribution as attempting one after the other, but a
// model that explains the history badly (attempts are geometric, the tail is heavy) costs 1/32 of the rounds, which
// is what a lockstep sweep over many chains waits for. scratch: per lane (max_len + 1) states.
template<bool LONG>
__global__ void __launch_bounds__(kThreads)
    k_state_history_rs_warp(DevModel M, const float* __restrict__ counts, long long stride, const int* __restrict__ sid,
                            long long N, HistoryArgs H, RngArgs ra, int* __restrict__ states_out, long long out_stride,
                            int* __restrict__ scratch, int* __restrict__ failed)
{
    long long const i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int const lane    = threadIdx.x & 31;
    if (i >= N) return;
    PhiloxRng g(ra.seed, ra.stream_base + (unsigned long long)i * 32ull + (unsigned long long)lane, ra.offset);
    float* c          = const_cast<float*>(counts) + i * stride; // STEP_KEEP never writes
    const Node* nodes = M.nodes + (long long)sid[i] * M.A * M.J;
    int* out          = states_out + i * out_stride;
    int* mine         = scratch + (i * 32 + lane) * (long long)(H.max_len + 1);
    long long attempts = 0;
    int first = 0, pos = 0;
    for (int e = 0; e < H.n_episodes; ++e)
    {
        int const len = H.episode_len[e];
        for (;;)
        {
            attempts += 32;
            if (attempts > H.max_attempts + 31)
            {
                if (lane == 0) *failed = 1;
                return;
            }
            int s   = sample_start_state(M, g);
            mine[0] = s;
            int t   = 0;
            for (; t < len; ++t)
            {
                int o;
                Feat x2;
                s = hyper_step<STEP_KEEP, PhiloxRng, false, LONG, false>(M, nodes + (long long)H.actions[first + t] * M.J, c, s,
                                                                         g, o, x2, nullptr);
                if (o != H.observations[first + t]) break;
                mine[1 + t] = s;
            }
            unsigned const ok = __ballot_sync(0xffffffffu, t == len);
            if (ok)
            {
                int const winner  = __ffs(ok) - 1;
                const int* theirs = scratch + (i * 32 + winner) * (long long)(H.max_len + 1);
                __syncwarp();
                for (int k = lane; k <= len; k += 32) out[pos + k] = theirs[k];
                break;
            }
        }
        first += len;
        pos += len + 1;
    }
}

// BABNModel::flattenT / flattenO (BABNModel.cpp:89-178) of every particle into scratch, laid out for the
// message passes below: Tt[a][s'][s] (the threads of a pass run over s) and Ot[a][o][s]. Every entry starts
// at 1.0f and is multiplied, feature after feature, by that node's expected multinomial (float sum, float
// divide). Only the actions that occur in the history are flattened (used[a]).
constexpr int kFlattenCells = 2 * kMaxJournalCells; // conditional expectations one (action, state) pair may hold

__global__ void __launch_bounds__(kThreads)
    k_flatten_model(DevModel M, const float* __restrict__ counts, long long stride, const int* __restrict__ sid,
                    const unsigned char* __restrict__ used, float* __restrict__ Tt_all, float* __restrict__ Ot_all)
{
    long long const p = blockIdx.y;
    const float* c    = counts + p * stride;
    float* Tt         = Tt_all + p * (long long)M.A * M.S * M.S;
    float* Ot         = Ot_all + p * (long long)M.A * M.O * M.S;
    bool const single_o = (M.FO == 1);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < (long long)M.A * M.S;
         k += (long long)gridDim.x * blockDim.x)
    {
        int const a = (int)(k / M.S), s = (int)(k - (long long)a * M.S);
        if (!used[a]) continue;
        const Node* nodes = M.nodes + ((long long)sid[p] * M.A + a) * M.J;
        Feat const x      = decode(s, M.step_s, M.FS, M.pow2_s, M.shift_s);
        // DBNNode::expectation of every node for this (action, parent state), once (BABNModel.cpp:105-111,152-158):
        // the S (or O) products below only pick entries of these rows
        float cond[kFlattenCells];
        int first[2 * FBA_MAX_FEATURES];
        int n_cells = 0;
        for (int f = 0; f < M.J; ++f)
        {
            Node const nd   = nodes[f];
            int const range = (f < M.FS) ? M.feat_s[f] : M.feat_o[f - M.FS];
            const float* row = c + nd.off + parent_config(M, nd.par, x) * range;
            first[f]         = n_cells;
            for (int v = 0; v < range; ++v) cond[n_cells + v] = expected_mult_at(row, range, v);
            n_cells += range;
        }
        // T[s][a][s'] = ((1 * c_0[x'_0]) * c_1[x'_1]) * ... in feature order. Successor states are walked in index
        // order (last feature fastest), so consecutive ones share the product over their common leading features:
        // prefix[k] = the product over features < k, redone only from the first feature that changed (about two
        // multiplies per state instead of FS, and no index decoding) — the same multiplications in the same order.
        float prefix[FBA_MAX_FEATURES + 1];
        int dig[FBA_MAX_FEATURES];
        prefix[0] = 1.0f;
        for (int f = 0; f < M.FS; ++f)
        {
            dig[f]        = 0;
            prefix[f + 1] = __fmul_rn(prefix[f], cond[first[f]]);
        }
        for (int s2 = 0; s2 < M.S; ++s2)
        {
            Tt[((long long)a * M.S + s2) * M.S + s] = prefix[M.FS];
            int f = M.FS - 1;
            while (f >= 0 && ++dig[f] == M.feat_s[f]) dig[f--] = 0;
            if (f < 0) break;
            for (int k = f; k < M.FS; ++k) prefix[k + 1] = __fmul_rn(prefix[k], cond[first[k] + dig[k]]);
        }
        for (int o = 0; o < M.O; ++o)
        { // the observation's parents are the features of the state it is made in
            Feat const of = decode(o, M.step_o, M.FO, M.pow2_o, M.shift_o);
            float pr      = 1.0f;
            for (int q = 0; q < M.FO; ++q) pr = __fmul_rn(pr, cond[first[M.FS + q] + of.get(q, single_o)]);
            Ot[((long long)a * M.O + o) * M.S + s] = pr;
        }
    }
}

// rnd::sample::Dir::sampleFromMult<double> (random.hpp:93-115)
template<class R>
__device__ __forceinline__ int sample_from_mult_d(const double* mult, int n, double total, R& g)
{
    double const p = __dmul_rn(draw_u(g), total);
    double sum     = mult[0];
    for (int i = 1; i < n; ++i)
    {
        if (p < sum) return i - 1;
        sum = __dadd_rn(sum, mult[i]);
    }
    return n - 1;
}

// msgSampleStateHistory (:96-213), one CTA per particle. Per episode: the backward pass — message[t][s] =
// sum_s' T[s][a_t][s'] message[t+1][s'] (a sequential double sum per state, as std::inner_product runs it; one
// thread per state), times the observation factor (or the state prior at t = 0), normalised by the SEQUENTIAL
// sum over states (thread 0) — then the forward pass: s_0 from message[0], s_t+1 from T[s_t][a_t][.] *
// message[t+1][.] with its sequential total (thread 0 draws). msg: per particle (max_len + 2) x S doubles of
// scratch (the last row holds the forward pass's products).
constexpr int kMsgThreads = 1024; // one state per thread up to S = 1024: the passes are bound by loads in flight

// (Measured and dropped: widening the float table entries to double on the integer pipe instead of F2F.F64.F32 —
// ncu showed the stalls on the converter were the wait for the loads it is the first consumer of, and the extra
// integer instructions made the pass slower, 9.6 -> 13.4 ms; profiles/r2n_*.)
// SH: the message row being read and the row being built live in shared memory (2 S doubles) — the sequential
// parts (thread 0 adds S values in order, then walks them to draw) then run at shared-memory latency instead of
// L2 latency while the other threads wait; models with more than 3072 states fall back to global rows.
template<bool REPLAY, bool SH>
__global__ void __launch_bounds__(kMsgThreads)
    k_state_history_msg(DevModel M, long long N, HistoryArgs H, const float* __restrict__ Tt_all,
                        const float* __restrict__ Ot_all, const float* __restrict__ state_prior,
                        double* __restrict__ msg_all, RngArgs ra, int* __restrict__ states_out, long long out_stride,
                        int* __restrict__ overrun, long long* __restrict__ words_used)
{
    extern __shared__ double sh_rows[];
    long long const p = blockIdx.x;
    if (p >= N) return;
    int const S     = M.S;
    const float* Tt = Tt_all + p * (long long)M.A * S * S;
    const float* Ot = Ot_all + p * (long long)M.A * M.O * S;
    double* msg     = msg_all + p * (long long)(H.max_len + 2) * S;
    double* work    = SH ? sh_rows + S : msg + (long long)(H.max_len + 1) * S; // the row being built / the forward products
    int* out        = states_out + p * out_stride;
    auto g          = RngOf<REPLAY>::make(ra, p);
    __shared__ double sh_tot;
    __shared__ int sh_state;
    int first = 0, pos = 0;
    for (int e = 0; e < H.n_episodes; ++e)
    {
        int const len  = H.episode_len[e];
        const int* act = H.actions + first;
        const int* obs = H.observations + first;
        for (int s = threadIdx.x; s < S; s += blockDim.x)
        {
            double const v              = (double)Ot[((long long)act[len - 1] * M.O + obs[len - 1]) * S + s];
            msg[(long long)len * S + s] = v;
            if (SH) sh_rows[s] = v;
        }
        __syncthreads();
        for (int step = len - 1; step >= 0; --step)
        {
            const float* Ta    = Tt + (long long)act[step] * S * S;
            const double* next = SH ? sh_rows : msg + (long long)(step + 1) * S;
            for (int s = threadIdx.x; s < S; s += blockDim.x)
            {
                // the sum itself is sequential (std::inner_product's order); the loads are not: eight rows of T in
                // flight per thread before the adds that need them
                // (software-pipelined: the next eight are requested before the current eight are added, 16 in flight)
                double acc = 0.0;
                int s2     = 0;
                float t[8], t_next[8];
                bool have = S >= 8;
                if (have)
                {
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = __ldcs(Ta + (long long)u * S + s);
                }
                while (have)
                {
                    bool const more = s2 + 16 <= S;
                    if (more)
                    {
#pragma unroll
                        for (int u = 0; u < 8; ++u) t_next[u] = __ldcs(Ta + (long long)(s2 + 8 + u) * S + s);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, __dmul_rn((double)t[u], next[s2 + u]));
                    s2 += 8;
                    have = more;
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = t_next[u];
                }
                for (; s2 < S; ++s2) acc = __dadd_rn(acc, __dmul_rn((double)Ta[(long long)s2 * S + s], next[s2]));
                float const factor = (step != 0) ? Ot[((long long)act[step - 1] * M.O + obs[step - 1]) * S + s]
                                                 : state_prior[s];
                work[s] = __dmul_rn(acc, (double)factor);
            }
            __syncthreads();
            if (threadIdx.x == 0)
            {
                double tot = 0.0;
                for (int s = 0; s < S; ++s) tot = __dadd_rn(tot, work[s]);
                sh_tot = tot;
            }
            __syncthreads();
            double const tot = sh_tot;
            for (int s = threadIdx.x; s < S; s += blockDim.x)
            {
                double const v               = __ddiv_rn(work[s], tot);
                msg[(long long)step * S + s] = v;
                if (SH) sh_rows[s] = v;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0)
        {
            sh_state   = sample_from_mult_d(SH ? sh_rows : msg, S, 1.0, g);
            out[pos++] = sh_state;
        }
        __syncthreads();
        for (int step = 0; step < len; ++step)
        {
            int const state    = sh_state;
            const float* Ta    = Tt + (long long)act[step] * S * S;
            const double* next = msg + (long long)(step + 1) * S;
            for (int s2 = threadIdx.x; s2 < S; s2 += blockDim.x)
                work[s2] = __dmul_rn((double)Ta[(long long)s2 * S + state], next[s2]);
            __syncthreads();
            if (threadIdx.x == 0)
            {
                double tot = 0.0;
                for (int s2 = 0; s2 < S; ++s2) tot = __dadd_rn(tot, work[s2]);
                sh_state   = sample_from_mult_d(work, S, tot, g);
                out[pos++] = sh_state;
            }
            __syncthreads();
        }
        if (threadIdx.x != 0) pos += len + 1;
        first += len;
    }
    if (threadIdx.x == 0 && g.overrun) *overrun = 1;
    if (threadIdx.x == 0 && words_used) words_used[p] = rng_words_used(g, p * ra.words_per_item);
}

// The same passes with one model spread over a thread-block CLUSTER of kMsgCluster CTAs (sm_90+ distributed
// shared memory): CTA r owns the states [r S / kMsgCluster, (r + 1) S / kMsgCluster) — their inner products are
// where the time goes, and they are bound by loads in flight, so four SMs' worth of them per model is the gain —
// while every CTA keeps a full copy of the current message row in its own shared memory. A backward step: each
// thread computes its state's new entry and stores it into the work row of ALL CTAs of the cluster
// (cluster.map_shared_rank), cluster.sync, every CTA's thread 0 adds the S entries in order (the same sum in every
// CTA, no broadcast needed), every CTA normalises its full copy, cluster.sync. The forward pass is short and runs in
// CTA 0 alone; a cluster.sync ends the episode. Arithmetic and order are those of the one-CTA kernel: bit-identical.
constexpr int kMsgCluster = 4;

template<bool REPLAY>
__global__ void __cluster_dims__(kMsgCluster, 1, 1) __launch_bounds__(kMsgThreads)
    k_state_history_msg_cluster(DevModel M, long long N, HistoryArgs H, const float* __restrict__ Tt_all,
                                const float* __restrict__ Ot_all, const float* __restrict__ state_prior,
                                double* __restrict__ msg_all, RngArgs ra, int* __restrict__ states_out,
                                long long out_stride, int* __restrict__ overrun, long long* __restrict__ words_used)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double sh_rows[];
    unsigned const rank = cluster.block_rank();
    long long const p   = blockIdx.x / kMsgCluster;
    int const S         = M.S;
    int const slice     = S / kMsgCluster, s_lo = (int)rank * slice;
    double* row         = sh_rows;     // the message row being read (normalised), all S entries
    double* work        = sh_rows + S; // the row being built (all S entries, filled by every CTA of the cluster)
    double* work_of[kMsgCluster];
#pragma unroll
    for (int r = 0; r < kMsgCluster; ++r) work_of[r] = cluster.map_shared_rank(work, r);
    const float* Tt = Tt_all + p * (long long)M.A * S * S;
    const float* Ot = Ot_all + p * (long long)M.A * M.O * S;
    double* msg     = msg_all + p * (long long)(H.max_len + 2) * S;
    int* out        = states_out + p * out_stride;
    auto g          = RngOf<REPLAY>::make(ra, p);
    __shared__ double sh_tot;
    __shared__ int sh_state;
    int first = 0, pos = 0;
    for (int e = 0; e < H.n_episodes; ++e)
    {
        int const len  = H.episode_len[e];
        const int* act = H.actions + first;
        const int* obs = H.observations + first;
        for (int s = threadIdx.x; s < S; s += blockDim.x)
        {
            double const v = (double)Ot[((long long)act[len - 1] * M.O + obs[len - 1]) * S + s];
            row[s]         = v;
            if (s >= s_lo && s < s_lo + slice) msg[(long long)len * S + s] = v;
        }
        __syncthreads();
        for (int step = len - 1; step >= 0; --step)
        {
            const float* Ta = Tt + (long long)act[step] * S * S;
            for (int s = s_lo + (int)threadIdx.x; s < s_lo + slice; s += blockDim.x)
            {
                double acc = 0.0;
                int s2     = 0;
                float t[8], t_next[8];
                bool have = S >= 8;
                if (have)
                {
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = __ldcs(Ta + (long long)u * S + s);
                }
                while (have)
                {
                    bool const more = s2 + 16 <= S;
                    if (more)
                    {
#pragma unroll
                        for (int u = 0; u < 8; ++u) t_next[u] = __ldcs(Ta + (long long)(s2 + 8 + u) * S + s);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, __dmul_rn((double)t[u], row[s2 + u]));
                    s2 += 8;
                    have = more;
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = t_next[u];
                }
                for (; s2 < S; ++s2) acc = __dadd_rn(acc, __dmul_rn((double)Ta[(long long)s2 * S + s], row[s2]));
                float const factor = (step != 0) ? Ot[((long long)act[step - 1] * M.O + obs[step - 1]) * S + s]
                                                 : state_prior[s];
                double const v = __dmul_rn(acc, (double)factor);
#pragma unroll
                for (int r = 0; r < kMsgCluster; ++r) work_of[r][s] = v;
            }
            cluster.sync();
            if (threadIdx.x == 0)
            {
                double tot = 0.0;
                for (int s = 0; s < S; ++s) tot = __dadd_rn(tot, work[s]);
                sh_tot = tot;
            }
            __syncthreads();
            double const tot = sh_tot;
            for (int s = threadIdx.x; s < S; s += blockDim.x)
            {
                double const v = __ddiv_rn(work[s], tot);
                row[s]         = v;
                if (s >= s_lo && s < s_lo + slice) msg[(long long)step * S + s] = v;
            }
            cluster.sync(); // nobody writes the next step's entries into a work row that is still being read
        }
        if (rank == 0)
        { // forward sampling (thread 0 draws; the products of a step by all threads of this CTA)
            __threadfence();
            if (threadIdx.x == 0)
            {
                sh_state   = sample_from_mult_d(row, S, 1.0, g);
                out[pos++] = sh_state;
            }
            __syncthreads();
            for (int step = 0; step < len; ++step)
            {
                int const state    = sh_state;
                const float* Ta    = Tt + (long long)act[step] * S * S;
                const double* next = msg + (long long)(step + 1) * S;
                for (int s2 = threadIdx.x; s2 < S; s2 += blockDim.x)
                    work[s2] = __dmul_rn((double)Ta[(long long)s2 * S + state], __ldcg(next + s2));
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    double tot = 0.0;
                    for (int s2 = 0; s2 < S; ++s2) tot = __dadd_rn(tot, work[s2]);
                    sh_state   = sample_from_mult_d(work, S, tot, g);
                    out[pos++] = sh_state;
                }
                __syncthreads();
            }
        }
        cluster.sync(); // the other CTAs start the next episode (and touch CTA 0's work row) only now
        first += len;
    }
    if (rank == 0 && threadIdx.x == 0 && g.overrun) *overrun = 1;
    if (rank == 0 && threadIdx.x == 0 && words_used) words_used[p] = rng_words_used(g, p * ra.words_per_item);
}

// MHwithinGibbs::computePosteriorCounts (:397-436): incrementCountsOf(s_t, a_t, o_t, s_t+1) for every step of
// every particle's state history, one thread per (particle, step). Every increment is exactly +1.0f, so the
// atomic adds give the reference's sequence of float sums whatever order they land in.
__global__ void __launch_bounds__(kThreads)
    k_add_history_counts(DevModel M, float* counts, long long stride, const int* __restrict__ sid, long long N,
                         int n_steps, const int* __restrict__ step_action, const int* __restrict__ step_obs,
                         const int* __restrict__ step_pos, const int* __restrict__ states, long long states_stride)
{
    long long const k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N * n_steps) return;
    long long const i = k / n_steps;
    int const t       = (int)(k - i * n_steps);
    const int* seq    = states + i * states_stride; // states_stride == 0: one history shared by all particles
    int const s = seq[step_pos[t]], s2 = seq[step_pos[t] + 1], a = step_action[t], o = step_obs[t];
    const Node* nodes = M.nodes + ((long long)sid[i] * M.A + a) * M.J;
    float* c          = counts + i * stride;
    bool const single_s = (M.FS == 1), single_o = (M.FO == 1);
    Feat const x  = decode(s, M.step_s, M.FS, M.pow2_s, M.shift_s);
    Feat const x2 = decode(s2, M.step_s, M.FS, M.pow2_s, M.shift_s);
    Feat const of = decode(o, M.step_o, M.FO, M.pow2_o, M.shift_o);
    for (int f = 0; f < M.FS; ++f)
    {
        Node const nd = nodes[f];
        atomicAdd(c + nd.off + parent_config(M, nd.par, x) * M.feat_s[f] + x2.get(f, single_s), 1.0f);
    }
    Feat const& xo = M.tabular ? x2 : x; // BABNModel.cpp:366,380: the observation CPTs are indexed by the OLD state
    for (int q = 0; q < M.FO; ++q)
    {
        Node const nd = nodes[M.FS + q];
        atomicAdd(c + nd.off + parent_config(M, nd.par, xo) * M.feat_o[q] + of.get(q, single_o), 1.0f);
    }
}

} // namespace fba
