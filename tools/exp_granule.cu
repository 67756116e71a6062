// Scratch micro-benchmark (not part of the product): what does the HBM deliver for k_propose's ACCESS
// PATTERN when nothing else is in the way? One thread per particle, J = 11 independent 8-byte reads at
// pseudo-random rows inside one 592-byte window of an 11 840-byte block (particle stride), followed by
// 11 single-cell increments written back in different ways — no random numbers, no arithmetic, all
// loads of a thread in flight together. Prints the best time of 10 launches per variant.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_granule tools/exp_granule.cu && ./exp_granule
#include <cstdio>
#include <cuda_runtime.h>

enum Variant
{
    READ_ONLY = 0,
    ST4,        // 4-byte store of the incremented cell (what k_propose does)
    ST8,        // 8-byte store of the whole row
    ST16,       // 16-byte aligned store (row + its neighbour, both read before)
    ST32,       // the whole 32-byte sector, two 16-byte stores
    ST4_CS,     // 4-byte store, streaming hint
    ST4_WT,     // 4-byte store, write-through hint
    RED4,       // red.global.add.f32 (fire-and-forget atomic at the L2)
    ST4_NOREAD, // 4-byte store without reading anything first
    N_VARIANTS
};
static const char* kName[] = {"read only", "st 4 B", "st 8 B (row)", "st 16 B", "st 32 B (sector)", "st 4 B .cs",
                              "st 4 B .wt", "red.add.f32", "st 4 B, no read"};

template<int V>
__global__ void __launch_bounds__(256) k_touch(float* buf, long long stride, long long n, int window_off, float* out)
{
    long long const i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* c   = buf + i * stride + window_off;
    unsigned h = (unsigned)i * 2654435761u;
    float4 v[11][2];
    int cell[11];
#pragma unroll
    for (int k = 0; k < 11; ++k)
    { // node k owns 16 floats (8 rows of 2) = two 32-byte sectors; the row is "chosen by the state"
        h       = h * 1664525u + 1013904223u;
        cell[k] = (k * 12 + 2 * ((h >> 24) & 7)) & ~1; // 12-float pitch keeps 8-byte alignment, 11 nodes < 148 floats
        if (V == ST4_NOREAD) continue;
        if (V == ST16 || V == ST32)
        {
            int const s0 = cell[k] & ~7; // the 32-byte sector of the row (window offsets are multiples of 4 floats)
            v[k][0]      = *reinterpret_cast<const float4*>(c + s0);
            v[k][1]      = *reinterpret_cast<const float4*>(c + s0 + 4);
        } else
        {
            float2 const r = *reinterpret_cast<const float2*>(c + cell[k]);
            v[k][0]        = make_float4(r.x, r.y, 0.f, 0.f);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k)
        if (V != ST4_NOREAD) acc += v[k][0].x + v[k][0].y;
#pragma unroll
    for (int k = 0; k < 11; ++k)
    {
        float* p = c + cell[k];
        if (V == ST4) *p = v[k][0].x + 1.0f;
        if (V == ST4_NOREAD) *p = 1.0f;
        if (V == ST8) *reinterpret_cast<float2*>(p) = make_float2(v[k][0].x + 1.0f, v[k][0].y);
        if (V == ST16 || V == ST32)
        {
            int const s0 = cell[k] & ~7, in = cell[k] - s0; // in = 0, 2, 4, 6
            float4 a = v[k][0], b = v[k][1];
            if (in == 0) a.x += 1.0f;
            if (in == 2) a.z += 1.0f;
            if (in == 4) b.x += 1.0f;
            if (in == 6) b.z += 1.0f;
            if (V == ST32 || in < 4) *reinterpret_cast<float4*>(c + s0) = a;
            if (V == ST32 || in >= 4) *reinterpret_cast<float4*>(c + s0 + 4) = b;
        }
        if (V == ST4_CS) __stcs(p, v[k][0].x + 1.0f);
        if (V == ST4_WT) __stwt(p, v[k][0].x + 1.0f);
        if (V == RED4) atomicAdd(p, 1.0f);
    }
    out[i] = acc;
}

template<int V>
static void run(float* buf, float* out, long long n, long long stride)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 12; ++rep)
    {
        int const window = 148 * (rep % 20); // another action's window every launch (multiple of 4 floats)
        cudaEventRecord(e0);
        k_touch<V><<<(unsigned)((n + 255) / 256), 256>>>(buf, stride, n, window, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    printf("%-18s %.3f ms\n", kName[V], best);
}

int main()
{
    long long const n = 1250000, stride = 2960;
    float *buf, *out;
    cudaMalloc(&buf, n * stride * sizeof(float));
    cudaMalloc(&out, n * sizeof(float));
    cudaMemset(buf, 0, n * stride * sizeof(float));
    run<READ_ONLY>(buf, out, n, stride);
    run<ST4>(buf, out, n, stride);
    run<ST8>(buf, out, n, stride);
    run<ST16>(buf, out, n, stride);
    run<ST32>(buf, out, n, stride);
    run<ST4_CS>(buf, out, n, stride);
    run<ST4_WT>(buf, out, n, stride);
    run<RED4>(buf, out, n, stride);
    run<ST4_NOREAD>(buf, out, n, stride);
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
