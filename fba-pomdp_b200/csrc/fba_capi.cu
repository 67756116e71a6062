// Host side of the C ABI declared in include/fba_pomdp_b200.h: owns device memory, feeds the
// replay stream, launches the kernels in fba_kernels.cuh. No CPU implementation of any operation
// lives here: without a CUDA device fba_ctx_create fails with FBA_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

#include "fba_kernels.cuh"

using namespace fba;

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct fba_ctx
{
    int device           = 0;
    cudaStream_t stream  = nullptr;
    int sm_count         = 148;
    int64_t launches     = 0;
    std::string err;
    // optional per-kernel CUDA-event timing (fba_ctx_profile_*)
    int bulk_copy    = 0;  // 1: full-copy gathers go through the TMA engine (k_gather_bulk)
    int rollout_coop = -1; // -1 auto (by batch size and row length), 0 thread per rollout, 1 warp per rollout
    void* big_scratch = nullptr; // grow-only: the flattened models + messages of fba_belief_sample_state_history
    size_t big_scratch_bytes = 0;
    bool msg_cluster  = true;  // Gibbs message passing: one model over a cluster of CTAs when the model allows it
    bool auto_compact = true;  // a full journal / delta list turns the belief into dense storage instead of failing
    bool nested_exact = false; // PHILOX NestedBelief updates: thread per top particle instead of warp per top particle
    bool inplace_resample = true; // PHILOX mode: survivors keep their slot (fba_ctx_set_option)
    bool fused_update     = true; // small beliefs: update + resample in ONE launch (k_runs_step, one CTA)
    long long parallel_chains_min = 8192; // REPLAY: beliefs at least this large evaluate the reference's
                                          // sequential weight chains in parallel (bit-identical)
    bool profiling       = false;
    struct Timed
    {
        const char* name;
        cudaEvent_t start, stop;
    };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> event_pool; // created at profile_begin so launches stay cheap
    std::map<std::string, std::pair<double, int64_t>> kernel_ms;
    // scratch shared by the beliefs of this context
    uint32_t* d_words    = nullptr;
    size_t words_cap     = 0;
    long long* d_offsets = nullptr;
    size_t offsets_cap   = 0;
    int* d_flag          = nullptr; // overrun flag
    unsigned long long* d_counters = nullptr; // [0] simulated steps executed by rollout kernels
    int* h_flag          = nullptr; // pinned
    double* h_scal       = nullptr; // pinned, 4 doubles
};

struct fba_model
{
    fba_ctx* ctx = nullptr;
    DevModel dev{};
    int max_structs = 0, n_structs = 0;
    bool long_rows = false; // some feature has more than 4 values: kernels load rows in chunks
    int delta_cap  = 0;     // > 0: tabular base+delta storage with this many increments per particle
    std::vector<uint32_t> t_par, o_par; // host structure table
    std::vector<int> sizes;
    std::vector<Node> h_nodes;
    std::unordered_map<std::string, int> index; // parent masks (bytes) -> structure id
    std::string key_buf;                        // reused by add_structures
    Node* d_nodes = nullptr;
    int* d_sizes  = nullptr;
    std::vector<void*> owned; // device copies of the descriptor's tables
};

struct fba_belief
{
    fba_ctx* ctx   = nullptr;
    fba_model* m   = nullptr;
    long long N    = 0, stride = 0; // stride: PHYSICAL floats per particle block
    long long lstride = 0;          // logical (dense) cells per particle, what the API calls "stride"
    bool weighted  = true;
    // base+delta storage: the dense tables of the prior prototypes, shared by all particles
    int delta_cap     = 0;
    float* base       = nullptr;
    int n_bases       = 0;
    std::vector<float> h_base;
    long long delta_used = 0;        // increments every particle's list holds (equal for all: one update appends J)
    int* d_proto_sid  = nullptr;     // prototype (base table) -> structure id
    std::vector<int> h_proto_sid;
    float* counts[2] = {nullptr, nullptr};
    int* state[2]    = {nullptr, nullptr};
    int* sid[2]      = {nullptr, nullptr};
    int cur          = 0;
    double* w        = nullptr;
    double* aux      = nullptr; // R (replay) or cdf (native), N doubles
    double* tile     = nullptr; // tile sums
    double* scal     = nullptr; // device scalars: [0] total, [1] total_weight
    int* anc         = nullptr; // ancestors
    long long anc_cap = 0;
    double total_weight = 1.0;  // WeightedFilter::_total_weight (host mirror)
    double uniform_total = -1;  // sequential sum of N x (1/N), computed on first use
    bool suffix_valid = false;  // aux holds R for the current weights
    bool cdf_valid    = false;  // aux holds the native cdf for the current weights
    bool uniform_now  = false;  // every weight is 1 / N: set where a resample / init leaves them so, cleared wherever the
                                // cdf cache is invalidated (i.e. wherever weights may have changed)
    bool w_escaped    = false;  // fba_belief_weight_ptr handed the weight array out: never assume anything about it
    // in-place resampling
    int *noff = nullptr, *escan = nullptr, *dead = nullptr, *totals = nullptr, *src_of = nullptr;
    long long* stats = nullptr; // [0] copies made by in-place resamples, [1] number of resamples,
                                // [2] surplus records dropped because the export buffer was full
    long long src_cap = 0;
    // peer-to-peer exchange (fba_belief_p2p_*): control block [P2PCtrl header | dead-slot list], mapped by peers
    PeerTable peers{};
    bool peers_open = false;
    int p2p_ranks = 0, p2p_rank = 0;
    long long* d_plan = nullptr;
    char* p2p_block   = nullptr;         // owns b->totals and b->dead once allocated
    unsigned long long* d_step = nullptr; // step stamp of the sharded updates
    std::vector<void*> opened; // cudaIpcOpenMemHandle'd base pointers
    long long* d_quota = nullptr; // device-side offspring quota (fba_belief_shard_resample_async)
    int2* tile_pairs = nullptr;
    bool inplace_last = false; // the last shard resample ran in place (import goes to dead slots)
    // pinned host mirrors of state / structure ids for the sequential part of reinvigoration
    int *h_sid = nullptr, *h_state = nullptr;
    BreedJob* d_jobs   = nullptr; // grow-only
    long long jobs_cap = 0;
    // in-place rejection sampling: the accepted attempts (N each) and scan scratch, allocated on first use
    int *rs_src = nullptr, *rs_state = nullptr, *rs_rec = nullptr, *rs_flag = nullptr, *rs_pos = nullptr,
        *rs_tiles = nullptr;
    // rejection sampling wave buffers
    long long wave_cap = 0;
    int *att_src = nullptr, *att_state = nullptr, *att_accept = nullptr, *att_pos = nullptr,
        *att_rec = nullptr, *d_total = nullptr, *att_tiles = nullptr;
    // rollout request / result staging (grow-only)
    long long roll_cap = 0, roll_cap_p = 0, step_cap = 0, step_cap_r = 0;
    int* step_i    = nullptr;
    double* step_r = nullptr;
    long long* roll_p  = nullptr;
    int *roll_s = nullptr, *roll_d = nullptr;
    double* roll_r = nullptr;
    // REPLAY, parallel evaluation of the sequential weight chains: per-segment scratch, grow-once
    double *seg_start = nullptr, *seg_end = nullptr, *seg_delta = nullptr;
    unsigned char* seg_tie = nullptr;
    long long* chain_stats = nullptr; // [0] segments recomputed sequentially, [1] segments walked
    // multi-GPU staging
    char* xport = nullptr;
    long long xport_cap = 0, xport_count = 0;
    char* import_buf = nullptr;
    long long import_cap = 0;
    long long local_kept = 0;
    long long imported   = 0; // records imported since the last shard resample
};

#define CU(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
        {                                                                                          \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            return FBA_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define REQUIRE(ctx, cond, msg)                                                                    \
    do {                                                                                           \
        if (!(cond))                                                                               \
        {                                                                                          \
            (ctx)->err = (msg);                                                                    \
            return FBA_ERR_INVALID;                                                                \
        }                                                                                          \
    } while (0)

static inline int blocks_for(long long n, int per_block = kThreads)
{
    return (int)std::max<long long>(1, (n + per_block - 1) / per_block);
}

// grid for warp-per-particle streaming kernels: enough CTAs to fill every SM several times over,
// capped so each warp loops (grid = multiple of the SM count)
static inline int stream_grid(const fba_ctx* ctx, long long n_particles)
{
    long long const need = (n_particles + (kThreads / 32) - 1) / (kThreads / 32);
    long long const cap  = (long long)ctx->sm_count * 8;
    return (int)std::max<long long>(1, std::min(need, cap));
}

static void profile_mark(fba_ctx* ctx, const char* name, bool begin)
{
    if (*name == '(') ++name; // template kernels with a comma are passed parenthesised
    if (begin)
    {
        fba_ctx::Timed t{name, nullptr, nullptr};
        for (cudaEvent_t* e : {&t.start, &t.stop})
        {
            if (ctx->event_pool.empty()) cudaEventCreate(e);
            else
            {
                *e = ctx->event_pool.back();
                ctx->event_pool.pop_back();
            }
        }
        cudaEventRecord(t.start, ctx->stream);
        ctx->timed.push_back(t);
    } else
        cudaEventRecord(ctx->timed.back().stop, ctx->stream);
}

#define LAUNCH(ctx, kernel, grid, block, ...)                                                      \
    do {                                                                                           \
        if ((ctx)->profiling) profile_mark(ctx, #kernel, true);                                    \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);                                \
        if ((ctx)->profiling) profile_mark(ctx, #kernel, false);                                   \
        ++(ctx)->launches;                                                                         \
        CU(ctx, cudaGetLastError());                                                               \
    } while (0)

// kernel<REPLAY, LONG, SAMPLED> chosen at run time (sampled-Dirichlet models never replay)
#define LAUNCH_RL(ctx, kernel, replay, longrows, grid, block, ...)                                 \
    do {                                                                                           \
        bool const smp_ = D.sampled != 0;                                                          \
        if (replay)                                                                                \
        {                                                                                          \
            if (longrows) LAUNCH(ctx, (kernel<true, true, false>), grid, block, __VA_ARGS__);      \
            else                                                                                   \
                LAUNCH(ctx, (kernel<true, false, false>), grid, block, __VA_ARGS__);               \
        } else if (smp_)                                                                           \
        {                                                                                          \
            if (longrows) LAUNCH(ctx, (kernel<false, true, true>), grid, block, __VA_ARGS__);      \
            else                                                                                   \
                LAUNCH(ctx, (kernel<false, false, true>), grid, block, __VA_ARGS__);               \
        } else                                                                                     \
        {                                                                                          \
            if (longrows) LAUNCH(ctx, (kernel<false, true, false>), grid, block, __VA_ARGS__);     \
            else                                                                                   \
                LAUNCH(ctx, (kernel<false, false, false>), grid, block, __VA_ARGS__);              \
        }                                                                                          \
    } while (0)

// delta kernel<REPLAY, SAMPLED>
#define LAUNCH_DELTA(ctx, kernel, replay, grid, block, ...)                                        \
    do {                                                                                           \
        if (replay) LAUNCH(ctx, (kernel<true, false>), grid, block, __VA_ARGS__);                  \
        else if (D.sampled)                                                                        \
            LAUNCH(ctx, (kernel<false, true>), grid, block, __VA_ARGS__);                          \
        else                                                                                       \
            LAUNCH(ctx, (kernel<false, false>), grid, block, __VA_ARGS__);                         \
    } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int fba_abi_version(void)
{
    return FBA_ABI_VERSION;
}

extern "C" void fba_ctx_destroy(fba_ctx* ctx);

extern "C" int fba_ctx_create(int device, fba_ctx** out)
{
    if (!out) return FBA_ERR_INVALID;
    *out      = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return FBA_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return FBA_ERR_INVALID;
    auto ctx    = new fba_ctx();
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (const char* g = getenv("FBA_B200_L2_FETCH")) // experiment knob: 32 / 64 / 128 bytes
        if (e == cudaSuccess) e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess)
    { // scratch of the short MCMC calls comes from the stream-ordered pool (PoolTmp): keep up to 256 MB cached
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool)
        {
            unsigned long long threshold = 256ull << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
        cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_counters, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_counters, 0, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_scal, 4 * sizeof(double));
    if (e != cudaSuccess)
    {
        fba_ctx_destroy(ctx);
        return FBA_ERR_CUDA;
    }
    *out = ctx;
    return FBA_OK;
}

extern "C" void fba_ctx_destroy(fba_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    cudaFree(ctx->d_words);
    cudaFree(ctx->d_offsets);
    cudaFree(ctx->big_scratch);
    cudaFree(ctx->d_flag);
    cudaFree(ctx->d_counters);
    cudaFreeHost(ctx->h_flag);
    cudaFreeHost(ctx->h_scal);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* fba_last_error(const fba_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : "no context";
}
extern "C" void* fba_ctx_stream(const fba_ctx* ctx)
{
    return ctx ? (void*)ctx->stream : nullptr;
}
extern "C" int fba_ctx_synchronize(fba_ctx* ctx)
{
    if (!ctx) return FBA_ERR_INVALID;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}
// which = 0: simulated steps executed by the rollout kernels since the context was created
// (rollouts end early at terminal states; the roofline of k_rollouts is quoted on this count)
extern "C" int64_t fba_ctx_counter(fba_ctx* ctx, int32_t which)
{
    if (!ctx || which < 0 || which >= 4) return -1;
    unsigned long long h = 0;
    cudaSetDevice(ctx->device);
    if (cudaMemcpyAsync(&h, ctx->d_counters + which, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    return (int64_t)h;
}

extern "C" int64_t fba_ctx_launch_count(const fba_ctx* ctx)
{
    return ctx ? ctx->launches : -1;
}

extern "C" int fba_ctx_set_option(fba_ctx* ctx, const char* name, int64_t value)
{
    if (!ctx || !name) return FBA_ERR_INVALID;
    if (!strcmp(name, "inplace_resample"))
    {
        ctx->inplace_resample = value != 0;
        return FBA_OK;
    }
    if (!strcmp(name, "fused_update"))
    {
        ctx->fused_update = value != 0;
        return FBA_OK;
    }
    if (!strcmp(name, "bulk_copy"))
    {
        ctx->bulk_copy = value != 0;
        return FBA_OK;
    }
    if (!strcmp(name, "parallel_chains_min"))
    { // 0: always the parallel evaluation of the REPLAY weight chains; a huge value: never
        ctx->parallel_chains_min = value;
        return FBA_OK;
    }
    if (!strcmp(name, "rollout_coop"))
    {
        ctx->rollout_coop = (int)value;
        return FBA_OK;
    }
    if (!strcmp(name, "msg_cluster"))
    { // 0: fba_belief_sample_state_history (messages) always runs one CTA per model
        ctx->msg_cluster = value != 0;
        return FBA_OK;
    }
    if (!strcmp(name, "auto_compact"))
    { // 0: a full increment list is an error (FBA_ERR_CAPACITY), as before
        ctx->auto_compact = value != 0;
        return FBA_OK;
    }
    if (!strcmp(name, "nested_exact"))
    { // 1: PHILOX-mode NestedBelief updates run the reference's loop attempt by attempt (one thread per top particle)
        ctx->nested_exact = value != 0;
        return FBA_OK;
    }
    ctx->err = std::string("unknown option ") + name;
    return FBA_ERR_INVALID;
}

// Per-kernel timing with CUDA events on the context's stream (bench.py's roofline numbers).
extern "C" int fba_ctx_profile_begin(fba_ctx* ctx)
{
    if (!ctx) return FBA_ERR_INVALID;
    ctx->kernel_ms.clear();
    while (ctx->event_pool.size() < 2048)
    {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) break;
        ctx->event_pool.push_back(e);
    }
    ctx->profiling = true;
    return FBA_OK;
}

extern "C" int fba_ctx_profile_end(fba_ctx* ctx)
{
    if (!ctx) return FBA_ERR_INVALID;
    ctx->profiling = false;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& t : ctx->timed)
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.start, t.stop);
        auto& acc = ctx->kernel_ms[t.name];
        acc.first += ms;
        acc.second += 1;
        ctx->event_pool.push_back(t.start);
        ctx->event_pool.push_back(t.stop);
    }
    ctx->timed.clear();
    return FBA_OK;
}

// total milliseconds and launch count of kernels whose name starts with `prefix`
extern "C" int fba_ctx_profile_get(const fba_ctx* ctx, const char* prefix, double* total_ms, int64_t* count)
{
    if (!ctx || !prefix) return FBA_ERR_INVALID;
    double ms = 0;
    int64_t n = 0;
    for (auto const& kv : ctx->kernel_ms)
        if (kv.first.compare(0, strlen(prefix), prefix) == 0) ms += kv.second.first, n += kv.second.second;
    if (total_ms) *total_ms = ms;
    if (count) *count = n;
    return FBA_OK;
}

// "name:ms:count;..." for every kernel timed since fba_ctx_profile_begin; returns the length needed
extern "C" int64_t fba_ctx_profile_list(const fba_ctx* ctx, char* buf, int64_t cap)
{
    if (!ctx) return -1;
    std::string out;
    for (auto const& kv : ctx->kernel_ms)
        out += kv.first + ":" + std::to_string(kv.second.first) + ":" + std::to_string(kv.second.second) + ";";
    if (buf && cap > 0)
    {
        size_t const n = std::min<size_t>(out.size(), (size_t)cap - 1);
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int64_t)out.size() + 1;
}

// upload words[first, first+n) of the replay stream into the context scratch
static int stage_words(fba_ctx* ctx, const fba_rng* rng, long long n)
{
    if (rng->cursor < 0 || rng->cursor + n > rng->n_words)
    {
        ctx->err = "replay stream underrun: need " + std::to_string(n) + " words at cursor "
                   + std::to_string(rng->cursor) + " of " + std::to_string(rng->n_words);
        return FBA_ERR_RNG_UNDERRUN;
    }
    if ((size_t)n > ctx->words_cap)
    {
        cudaFree(ctx->d_words); // (synchronises: no kernel still reads the old buffer)
        ctx->d_words     = nullptr;
        ctx->words_cap   = 0;
        size_t const cap = (size_t)n + (size_t)n / 2 + 1024;
        CU(ctx, cudaMalloc(&ctx->d_words, cap * sizeof(uint32_t)));
        ctx->words_cap = cap;
    }
    if (n)
        CU(ctx, cudaMemcpyAsync(ctx->d_words, rng->words + rng->cursor, (size_t)n * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, ctx->stream));
    return FBA_OK;
}

static int stage_offsets(fba_ctx* ctx, const std::vector<long long>& off)
{
    if (off.size() > ctx->offsets_cap)
    {
        cudaFree(ctx->d_offsets);
        ctx->d_offsets   = nullptr;
        ctx->offsets_cap = 0;
        size_t const cap = off.size() + off.size() / 2 + 1024;
        CU(ctx, cudaMalloc(&ctx->d_offsets, cap * sizeof(long long)));
        ctx->offsets_cap = cap;
    }
    // pageable source: the copy is staged before the call returns, so `off` may die afterwards
    CU(ctx, cudaMemcpyAsync(ctx->d_offsets, off.data(), off.size() * sizeof(long long),
                            cudaMemcpyHostToDevice, ctx->stream));
    return FBA_OK;
}

static int clear_flag(fba_ctx* ctx)
{
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    return FBA_OK;
}

static int check_flag(fba_ctx* ctx)
{
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag == 2)
    {
        ctx->err = "base+delta particle ran out of increment slots (raise fba_model_desc.delta_capacity)";
        return FBA_ERR_CAPACITY;
    }
    if (*ctx->h_flag)
    {
        ctx->err = "replay stream underrun on device";
        return FBA_ERR_RNG_UNDERRUN;
    }
    return FBA_OK;
}

static RngArgs replay_args(fba_ctx* ctx, long long n_words, long long per_item, bool use_offsets)
{
    RngArgs ra{};
    ra.words          = ctx->d_words;
    ra.n_words        = n_words;
    ra.offsets        = use_offsets ? ctx->d_offsets : nullptr;
    ra.words_per_item = per_item;
    return ra;
}

static RngArgs philox_args(fba_rng* rng, unsigned long long stream_base = 0)
{
    RngArgs ra{};
    ra.seed        = rng->seed;
    ra.offset      = rng->offset++;
    ra.stream_base = stream_base;
    return ra;
}

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
template<class T>
static int to_device(fba_model* m, const T* host, size_t n, const T** out)
{
    *out = nullptr;
    if (!host || !n) return FBA_OK;
    T* d = nullptr;
    CU(m->ctx, cudaMalloc(&d, n * sizeof(T)));
    m->owned.push_back(d); // owned from here on: freed by fba_model_destroy even if the copy fails
    CU(m->ctx, cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = d;
    return FBA_OK;
}

extern "C" int fba_model_create(fba_ctx* ctx, const fba_model_desc* d, int32_t max_structures,
                                fba_model** out)
{
    if (!ctx || !d || !out) return FBA_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    REQUIRE(ctx, d->S > 0 && d->A > 0 && d->O > 0, "model: S, A and O must be positive");
    REQUIRE(ctx, d->n_state_features >= 1 && d->n_state_features <= FBA_MAX_FEATURES
                     && d->n_obs_features >= 1 && d->n_obs_features <= FBA_MAX_FEATURES,
            "model: feature counts must be in [1, 16]");
    REQUIRE(ctx, max_structures >= 1, "model: max_structures must be >= 1");
    long long ps = 1, po = 1;
    for (int f = 0; f < d->n_state_features; ++f)
    {
        REQUIRE(ctx, d->state_feature_sizes[f] >= 1, "model: state feature size < 1");
        REQUIRE(ctx, d->n_state_features == 1 || d->state_feature_sizes[f] <= 256,
                "model: with several state features each must have <= 256 values");
        ps *= d->state_feature_sizes[f];
    }
    for (int f = 0; f < d->n_obs_features; ++f)
    {
        REQUIRE(ctx, d->obs_feature_sizes[f] >= 1, "model: observation feature size < 1");
        REQUIRE(ctx, d->n_obs_features == 1 || d->obs_feature_sizes[f] <= 256,
                "model: with several observation features each must have <= 256 values");
        po *= d->obs_feature_sizes[f];
    }
    REQUIRE(ctx, ps == d->S && po == d->O, "model: feature sizes do not multiply to S / O");
    REQUIRE(ctx, !d->tabular || (d->n_state_features == 1 && d->n_obs_features == 1),
            "model: tabular models have exactly one state and one observation feature");
    REQUIRE(ctx, d->delta_capacity >= 0, "model: negative delta_capacity");
    if (d->delta_capacity > 0 && !d->tabular)
    { // factored journal: the rows gathered in one pass must fit kMaxJournalCells
        int sum_s = 0, sum_o = 0;
        for (int f = 0; f < d->n_state_features; ++f) sum_s += d->state_feature_sizes[f];
        for (int f = 0; f < d->n_obs_features; ++f) sum_o += d->obs_feature_sizes[f];
        REQUIRE(ctx, sum_s <= kMaxJournalCells && sum_o <= kMaxJournalCells,
                "model: base+journal storage needs the feature ranges to sum to at most 96");
        REQUIRE(ctx, d->delta_capacity % (d->n_state_features + d->n_obs_features) == 0,
                "model: a factored model's delta_capacity must be a multiple of its number of nodes per action");
    }
    REQUIRE(ctx, d->delta_capacity == 0 || !d->tabular || (d->S <= kMaxDeltaRow && d->O <= kMaxDeltaRow),
            "model: base+delta storage supports rows of at most " + std::to_string(kMaxDeltaRow) + " cells");

    auto m         = new fba_model();
    m->ctx         = ctx;
    m->max_structs = max_structures;
    m->delta_cap   = d->delta_capacity;
    DevModel& D    = m->dev;
    D.S = d->S, D.A = d->A, D.O = d->O;
    D.FS = d->n_state_features, D.FO = d->n_obs_features, D.J = D.FS + D.FO;
    for (int f = 0; f < D.FS; ++f) D.feat_s[f] = d->state_feature_sizes[f];
    for (int f = 0; f < D.FO; ++f) D.feat_o[f] = d->obs_feature_sizes[f];
    D.step_s[D.FS - 1] = 1;
    for (int f = D.FS - 2; f >= 0; --f) D.step_s[f] = D.step_s[f + 1] * D.feat_s[f + 1];
    D.step_o[D.FO - 1] = 1;
    for (int f = D.FO - 2; f >= 0; --f) D.step_o[f] = D.step_o[f + 1] * D.feat_o[f + 1];
    auto log2_exact = [](int v) {
        int k = 0;
        while ((1 << k) < v) ++k;
        return ((1 << k) == v) ? k : -1;
    };
    D.pow2_s = D.pow2_o = 1;
    for (int f = 0; f < D.FS; ++f) D.pow2_s &= log2_exact(D.feat_s[f]) >= 0, D.shift_s[f] = std::max(0, log2_exact(D.step_s[f]));
    for (int f = 0; f < D.FO; ++f) D.pow2_o &= log2_exact(D.feat_o[f]) >= 0, D.shift_o[f] = std::max(0, log2_exact(D.step_o[f]));
    D.tabular = d->tabular, D.domain = d->domain, D.action_draw = d->action_draw;
    D.sampled = d->dirichlet_sampling != 0;
    for (int f = 0; f < D.FS; ++f) m->long_rows |= D.feat_s[f] > 4;
    for (int f = 0; f < D.FO; ++f) m->long_rows |= D.feat_o[f] > 4;
    memcpy(D.dom_ip, d->dom_ip, sizeof(D.dom_ip));
    memcpy(D.dom_dp, d->dom_dp, sizeof(D.dom_dp));
    D.start_kind = d->start_kind;
    memcpy(D.start_ip, d->start_ip, sizeof(D.start_ip));
    D.start_total = d->start_total;

    int rc = FBA_OK;
    if (d->domain == FBA_DOM_TABLE)
    {
        if (!rc) rc = to_device(m, d->rew_sa, (size_t)d->S * d->A, &D.rew_sa);
        if (!rc) rc = to_device(m, d->rew_as2, (size_t)d->S * d->A, &D.rew_as2);
        if (!rc) rc = to_device(m, d->term_sa, (size_t)d->S * d->A, &D.term_sa);
        if (!rc) rc = to_device(m, d->term_as2, (size_t)d->S * d->A, &D.term_as2);
    }
    if (!rc && d->start_kind == FBA_START_CATEGORICAL)
    {
        if (!d->start_values) ctx->err = "model: categorical start needs start_values", rc = FBA_ERR_INVALID;
        else
            rc = to_device(m, d->start_values, (size_t)d->start_ip[0], &D.start_values);
    }
    if (!rc && d->start_kind == FBA_START_SLOW2)
    {
        if (!d->start_table) ctx->err = "model: slow2 start needs start_table", rc = FBA_ERR_INVALID;
        else
            rc = to_device(m, d->start_table, (size_t)d->start_ip[0] * d->start_ip[1], &D.start_table);
    }
    if (!rc)
    {
        size_t const nn = (size_t)max_structures * D.A * D.J;
        m->h_nodes.resize(nn);
        cudaError_t e = cudaMalloc(&m->d_nodes, nn * sizeof(Node));
        if (e == cudaSuccess) e = cudaMalloc(&m->d_sizes, (size_t)max_structures * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(m->d_sizes, 0, (size_t)max_structures * sizeof(int));
        if (e != cudaSuccess) ctx->err = std::string("model alloc: ") + cudaGetErrorString(e), rc = FBA_ERR_CUDA;
        D.nodes       = m->d_nodes;
        D.struct_size = m->d_sizes;
    }
    if (rc)
    {
        fba_model_destroy(m);
        return rc;
    }
    *out = m;
    return FBA_OK;
}

extern "C" void fba_model_destroy(fba_model* m)
{
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    for (auto p : m->owned) cudaFree(p);
    cudaFree(m->d_nodes);
    cudaFree(m->d_sizes);
    delete m;
}

static int add_structures(fba_model* m, int32_t n, const uint32_t* t_par, const uint32_t* o_par,
                          int32_t* ids_out, bool upload);

// uploads the node tables of structures [first_new, n_structs) to the device
static int upload_structures(fba_model* m, int first_new)
{
    fba_ctx* ctx = m->ctx;
    if (m->n_structs <= first_new) return FBA_OK;
    size_t const per = (size_t)m->dev.A * m->dev.J;
    CU(ctx, cudaMemcpy(m->d_nodes + first_new * per, m->h_nodes.data() + first_new * per,
                       (m->n_structs - first_new) * per * sizeof(Node), cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(m->d_sizes + first_new, m->sizes.data() + first_new,
                       (m->n_structs - first_new) * sizeof(int), cudaMemcpyHostToDevice));
    return FBA_OK;
}

extern "C" int fba_model_add_structures(fba_model* m, int32_t n, const uint32_t* t_par,
                                        const uint32_t* o_par, int32_t* ids_out)
{
    return add_structures(m, n, t_par, o_par, ids_out, true);
}

static int add_structures(fba_model* m, int32_t n, const uint32_t* t_par, const uint32_t* o_par,
                          int32_t* ids_out, bool upload)
{
    if (!m || !t_par || !o_par || n < 0) return FBA_ERR_INVALID;
    fba_ctx* ctx      = m->ctx;
    DevModel const& D = m->dev;
    CU(ctx, cudaSetDevice(ctx->device));
    int const nT = D.A * D.FS, nO = D.A * D.FO;
    int const first_new = m->n_structs;
    uint32_t const full = (D.FS >= 32) ? 0xffffffffu : ((1u << D.FS) - 1u);
    int rc = FBA_OK;
    std::vector<Node> fresh((size_t)D.A * D.J);
    for (int k = 0; k < n && rc == FBA_OK; ++k)
    {
        const uint32_t* tp = t_par + (size_t)k * nT;
        const uint32_t* op = o_par + (size_t)k * nO;
        std::string& key = m->key_buf;
        key.assign((const char*)tp, nT * sizeof(uint32_t));
        key.append((const char*)op, nO * sizeof(uint32_t));
        auto it = m->index.find(key);
        int id;
        if (it != m->index.end()) id = it->second;
        else
        {
            // lay the structure out first; it is registered only once it is known to be valid
            long long off = 0;
            for (int a = 0; a < D.A && rc == FBA_OK; ++a)
                for (int j = 0; j < D.J; ++j)
                {
                    uint32_t const par = (j < D.FS) ? tp[a * D.FS + j] : op[a * D.FO + (j - D.FS)];
                    long long cfgs     = 1;
                    for (int f = 0; f < D.FS; ++f)
                        if (par & (1u << f)) cfgs *= D.feat_s[f];
                    int const range        = (j < D.FS) ? D.feat_s[j] : D.feat_o[j - D.FS];
                    fresh[a * D.J + j].par = par;
                    fresh[a * D.J + j].off = (int32_t)off;
                    off += cfgs * range;
                    if ((par & ~full) != 0) ctx->err = "structure: parent mask names a missing feature", rc = FBA_ERR_INVALID;
                    else if (off >= (1ll << 31))
                        ctx->err = "structure: count block exceeds 2^31 cells", rc = FBA_ERR_INVALID;
                    if (rc != FBA_OK) break;
                }
            if (rc == FBA_OK && m->n_structs >= m->max_structs)
            {
                ctx->err = "structure table full (" + std::to_string(m->max_structs) + ")";
                rc       = FBA_ERR_CAPACITY;
            }
            if (rc != FBA_OK) break;
            id            = m->n_structs++;
            m->index[key] = id;
            m->t_par.insert(m->t_par.end(), tp, tp + nT);
            m->o_par.insert(m->o_par.end(), op, op + nO);
            std::copy(fresh.begin(), fresh.end(), m->h_nodes.begin() + (size_t)id * D.A * D.J);
            m->sizes.push_back((int)off);
        }
        if (ids_out) ids_out[k] = id;
    }
    // structures registered before a failing one stay registered, so they are uploaded either way
    // (only entries [first_new, n_structs) are written: no particle names them yet, so kernels in
    // flight on the context's stream never read what this copy writes)
    if (upload)
    {
        std::string const keep = ctx->err;
        int const up           = upload_structures(m, first_new);
        if (rc != FBA_OK) ctx->err = keep;
        else
            rc = up;
    }
    return rc;
}

extern "C" int32_t fba_model_num_structures(const fba_model* m)
{
    return m ? m->n_structs : 0;
}
extern "C" int64_t fba_model_structure_size(const fba_model* m, int32_t id)
{
    return (m && id >= 0 && id < m->n_structs) ? m->sizes[id] : -1;
}
extern "C" int fba_model_get_structure(const fba_model* m, int32_t id, uint32_t* t_par, uint32_t* o_par)
{
    if (!m || id < 0 || id >= m->n_structs) return FBA_ERR_INVALID;
    int const nT = m->dev.A * m->dev.FS, nO = m->dev.A * m->dev.FO;
    if (t_par) memcpy(t_par, m->t_par.data() + (size_t)id * nT, nT * sizeof(uint32_t));
    if (o_par) memcpy(o_par, m->o_par.data() + (size_t)id * nO, nO * sizeof(uint32_t));
    return FBA_OK;
}

// ------------------------------------------------------------------------------------------------
// belief
// ------------------------------------------------------------------------------------------------
extern "C" int fba_belief_create(fba_ctx* ctx, fba_model* m, int64_t N, int64_t stride, int32_t weighted,
                                 fba_belief** out)
{
    if (!ctx || !m || !out) return FBA_ERR_INVALID;
    *out = nullptr;
    // the reference's belief constructors throw on n < 1 (BAImportanceSampling.cpp:19-29)
    REQUIRE(ctx, N >= 1, "cannot initiate belief with n " + std::to_string(N));
    REQUIRE(ctx, N < (1ll << 31), "belief: at most 2^31-1 particles per GPU");
    REQUIRE(ctx, m->n_structs >= 1, "belief: register at least one structure first");
    long long need = 0;
    for (int s : m->sizes) need = std::max<long long>(need, s);
    if (stride <= 0) stride = need;
    stride = (stride + 3) & ~3ll;
    REQUIRE(ctx, stride >= need, "belief: stride smaller than the largest registered structure");
    long long const lstride = stride;
    if (m->delta_cap > 0)
    {
        REQUIRE(ctx, m->n_structs == 1 || !m->dev.tabular, "belief: tabular base+delta storage needs exactly one structure");
        if (m->dev.tabular) stride = ((long long)m->delta_cap + 1 + 3) & ~3ll; // header + increments, in 4-byte words
        else // journal: a 4-word header, then delta_cap / J updates of J entries padded to a multiple of 4
            stride = kJournalHeader + (long long)(m->delta_cap / m->dev.J) * ((m->dev.J + 1 + 3) & ~3);
    }
    CU(ctx, cudaSetDevice(ctx->device));

    auto b       = new fba_belief();
    b->ctx       = ctx;
    b->m         = m;
    b->N         = N;
    b->stride    = stride;
    b->lstride   = lstride;
    b->delta_cap = m->delta_cap;
    b->weighted = weighted != 0;
    // only the front buffer: the back buffer (full-copy resampling, rejection sampling) is allocated
    // on first use (ensure_next), so a production belief that resamples in place can fill the HBM
    cudaError_t e = cudaMalloc(&b->counts[0], (size_t)N * stride * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&b->state[0], (size_t)N * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->sid[0], (size_t)N * sizeof(int));
    long long const n_tiles = (N + kTile - 1) / kTile;
    if (e == cudaSuccess) e = cudaMalloc(&b->w, (size_t)N * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->aux, (size_t)N * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->tile, (size_t)n_tiles * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->scal, 4 * sizeof(double));
    b->anc_cap = N + N / 8 + 1024; // room for an over-quota shard's surplus offspring
    if (e == cudaSuccess) e = cudaMalloc(&b->anc, (size_t)b->anc_cap * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_total, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->noff, (size_t)N * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->escan, (size_t)N * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->dead, (size_t)N * sizeof(int));
    b->src_cap = 2 * N + 1024; // extras <= offspring total; a shard's quota never exceeds 2N here
    if (e == cudaSuccess) e = cudaMalloc(&b->src_of, (size_t)b->src_cap * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->totals, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->stats, 4 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemset(b->stats, 0, 4 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_quota, sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&b->tile_pairs, (size_t)n_tiles * sizeof(int2));
    if (e != cudaSuccess)
    {
        ctx->err = std::string("belief alloc: ") + cudaGetErrorString(e);
        fba_belief_destroy(b);
        return FBA_ERR_CUDA;
    }
    *out = b;
    return FBA_OK;
}

extern "C" void fba_belief_destroy(fba_belief* b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (int k = 0; k < 2; ++k)
    {
        cudaFree(b->counts[k]);
        cudaFree(b->state[k]);
        cudaFree(b->sid[k]);
    }
    cudaFree(b->base), cudaFree(b->d_proto_sid);
    cudaFreeHost(b->h_sid), cudaFreeHost(b->h_state);
    cudaFree(b->d_jobs);
    cudaFree(b->rs_src), cudaFree(b->rs_state), cudaFree(b->rs_rec), cudaFree(b->rs_flag), cudaFree(b->rs_pos), cudaFree(b->rs_tiles);
    cudaFree(b->w), cudaFree(b->aux), cudaFree(b->tile), cudaFree(b->scal), cudaFree(b->anc);
    cudaFree(b->att_src), cudaFree(b->att_state), cudaFree(b->att_accept), cudaFree(b->att_pos);
    cudaFree(b->roll_p), cudaFree(b->roll_s), cudaFree(b->roll_d), cudaFree(b->roll_r);
    cudaFree(b->step_i), cudaFree(b->step_r);
    for (auto p : b->opened) cudaIpcCloseMemHandle(p);
    cudaFree(b->d_plan);
    cudaFree(b->att_tiles);
    cudaFree(b->att_rec), cudaFree(b->d_total), cudaFree(b->xport), cudaFree(b->import_buf);
    cudaFree(b->src_of), cudaFree(b->d_quota), cudaFree(b->stats), cudaFree(b->noff), cudaFree(b->escan), cudaFree(b->tile_pairs);
    if (b->p2p_block) cudaFree(b->p2p_block); // b->dead and b->totals live inside it
    else
        cudaFree(b->dead), cudaFree(b->totals);
    cudaFree(b->d_step);
    cudaFree(b->seg_start), cudaFree(b->seg_end), cudaFree(b->seg_delta), cudaFree(b->seg_tie), cudaFree(b->chain_stats);
    delete b;
}

extern "C" int64_t fba_belief_size(const fba_belief* b)
{
    return b ? b->N : 0;
}
extern "C" int64_t fba_belief_stride(const fba_belief* b)
{
    return b ? b->lstride : 0;
}
extern "C" void* fba_belief_counts_ptr(fba_belief* b)
{
    return b ? b->counts[b->cur] : nullptr;
}
extern "C" void* fba_belief_state_ptr(fba_belief* b)
{
    return b ? b->state[b->cur] : nullptr;
}
extern "C" void* fba_belief_weight_ptr(fba_belief* b)
{
    if (b) b->w_escaped = true, b->uniform_now = false; // the host language may write through this pointer
    return b ? b->w : nullptr;
}
extern "C" void* fba_belief_aux_ptr(fba_belief* b)
{
    return b ? b->aux : nullptr;
}
extern "C" void* fba_belief_scalars_ptr(fba_belief* b)
{
    return b ? b->scal : nullptr;
}

// WeightedFilter::_total_weight after N x add(s, 1/N) (WeightedFilter.cpp:60-66): data independent
static double uniform_total(fba_belief* b)
{
    if (b->uniform_total < 0)
    {
        double const w = 1.0 / (double)b->N;
        volatile double acc = 0.0; // volatile: keep the sequential, separately rounded adds
        for (long long i = 0; i < b->N; ++i) acc = acc + w;
        b->uniform_total = acc;
    }
    return b->uniform_total;
}

static void weights_became_uniform(fba_belief* b)
{
    b->total_weight = uniform_total(b);
    b->suffix_valid = false;
    b->cdf_valid    = false;
    b->uniform_now  = !b->w_escaped;
}

// device scratch that dies with the call, whichever way the call returns
template<class T>
struct DevTmp
{
    T* p = nullptr;
    ~DevTmp() { cudaFree(p); }
    operator T*() const { return p; }
    T** operator&() { return &p; }
    DevTmp()              = default;
    DevTmp(DevTmp const&) = delete;
    DevTmp& operator=(DevTmp const&) = delete;
};

// the same from the device's stream-ordered memory pool (cudaMallocAsync on the context's stream): for the short
// calls of the MCMC beliefs, which would otherwise spend most of their time in cudaMalloc / cudaFree. Freed in
// stream order when the call returns; the pool keeps up to 256 MB cached (fba_ctx_create).
template<class T>
struct PoolTmp
{
    T* p            = nullptr;
    cudaStream_t st = nullptr;
    ~PoolTmp()
    {
        if (p) cudaFreeAsync(p, st);
    }
    operator T*() const { return p; }
    PoolTmp()               = default;
    PoolTmp(PoolTmp const&) = delete;
    PoolTmp& operator=(PoolTmp const&) = delete;
    int alloc(fba_ctx* ctx, size_t n)
    {
        st = ctx->stream;
        CU(ctx, cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st));
        return FBA_OK;
    }
};

// base+delta storage: the prior prototypes become the shared base tables
static int install_base_tables(fba_belief* b, int n_protos, const float* proto_counts, const int32_t* proto_struct_id)
{
    fba_ctx* ctx = b->ctx;
    size_t const n = (size_t)n_protos * b->lstride;
    cudaFree(b->base), cudaFree(b->d_proto_sid);
    b->base = nullptr, b->d_proto_sid = nullptr;
    b->h_proto_sid.assign(proto_struct_id, proto_struct_id + n_protos);
    CU(ctx, cudaMalloc(&b->d_proto_sid, (size_t)n_protos * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(b->d_proto_sid, b->h_proto_sid.data(), (size_t)n_protos * sizeof(int), cudaMemcpyHostToDevice,
                            ctx->stream));
    CU(ctx, cudaMalloc(&b->base, n * sizeof(float)));
    CU(ctx, cudaMemcpyAsync(b->base, proto_counts, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    b->h_base.assign(proto_counts, proto_counts + n);
    b->n_bases = n_protos;
    return FBA_OK;
}

extern "C" int fba_belief_init(fba_belief* b, int32_t n_protos, const int32_t* proto_struct_id,
                               const float* proto_counts, const int32_t* particle_proto,
                               const int32_t* particle_state)
{
    if (!b) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, n_protos >= 1 && proto_struct_id && proto_counts && particle_state,
            "belief_init: prototypes and particle states are required");
    for (int p = 0; p < n_protos; ++p)
        REQUIRE(ctx, proto_struct_id[p] >= 0 && proto_struct_id[p] < b->m->n_structs,
                "belief_init: unknown structure id");
    CU(ctx, cudaSetDevice(ctx->device));
    if (particle_proto)
        for (long long i = 0; i < b->N; ++i)
            REQUIRE(ctx, particle_proto[i] >= 0 && particle_proto[i] < n_protos, "belief_init: bad prototype index");
    for (long long i = 0; i < b->N; ++i)
        REQUIRE(ctx, particle_state[i] >= 0 && particle_state[i] < b->m->dev.S, "belief_init: state out of range");
    DevTmp<float> d_protos;
    DevTmp<int> d_psid, d_pp, d_ps;
    size_t const pc = (size_t)n_protos * b->lstride;
    CU(ctx, cudaMalloc(&d_ps, (size_t)b->N * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(d_ps, particle_state, (size_t)b->N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (particle_proto)
    {
        CU(ctx, cudaMalloc(&d_pp, (size_t)b->N * sizeof(int)));
        CU(ctx, cudaMemcpyAsync(d_pp, particle_proto, (size_t)b->N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (b->delta_cap > 0)
    {
        int const rc = install_base_tables(b, n_protos, proto_counts, proto_struct_id);
        if (rc) return rc;
        LAUNCH(ctx, k_init_delta, blocks_for(b->N), kThreads, b->counts[b->cur], b->stride, b->state[b->cur],
               b->sid[b->cur], b->weighted ? b->w : nullptr, b->N, d_pp, d_ps, b->m->dev.tabular ? 0 : kJournalHeader - 1);
        b->delta_used = 0;
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        weights_became_uniform(b);
        return FBA_OK;
    }
    CU(ctx, cudaMalloc(&d_protos, pc * sizeof(float)));
    CU(ctx, cudaMalloc(&d_psid, n_protos * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(d_protos, proto_counts, pc * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_psid, proto_struct_id, n_protos * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_init_from_protos, stream_grid(ctx, b->N), kThreads, b->counts[b->cur], b->stride,
           b->state[b->cur], b->sid[b->cur], b->weighted ? b->w : nullptr, b->N, d_protos, d_psid, d_pp,
           d_ps);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    weights_became_uniform(b);
    return FBA_OK;
}

extern "C" int fba_belief_init_sampled(fba_belief* b, int32_t n_protos, const int32_t* proto_struct_id,
                                       const float* proto_counts, const double* proto_probs, fba_rng* rng)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "belief_init_sampled: PHILOX mode only");
    REQUIRE(ctx, n_protos >= 1 && proto_struct_id && proto_counts, "belief_init_sampled: prototypes required");
    for (int p = 0; p < n_protos; ++p)
        REQUIRE(ctx, proto_struct_id[p] >= 0 && proto_struct_id[p] < b->m->n_structs,
                "belief_init_sampled: unknown structure id");
    CU(ctx, cudaSetDevice(ctx->device));
    DevTmp<float> d_protos;
    DevTmp<int> d_psid, d_pp, d_ps;
    DevTmp<double> d_cdf;
    size_t const pc = (size_t)n_protos * b->lstride;
    CU(ctx, cudaMalloc(&d_protos, pc * sizeof(float)));
    CU(ctx, cudaMalloc(&d_psid, n_protos * sizeof(int)));
    CU(ctx, cudaMalloc(&d_pp, (size_t)b->N * sizeof(int)));
    CU(ctx, cudaMalloc(&d_ps, (size_t)b->N * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(d_protos, proto_counts, pc * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_psid, proto_struct_id, n_protos * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double> cdf;
    if (proto_probs)
    {
        double acc = 0, tot = 0;
        for (int p = 0; p < n_protos; ++p) tot += proto_probs[p];
        for (int p = 0; p < n_protos; ++p) cdf.push_back((acc += proto_probs[p]) / tot);
        CU(ctx, cudaMalloc(&d_cdf, n_protos * sizeof(double)));
        CU(ctx, cudaMemcpyAsync(d_cdf, cdf.data(), n_protos * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    LAUNCH(ctx, k_draw_init<false>, blocks_for(b->N), kThreads, b->m->dev, b->N, n_protos, d_cdf, d_pp,
           d_ps, philox_args(rng));
    if (b->delta_cap > 0)
    {
        int const rc = install_base_tables(b, n_protos, proto_counts, proto_struct_id);
        if (rc) return rc;
        LAUNCH(ctx, k_init_delta, blocks_for(b->N), kThreads, b->counts[b->cur], b->stride, b->state[b->cur],
               b->sid[b->cur], b->weighted ? b->w : nullptr, b->N, d_pp, d_ps, b->m->dev.tabular ? 0 : kJournalHeader - 1);
        b->delta_used = 0;
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        b->total_weight = 1.0;
        b->suffix_valid = b->cdf_valid = false;
        b->uniform_now  = !b->w_escaped; // weights 1 / N
        return FBA_OK;
    }
    LAUNCH(ctx, k_init_from_protos, stream_grid(ctx, b->N), kThreads, b->counts[b->cur], b->stride,
           b->state[b->cur], b->sid[b->cur], b->weighted ? b->w : nullptr, b->N, d_protos, d_psid, d_pp,
           d_ps);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    b->total_weight = 1.0;
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now  = !b->w_escaped; // weights 1 / N
    return FBA_OK;
}

extern "C" int fba_belief_upload(fba_belief* b, int64_t first, int64_t count, const int32_t* state,
                                 const int32_t* struct_id, const float* counts, const double* w)
{
    if (!b) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= b->N, "belief_upload: range out of bounds");
    CU(ctx, cudaSetDevice(ctx->device));
    if (state)
    {
        for (int64_t i = 0; i < count; ++i)
            REQUIRE(ctx, state[i] >= 0 && state[i] < b->m->dev.S, "belief_upload: state out of range");
        CU(ctx, cudaMemcpyAsync(b->state[b->cur] + first, state, count * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    REQUIRE(ctx, !(b->delta_cap > 0 && (counts || struct_id)),
            "belief_upload: base+delta beliefs take their counts from fba_belief_init prototypes");
    if (struct_id)
    {
        for (int64_t i = 0; i < count; ++i)
            REQUIRE(ctx, struct_id[i] >= 0 && struct_id[i] < b->m->n_structs, "belief_upload: unknown structure id");
        CU(ctx, cudaMemcpyAsync(b->sid[b->cur] + first, struct_id, count * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (counts)
        CU(ctx, cudaMemcpyAsync(b->counts[b->cur] + first * b->stride, counts,
                                (size_t)count * b->stride * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (w && b->weighted)
    {
        CU(ctx, cudaMemcpyAsync(b->w + first, w, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        b->suffix_valid = b->cdf_valid = false;
        b->uniform_now = false;
        if (first == 0 && count == b->N)
        { // WeightedFilter::_total_weight = the sequential sum of the weights as added
            volatile double acc = 0.0;
            for (int64_t i = 0; i < count; ++i) acc = acc + w[i];
            b->total_weight = acc;
        }
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

extern "C" int fba_belief_download(fba_belief* b, int64_t first, int64_t count, int32_t* state,
                                   int32_t* struct_id, float* counts, double* w)
{
    if (!b) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= b->N, "belief_download: range out of bounds");
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<int> sid_tmp;
    int* sid_host = struct_id;
    if (counts && !sid_host)
    {
        sid_tmp.resize(count);
        sid_host = sid_tmp.data();
    }
    if (state)
        CU(ctx, cudaMemcpyAsync(state, b->state[b->cur] + first, count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (sid_host)
        CU(ctx, cudaMemcpyAsync(sid_host, b->sid[b->cur] + first, count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int> blocks;
    if (counts && b->delta_cap > 0)
    {
        blocks.resize((size_t)count * b->stride);
        CU(ctx, cudaMemcpyAsync(blocks.data(), b->counts[b->cur] + first * b->stride,
                                (size_t)count * b->stride * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    } else if (counts)
        CU(ctx, cudaMemcpyAsync(counts, b->counts[b->cur] + first * b->stride,
                                (size_t)count * b->stride * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (w && b->weighted)
        CU(ctx, cudaMemcpyAsync(w, b->w + first, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (counts && b->delta_cap > 0)
    { // dense view = the particle's base table + its increments, each exactly +1.0f in time order
        for (int64_t i = 0; i < count; ++i)
        {
            float* out = counts + i * b->lstride;
            memcpy(out, b->h_base.data() + (size_t)sid_host[i] * b->lstride, (size_t)b->lstride * sizeof(float));
            const int* blk = blocks.data() + (size_t)i * b->stride;
            auto plus_one  = [&](int cell) {
                volatile float v = out[cell];
                v                = v + 1.0f;
                out[cell]        = v;
            };
            if (b->m->dev.tabular)
                for (int e = 1; e <= blk[0]; ++e) plus_one(blk[e]);
            else
            { // journal layout: 4-word header, updates of J entries padded to Jp
                int const J = b->m->dev.J, Jp = (J + 1 + 3) & ~3, nu = (blk[0] - (kJournalHeader - 1)) / Jp;
                for (int u = 0; u < nu; ++u)
                    for (int j = 0; j < J; ++j) plus_one(blk[kJournalHeader + u * Jp + j]);
            }
            if (struct_id) struct_id[i] = b->h_proto_sid[(size_t)sid_host[i]]; // sid holds the base table (prototype) id
        }
        return FBA_OK;
    }
    if (counts)
    { // cells past a particle's own structure are padding: report zeros (copies skip them)
        for (int64_t i = 0; i < count; ++i)
        {
            long long const used = b->m->sizes[sid_host[i]];
            if (used < b->stride)
                memset(counts + i * b->stride + used, 0, (size_t)(b->stride - used) * sizeof(float));
        }
    }
    return FBA_OK;
}

extern "C" int fba_belief_total_weight(fba_belief* b, double* total)
{
    if (!b || !total) return FBA_ERR_INVALID;
    *total = b->total_weight;
    return FBA_OK;
}

// base + delta / base + journal storage -> dense private blocks, in place of the belief's storage
static int compact_to_dense(fba_belief* b)
{
    fba_ctx* ctx = b->ctx;
    if (b->delta_cap == 0) return FBA_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    size_t free_b = 0, total_b = 0;
    CU(ctx, cudaMemGetInfo(&free_b, &total_b));
    size_t const need = (size_t)b->N * b->lstride * sizeof(float);
    if (need + (512ull << 20) > free_b)
    {
        ctx->err = "compact: dense storage of this belief needs " + std::to_string(need >> 20) + " MB, "
                   + std::to_string(free_b >> 20) + " MB are free";
        return FBA_ERR_CAPACITY;
    }
    float* dense = nullptr;
    CU(ctx, cudaMalloc(&dense, need));
    LAUNCH(ctx, k_compact, stream_grid(ctx, b->N), kThreads, (const float*)b->base, b->lstride,
           (const float*)b->counts[b->cur], b->stride, b->sid[b->cur], (const int*)b->d_proto_sid, dense, b->N,
           b->m->dev.tabular ? 0 : b->m->dev.J);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    int const nx = b->cur ^ 1;
    cudaFree(b->counts[b->cur]), cudaFree(b->counts[nx]), cudaFree(b->state[nx]), cudaFree(b->sid[nx]);
    b->counts[b->cur] = dense;
    b->counts[nx] = nullptr, b->state[nx] = nullptr, b->sid[nx] = nullptr; // the back buffer returns on first use
    cudaFree(b->base), cudaFree(b->d_proto_sid);
    b->base = nullptr, b->d_proto_sid = nullptr;
    b->h_base.clear(), b->h_base.shrink_to_fit(), b->h_proto_sid.clear();
    b->n_bases    = 0;
    b->stride     = b->lstride;
    b->delta_cap  = 0;
    b->delta_used = 0;
    return FBA_OK;
}

extern "C" int fba_belief_compact(fba_belief* b)
{
    if (!b) return FBA_ERR_INVALID;
    return compact_to_dense(b);
}

extern "C" int32_t fba_belief_delta_capacity(const fba_belief* b)
{
    return b ? b->delta_cap : 0;
}

// ---- importance sampling ---------------------------------------------------------------------

static int propose(fba_belief* b, int a, int o, fba_rng* rng, unsigned long long stream_base)
{
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, b->weighted, "importance sampling needs a weighted belief");
    REQUIRE(ctx, !(D.sampled && rng->mode == FBA_RNG_REPLAY),
            "sampled-Dirichlet models run in PHILOX mode (their gamma draws cannot replay libm bit for bit)");
    REQUIRE(ctx, a >= 0 && a < D.A, "action out of range");
    REQUIRE(ctx, o >= 0 && o < D.O, "observation out of range");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (b->delta_cap > 0 && b->delta_used + D.J > b->delta_cap && ctx->auto_compact && !b->peers_open)
    { // the increment lists are full: continue on dense private blocks (if they fit; otherwise the error below)
        rc = compact_to_dense(b);
        if (rc && rc != FBA_ERR_CAPACITY) return rc;
    }
    if (b->delta_cap > 0)
    {
        // every particle's increment list grows by J per update and copies keep the length, so the host
        // knows when the lists are full without asking the device
        if (b->delta_used + D.J > b->delta_cap)
        {
            ctx->err = "base+delta particles ran out of increment slots after " + std::to_string(b->delta_used / D.J)
                       + " updates (raise fba_model_desc.delta_capacity)";
            return FBA_ERR_CAPACITY;
        }
        b->delta_used += D.J;
        bool binary = !D.tabular && !D.sampled && D.FS <= FBA_MAX_FEATURES && D.FO <= 2 && b->delta_cap / D.J <= 255;
        for (int f = 0; f < D.FS; ++f) binary = binary && D.feat_s[f] == 2;
        for (int f = 0; f < D.FO; ++f) binary = binary && D.feat_o[f] == 2;
        if (binary) // the register-only step histograms whole observation CPTs: at most 8 cells = 2 parents
            for (size_t k = 0; k < b->m->o_par.size(); ++k) binary = binary && __builtin_popcount(b->m->o_par[k]) <= 2;
#define PROPOSE_DELTA_ARGS(rargs)                                                                              \
    D, b->base, b->lstride, b->counts[b->cur], b->stride, b->delta_cap, b->state[b->cur], b->sid[b->cur],         \
        (const int*)b->d_proto_sid, b->w, b->N, a, o, rargs, ctx->d_flag
        bool const replay = rng->mode == FBA_RNG_REPLAY;
        long long const per = 2ll * D.J, need = per * b->N;
        if (replay)
        {
            if ((rc = stage_words(ctx, rng, need))) return rc;
            if ((rc = clear_flag(ctx))) return rc;
        }
        RngArgs const ra = replay ? replay_args(ctx, need, per, false) : philox_args(rng, stream_base);
        if (D.tabular) LAUNCH_DELTA(ctx, k_propose_delta, replay, blocks_for(b->N), kThreads, PROPOSE_DELTA_ARGS(ra));
        else if (binary)
        { // all particles hold the same number of updates: (delta_used - J) / J before this one
            int const nu   = (int)((b->delta_used - D.J) / D.J);
            int const grid = blocks_for(b->N, kStageWarps * 32);
            if (replay)
                LAUNCH(ctx, (k_propose_journal_staged<true>), grid, kStageWarps * 32, PROPOSE_DELTA_ARGS(ra), nu);
            else
                LAUNCH(ctx, (k_propose_journal_staged<false>), grid, kStageWarps * 32, PROPOSE_DELTA_ARGS(ra), nu);
        } else if (replay)
            LAUNCH(ctx, (k_propose_journal<true, false>), blocks_for(b->N), kThreads, PROPOSE_DELTA_ARGS(ra));
        else if (D.sampled)
            LAUNCH(ctx, (k_propose_journal<false, true>), blocks_for(b->N), kThreads, PROPOSE_DELTA_ARGS(ra));
        else
            LAUNCH(ctx, (k_propose_journal<false, false>), blocks_for(b->N), kThreads, PROPOSE_DELTA_ARGS(ra));
#undef PROPOSE_DELTA_ARGS
        if (replay) rng->cursor += need;
        b->suffix_valid = b->cdf_valid = false;
        b->uniform_now = false;
        return FBA_OK;
    }
    if (rng->mode == FBA_RNG_REPLAY)
    {
        // particle-major: particle i consumes J uniforms = 2J words (SURVEY.md §8 a')
        long long const per = 2ll * D.J, need = per * b->N;
        if ((rc = stage_words(ctx, rng, need))) return rc;
        if ((rc = clear_flag(ctx))) return rc;
        LAUNCH_RL(ctx, k_propose, true, b->m->long_rows, blocks_for(b->N), kThreads, D, b->counts[b->cur],
                  b->stride, b->state[b->cur], b->sid[b->cur], b->w, b->N, a, o,
                  replay_args(ctx, need, per, false), ctx->d_flag);
        rng->cursor += need;
    } else
    {
        LAUNCH_RL(ctx, k_propose, false, b->m->long_rows, blocks_for(b->N), kThreads, D, b->counts[b->cur],
                  b->stride, b->state[b->cur], b->sid[b->cur], b->w, b->N, a, o, philox_args(rng, stream_base),
                  ctx->d_flag);
    }
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now = false;
    return FBA_OK;
}

// native: tile sums -> scan -> (divide, cdf). Leaves the cdf of the normalised weights in aux.
static int native_normalize(fba_belief* b, bool use_device_total, double divide_by)
{
    fba_ctx* ctx      = b->ctx;
    int const n_tiles = (int)((b->N + kTile - 1) / kTile);
    LAUNCH(ctx, k_tile_sums, n_tiles, kThreads, b->w, b->N, b->tile);
    LAUNCH(ctx, k_scan_tile_sums, 1, kThreads, b->tile, n_tiles, b->scal);
    LAUNCH(ctx, k_scale_and_scan, n_tiles, kThreads, b->w, b->N, b->tile,
           use_device_total ? b->scal : (const double*)nullptr, divide_by, b->aux);
    b->cdf_valid    = true;
    b->suffix_valid = false;
    return FBA_OK;
}

static int read_scal(fba_belief* b)
{
    fba_ctx* ctx = b->ctx;
    CU(ctx, cudaMemcpyAsync(ctx->h_scal, b->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

// REPLAY: the reference's sequential chains (k_seq_normalize's A, B, C; only C if !do_normalise, with
// scal[1] = _total_weight given). Small beliefs: one thread walks them. Large beliefs: segments walked in
// parallel from speculated starts + a short serial shift / recompute pass — bit-identical results
// (fba_kernels.cuh, "REPLAY normalisation for LARGE beliefs").
static int replay_chains(fba_belief* b, bool do_normalise)
{
    fba_ctx* ctx = b->ctx;
    if (b->N < ctx->parallel_chains_min)
    {
        LAUNCH(ctx, k_seq_normalize, 1, kThreads, b->w, b->N, b->scal, b->aux, do_normalise ? 1 : 0);
        return FBA_OK;
    }
    int const n_seg = (int)((b->N + kTile - 1) / kTile);
    if (!b->seg_start)
    {
        CU(ctx, cudaMalloc(&b->seg_start, (size_t)n_seg * sizeof(double)));
        CU(ctx, cudaMalloc(&b->seg_end, (size_t)n_seg * sizeof(double)));
        CU(ctx, cudaMalloc(&b->seg_delta, (size_t)n_seg * sizeof(double)));
        CU(ctx, cudaMalloc(&b->seg_tie, (size_t)n_seg));
        CU(ctx, cudaMalloc(&b->chain_stats, 2 * sizeof(long long)));
        CU(ctx, cudaMemsetAsync(b->chain_stats, 0, 2 * sizeof(long long), ctx->stream));
    }
    int const seg_grid = blocks_for(n_seg, 64);
    double* const spec_total = b->scal + 3;
    const double* none       = nullptr;
    auto sum_chain = [&](double* result) -> int { // exact sequential sum of the current weights
        LAUNCH(ctx, k_scan_tile_sums, 1, kThreads, b->tile, n_seg, spec_total);
        LAUNCH(ctx, (k_chain_segments<false>), seg_grid, 64, b->w, b->N, b->tile, spec_total, none, b->seg_start,
               b->seg_end, b->seg_tie, (double*)nullptr);
        LAUNCH(ctx, (k_chain_fix<false>), 1, kThreads, b->w, b->N, b->seg_start, b->seg_end, b->seg_tie, none, result,
               (double*)nullptr, (double*)nullptr, b->chain_stats);
        return FBA_OK;
    };
    int rc;
    LAUNCH(ctx, k_tile_sums, n_seg, kThreads, b->w, b->N, b->tile);
    if (do_normalise)
    {
        if ((rc = sum_chain(b->scal))) return rc;                                                // A
        LAUNCH(ctx, k_divide_tile_sums, n_seg, kThreads, b->w, b->N, (const double*)b->scal, b->tile); // B
        if ((rc = sum_chain(b->scal + 1))) return rc;
    } else
        LAUNCH(ctx, k_scan_tile_sums, 1, kThreads, b->tile, n_seg, spec_total);
    // C: b->tile holds the tree-order prefix of the current weights, spec_total their tree-order total
    LAUNCH(ctx, (k_chain_segments<true>), seg_grid, 64, b->w, b->N, b->tile, spec_total, (const double*)(b->scal + 1),
           b->seg_start, b->seg_end, b->seg_tie, b->aux);
    LAUNCH(ctx, (k_chain_fix<true>), 1, kThreads, b->w, b->N, b->seg_start, b->seg_end, b->seg_tie,
           (const double*)(b->scal + 1), b->scal + 2, b->seg_delta, b->aux, b->chain_stats);
    LAUNCH(ctx, k_chain_apply_shift, blocks_for(b->N), kThreads, b->aux, b->N, (const double*)b->seg_delta);
    return FBA_OK;
}

extern "C" int fba_belief_update(fba_belief* b, int32_t a, int32_t o, fba_rng* rng, double* likelihood)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    int rc       = propose(b, a, o, rng, 0);
    if (rc) return rc;
    if (rng->mode == FBA_RNG_REPLAY)
    {
        if ((rc = replay_chains(b, true))) return rc;
        if ((rc = read_scal(b))) return rc;
        if ((rc = check_flag(ctx))) return rc;
        b->total_weight = ctx->h_scal[1];
        b->suffix_valid = true;
        if (likelihood) *likelihood = ctx->h_scal[0];
    } else
    {
        if ((rc = native_normalize(b, true, 1.0))) return rc;
        b->total_weight = 1.0;
        if (likelihood)
        {
            if ((rc = read_scal(b))) return rc;
            *likelihood = ctx->h_scal[0];
        }
    }
    return FBA_OK;
}

static void flip(fba_belief* b)
{
    b->cur ^= 1;
}

// the back buffer of the double-buffered operations, allocated on first use
static int ensure_next(fba_belief* b)
{
    fba_ctx* ctx = b->ctx;
    int const nx = b->cur ^ 1;
    if (b->counts[nx]) return FBA_OK;
    cudaError_t e = cudaMalloc(&b->counts[nx], (size_t)b->N * b->stride * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&b->state[nx], (size_t)b->N * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->sid[nx], (size_t)b->N * sizeof(int));
    if (e != cudaSuccess)
    {
        cudaFree(b->counts[nx]), cudaFree(b->state[nx]), cudaFree(b->sid[nx]);
        b->counts[nx] = nullptr, b->state[nx] = nullptr, b->sid[nx] = nullptr;
        ctx->err = std::string("belief back buffer: ") + cudaGetErrorString(e);
        return FBA_ERR_CUDA;
    }
    return FBA_OK;
}

// pick N ancestors from the current weights into b->anc
static int pick_ancestors(fba_belief* b, fba_rng* rng, long long n_out, long long words_per_item,
                          bool use_offsets, long long n_words)
{
    fba_ctx* ctx = b->ctx;
    int rc;
    if (rng->mode == FBA_RNG_REPLAY)
    {
        if (!b->suffix_valid)
        { // R for the current weights (e.g. uniform after a resample): chain C only
            double const tw = b->total_weight;
            CU(ctx, cudaMemcpyAsync(b->scal + 1, &tw, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            if ((rc = replay_chains(b, false))) return rc;
            b->suffix_valid = true;
            b->cdf_valid    = false;
            b->uniform_now = false;
        }
        if ((rc = clear_flag(ctx))) return rc;
        LAUNCH(ctx, k_pick_replay, blocks_for(n_out), kThreads, b->aux, b->N, b->scal,
               replay_args(ctx, n_words, words_per_item, use_offsets), b->anc, n_out, ctx->d_flag);
    } else
    {
        if (!b->cdf_valid)
            if ((rc = native_normalize(b, false, 1.0))) return rc;
        LAUNCH(ctx, k_pick_native, blocks_for(n_out), kThreads, b->aux, b->N, n_out, 1, philox_args(rng), b->anc);
    }
    return FBA_OK;
}

static int gather_into_next(fba_belief* b, long long n_out, bool copy_state)
{
    fba_ctx* ctx = b->ctx;
    int const nx = b->cur ^ 1;
    if (int const rc = ensure_next(b)) return rc;
    long long const stage_bytes = b->stride * (long long)sizeof(float);
    if (ctx->bulk_copy && b->delta_cap == 0 && stage_bytes >= 1024
        && kBulkStages * stage_bytes + kBulkStages * 8 <= 200 * 1024)
    {
        size_t const shmem = (size_t)kBulkStages * stage_bytes + kBulkStages * sizeof(unsigned long long);
        CU(ctx, cudaFuncSetAttribute(k_gather_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        int const grid = (int)std::min<long long>(n_out, (long long)ctx->sm_count * 2);
        if (ctx->profiling) profile_mark(ctx, "k_gather_bulk", true);
        k_gather_bulk<<<grid, 32, shmem, ctx->stream>>>(
            b->counts[b->cur], b->counts[nx], b->stride, b->state[b->cur], copy_state ? b->state[nx] : nullptr,
            b->sid[b->cur], b->sid[nx], b->m->d_sizes, b->weighted ? b->w : nullptr, 1.0 / (double)b->N, b->anc,
            n_out, (int)stage_bytes);
        if (ctx->profiling) profile_mark(ctx, "k_gather_bulk", false);
        ++ctx->launches;
        CU(ctx, cudaGetLastError());
        return FBA_OK;
    }
    LAUNCH(ctx, k_gather, stream_grid(ctx, n_out), kThreads, b->counts[b->cur], b->counts[nx], b->stride,
           b->state[b->cur], copy_state ? b->state[nx] : nullptr, b->sid[b->cur], b->sid[nx],
           b->m->d_sizes, b->weighted ? b->w : nullptr, 1.0 / (double)b->N, b->anc, n_out,
           b->delta_cap > 0 ? 1 : 0);
    return FBA_OK;
}

// PHILOX: systematic resampling to n_out offspring without moving survivors (see fba_kernels.cuh).
// Surplus offspring beyond the N slots go to the export buffer; missing ones leave dead slots empty.
static int ensure_export(fba_belief* b, long long records)
{
    fba_ctx* ctx = b->ctx;
    if (records <= b->xport_cap) return FBA_OK;
    cudaFree(b->xport);
    b->xport     = nullptr;
    b->xport_cap = 0;
    long long const cap = records + records / 4 + 64;
    CU(ctx, cudaMalloc(&b->xport, cap * fba_belief_record_bytes(b)));
    b->xport_cap = cap;
    return FBA_OK;
}

// n_out_dev != NULL: the quota is read from device memory (no host sync needed beforehand)
static int resample_inplace(fba_belief* b, fba_rng* rng, long long n_out, const long long* n_out_dev = nullptr,
                            bool p2p = false)
{
    fba_ctx* ctx = b->ctx;
    int rc;
    if (!b->cdf_valid)
        if ((rc = native_normalize(b, false, 1.0))) return rc;
    int const n_tiles = (int)((b->N + kTile - 1) / kTile);
    long long const rb = fba_belief_record_bytes(b);
    if (!n_out_dev)
        if ((rc = ensure_export(b, std::max(0ll, n_out - b->N)))) return rc;
    LAUNCH(ctx, k_offspring, n_tiles, kThreads, b->aux, b->N, n_out, n_out_dev, philox_args(rng), b->noff,
           b->tile_pairs);
    LAUNCH(ctx, k_scan_tile_pairs, 1, kThreads, b->tile_pairs, n_tiles, b->totals, b->stats);
    CU(ctx, cudaMemsetAsync(b->src_of, 0xFF, (size_t)b->src_cap * sizeof(int), ctx->stream));
    LAUNCH(ctx, k_offspring_apply, n_tiles, kThreads, b->noff, b->N, b->tile_pairs, b->dead, b->escan, b->src_of,
           b->src_cap);
    // peer-to-peer: tell every rank that this rank's dead-slot list is complete (senders wait for it)
    if (p2p) LAUNCH(ctx, k_p2p_signal, 1, 32, b->peers, b->p2p_ranks, b->p2p_rank, 0, b->d_step);
    LAUNCH(ctx, k_copy_inplace, stream_grid(ctx, b->N), kThreads, b->counts[b->cur], b->stride,
           b->state[b->cur], b->sid[b->cur], b->m->d_sizes, b->escan, b->src_of, b->N, b->dead, b->totals,
           b->xport, rb, b->xport_cap, b->stats, b->src_cap, p2p ? b->d_plan : (const long long*)nullptr,
           b->p2p_ranks, b->p2p_rank, b->peers, b->delta_cap > 0 ? 1 : 0, b->d_step);
    // ... and that this rank's peer stores are done; wait for the ranks that ship records here
    if (p2p) LAUNCH(ctx, k_p2p_signal, 1, 32, b->peers, b->p2p_ranks, b->p2p_rank, 1, b->d_step);
    LAUNCH(ctx, k_fill, blocks_for(b->N), kThreads, b->w, b->N, 1.0 / (double)b->N);
    if (p2p)
        LAUNCH(ctx, k_p2p_wait_landed, 1, 32, (const P2PCtrl*)b->p2p_block, b->p2p_ranks, b->p2p_rank, b->d_plan,
               b->d_step, b->peers.timeout_ns, b->stats);
    b->total_weight = 1.0;
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now  = !b->w_escaped; // k_fill: every weight is 1 / N
    b->inplace_last = true;
    return FBA_OK;
}

extern "C" int fba_belief_resample(fba_belief* b, fba_rng* rng)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, b->weighted, "resample needs a weighted belief");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (rng->mode == FBA_RNG_PHILOX && ctx->inplace_resample) return resample_inplace(b, rng, b->N);
    if (rng->mode == FBA_RNG_REPLAY)
    {
        long long const need = 2 * b->N; // N draws of one uniform (ImportanceSampler.hpp:79-82)
        if ((rc = stage_words(ctx, rng, need))) return rc;
        if ((rc = pick_ancestors(b, rng, b->N, 2, false, need))) return rc;
        rng->cursor += need;
    } else if ((rc = pick_ancestors(b, rng, b->N, 0, false, 0)))
        return rc;
    if ((rc = gather_into_next(b, b->N, true))) return rc;
    flip(b);
    if (rng->mode == FBA_RNG_REPLAY)
    {
        if ((rc = check_flag(ctx))) return rc;
        weights_became_uniform(b);
    } else
    {
        b->total_weight = 1.0;
        b->suffix_valid = b->cdf_valid = false;
        b->uniform_now  = !b->w_escaped; // the gather wrote 1 / N into every weight
    }
    return FBA_OK;
}

// A belief this small is launch-latency-bound (nine launches, ~40 us): one CTA does the whole
// update + resample with the very same *_body functions (k_runs_step with one run), bit-identical
// to the launch-per-phase path (tests/test_cuda_runs.py compares the two).
constexpr long long kFusedMaxParticles = 2048;

extern "C" int fba_belief_update_estimation(fba_belief* b, int32_t a, int32_t o, fba_rng* rng, double* likelihood)
{
    if (b && rng && rng->mode == FBA_RNG_PHILOX && b->ctx->fused_update && b->ctx->inplace_resample &&
        !b->ctx->profiling && b->weighted && b->delta_cap == 0 && b->N <= kFusedMaxParticles)
    {
        fba_ctx* ctx      = b->ctx;
        DevModel const& D = b->m->dev;
        REQUIRE(ctx, a >= 0 && a < D.A, "action out of range");
        REQUIRE(ctx, o >= 0 && o < D.O, "observation out of range");
        CU(ctx, cudaSetDevice(ctx->device));
        RunsArgs A{};
        A.counts = b->counts[b->cur], A.stride = b->stride, A.state = b->state[b->cur], A.sid = b->sid[b->cur];
        A.w = b->w, A.cdf = b->aux, A.noff = b->noff, A.escan = b->escan, A.dead = b->dead, A.src_of = b->src_of;
        A.tile = b->tile, A.tile_pairs = b->tile_pairs, A.totals = b->totals, A.scal = b->scal;
        A.n = b->N, A.n_tiles = (int)((b->N + kTile - 1) / kTile);
        A.action0 = a, A.observation0 = o;
        A.struct_size = b->m->d_sizes;
        A.stats       = b->stats;
        RngArgs ra{};
        ra.seed   = rng->seed;
        ra.offset = rng->offset;
        rng->offset += 2; // propose + resample, as the two calls below
        bool const lr = b->m->long_rows;
        if (D.sampled)
        {
            if (lr) LAUNCH(ctx, (k_runs_step<true, true, 0>), 1, kThreads, D, A, ra);
            else
                LAUNCH(ctx, (k_runs_step<false, true, 0>), 1, kThreads, D, A, ra);
        } else
        {
            if (lr) LAUNCH(ctx, (k_runs_step<true, false, 0>), 1, kThreads, D, A, ra);
            else
                LAUNCH(ctx, (k_runs_step<false, false, 0>), 1, kThreads, D, A, ra);
        }
        b->total_weight = 1.0;
        b->suffix_valid = b->cdf_valid = false;
        b->uniform_now  = !b->w_escaped; // the fused kernel ends with uniform weights, as the phases do
        b->inplace_last = true;
        if (likelihood)
        {
            int const rc = read_scal(b);
            if (rc) return rc;
            *likelihood = ctx->h_scal[0];
        }
        return FBA_OK;
    }
    int rc = fba_belief_update(b, a, o, rng, likelihood);
    if (rc) return rc;
    return fba_belief_resample(b, rng);
}

// words one sampleStartState consumes at stream position `pos` (host-side scan of the replay stream)
static long long start_state_words(const DevModel& D, const fba_rng* rng, long long pos)
{
    switch (D.start_kind)
    {
        case FBA_START_CONST: return 0;
        case FBA_START_BOOL: return 2;
        case FBA_START_UNIFORM_INT: {
            ReplayRng g(rng->words, pos, rng->n_words);
            draw_k(g, (uint32_t)D.start_ip[0]);
            return g.pos - pos;
        }
        case FBA_START_SLOW2: return 4;
        default: return 2;
    }
}

// with_resample: BAImportanceSampling::resetDomainStateDistribution on a weighted belief (N weighted draws
// first); otherwise BAPOMDP::resetDomainState on every particle where it is (BAPOMDP.cpp:69-77)
static int reset_states(fba_belief* b, fba_rng* rng, bool with_resample)
{
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    bool const weighted = b->weighted && with_resample;
    int const skip = weighted ? 2 : 0; // weighted: one pick uniform precedes the start draws
    if (rng->mode == FBA_RNG_REPLAY)
    {
        // item j: [weighted pick u] + start-state draws; the latter may vary in length
        std::vector<long long> off((size_t)b->N);
        long long pos = rng->cursor;
        for (long long j = 0; j < b->N; ++j)
        {
            off[j] = pos - rng->cursor;
            pos += skip;
            pos += start_state_words(D, rng, pos);
        }
        long long const need = pos - rng->cursor;
        if ((rc = stage_words(ctx, rng, need))) return rc;
        if ((rc = stage_offsets(ctx, off))) return rc;
        CU(ctx, cudaStreamSynchronize(ctx->stream)); // `off` is about to go out of scope
        if (weighted)
        {
            if ((rc = pick_ancestors(b, rng, b->N, 0, true, need))) return rc;
            if ((rc = gather_into_next(b, b->N, false))) return rc;
            flip(b);
        } else if ((rc = clear_flag(ctx)))
            return rc;
        LAUNCH(ctx, k_reset_states<true>, blocks_for(b->N), kThreads, D, b->state[b->cur], b->N,
               replay_args(ctx, need, 0, true), skip, ctx->d_flag);
        rng->cursor += need;
        if ((rc = check_flag(ctx))) return rc;
        if (weighted) weights_became_uniform(b);
    } else
    {
        if (weighted && ctx->inplace_resample)
        { // survivors keep their slot (and the back buffer stays unallocated); states are redrawn below
            if ((rc = resample_inplace(b, rng, b->N))) return rc;
        } else if (weighted)
        {
            if ((rc = pick_ancestors(b, rng, b->N, 0, false, 0))) return rc;
            if ((rc = gather_into_next(b, b->N, false))) return rc;
            flip(b);
            b->total_weight = 1.0;
            b->suffix_valid = b->cdf_valid = false;
            b->uniform_now  = !b->w_escaped; // the gather wrote 1 / N into every weight
        }
        LAUNCH(ctx, k_reset_states<false>, blocks_for(b->N), kThreads, D, b->state[b->cur], b->N,
               philox_args(rng), 0, ctx->d_flag);
    }
    return FBA_OK;
}

extern "C" int fba_belief_reset_domain_states(fba_belief* b, fba_rng* rng)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    return reset_states(b, rng, true);
}

extern "C" int fba_belief_redraw_domain_states(fba_belief* b, fba_rng* rng)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    return reset_states(b, rng, false);
}

extern "C" int fba_belief_sample(fba_belief* b, fba_rng* rng, int64_t* index)
{
    if (!b || !rng || !index) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (!b->weighted)
    { // FlatFilter::sample (FlatFilter.cpp:97-102): one uniform int, host side
        if (rng->mode == FBA_RNG_REPLAY)
        {
            ReplayRng g(rng->words, rng->cursor, rng->n_words);
            int const i = draw_k(g, (uint32_t)b->N);
            if (g.overrun) return ctx->err = "replay stream underrun", FBA_ERR_RNG_UNDERRUN;
            rng->cursor = g.pos;
            *index      = i;
        } else
        {
            PhiloxRng g(rng->seed, 0, rng->offset++);
            *index = draw_k(g, (uint32_t)b->N);
        }
        return FBA_OK;
    }
    if (rng->mode == FBA_RNG_PHILOX && b->uniform_now)
    { // every weight is 1 / N (init, resample, reset): a weighted draw is floor(u N) — no normalisation pass, no
      // launch, no read-back; the same u and the same rule as fba_runs_sample (k_runs_step<.., 2>), so a run stays
      // bit-identical to its stand-alone belief
        PhiloxRng g(rng->seed, 0, rng->offset++);
        double const u = draw_u(g);
        *index         = std::min<long long>(b->N - 1, (long long)std::floor(u * (double)b->N));
        return FBA_OK;
    }
    if (rng->mode == FBA_RNG_REPLAY)
    {
        if ((rc = stage_words(ctx, rng, 2))) return rc;
        if ((rc = pick_ancestors(b, rng, 1, 2, false, 2))) return rc;
        rng->cursor += 2;
    } else
    {
        if (!b->cdf_valid)
            if ((rc = native_normalize(b, false, 1.0))) return rc;
        LAUNCH(ctx, k_pick_native, 1, kThreads, b->aux, b->N, 1ll, 0, philox_args(rng), b->anc);
    }
    int h = 0;
    CU(ctx, cudaMemcpyAsync(&h, b->anc, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *index = h;
    return FBA_OK;
}

// ---- rejection sampling ----------------------------------------------------------------------

static int ensure_wave(fba_belief* b, long long cap)
{
    fba_ctx* ctx = b->ctx;
    if (cap <= b->wave_cap) return FBA_OK;
    cudaFree(b->att_src), cudaFree(b->att_state), cudaFree(b->att_accept), cudaFree(b->att_pos), cudaFree(b->att_rec);
    cudaFree(b->att_tiles);
    b->att_src = b->att_state = b->att_accept = b->att_pos = b->att_rec = b->att_tiles = nullptr;
    b->wave_cap = 0;
    CU(ctx, cudaMalloc(&b->att_tiles, ((cap + kFlagTile - 1) / kFlagTile + 1) * sizeof(int)));
    CU(ctx, cudaMalloc(&b->att_src, cap * sizeof(int)));
    CU(ctx, cudaMalloc(&b->att_state, cap * sizeof(int)));
    CU(ctx, cudaMalloc(&b->att_accept, cap * sizeof(int)));
    CU(ctx, cudaMalloc(&b->att_pos, cap * sizeof(int)));
    CU(ctx, cudaMalloc(&b->att_rec, cap * b->m->dev.J * sizeof(int)));
    b->wave_cap = cap;
    return FBA_OK;
}

extern "C" int fba_belief_reject_sample(fba_belief* b, int32_t a, int32_t o, fba_rng* rng, int64_t* attempts_out)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, a >= 0 && a < D.A, "action out of range");
    REQUIRE(ctx, o >= 0 && o < D.O, "observation out of range");
    REQUIRE(ctx, b->delta_cap == 0 || D.tabular, "reject_sample: factored beliefs need dense storage");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    // PHILOX + dense storage: accepted sources keep their slot, only repeated acceptances are copied
    // (k_rs_*), no second buffer. REPLAY keeps the reference's attempt order in the other buffer.
    bool const inplace = rng->mode == FBA_RNG_PHILOX && ctx->inplace_resample && b->delta_cap == 0;
    int const nx       = b->cur ^ 1;
    if (!inplace)
    {
        if ((rc = ensure_next(b))) return rc;
    } else if (!b->rs_src)
    {
        size_t const n = (size_t)b->N;
        CU(ctx, cudaMalloc(&b->rs_src, n * sizeof(int)));
        CU(ctx, cudaMalloc(&b->rs_state, n * sizeof(int)));
        CU(ctx, cudaMalloc(&b->rs_rec, n * (size_t)D.J * sizeof(int)));
        CU(ctx, cudaMalloc(&b->rs_flag, n * sizeof(int)));
        CU(ctx, cudaMalloc(&b->rs_pos, n * sizeof(int)));
        CU(ctx, cudaMalloc(&b->rs_tiles, ((n + kFlagTile - 1) / kFlagTile + 1) * sizeof(int)));
    }
    long long accepted = 0, attempts = 0;
    long long const cap = 1ll << 22;
    double rate         = 0.5; // running estimate of the acceptance rate
    unsigned long long const philox_off = (rng->mode == FBA_RNG_PHILOX) ? rng->offset++ : 0;
    std::vector<int> h_accept;
    std::vector<long long> off;
    int empty_waves = 0;

    while (accepted < b->N)
    {
        long long wave = (long long)((double)(b->N - accepted) / rate * 1.05) + 64;
        wave           = std::min(cap, std::max<long long>(1024, wave));
        RngArgs ra{};
        long long n_words = 0;
        if (rng->mode == FBA_RNG_REPLAY)
        {
            // attempt t: one uniform int (1 word, rarely more), then J uniforms (2J words)
            off.assign((size_t)wave + 1, 0);
            long long pos = rng->cursor, n_att = 0;
            for (; n_att < wave; ++n_att)
            {
                ReplayRng g(rng->words, pos, rng->n_words);
                draw_k(g, (uint32_t)b->N);
                long long const end = g.pos + 2ll * D.J;
                if (g.overrun || end > rng->n_words) break; // the stream ends inside this attempt
                off[n_att] = pos - rng->cursor;
                pos        = end;
            }
            if (n_att == 0) return ctx->err = "replay stream underrun in rejection sampling", FBA_ERR_RNG_UNDERRUN;
            wave      = n_att;
            n_words   = pos - rng->cursor;
            off[wave] = n_words;
            if ((rc = ensure_wave(b, wave))) return rc;
            if ((rc = stage_words(ctx, rng, n_words))) return rc;
            if ((rc = stage_offsets(ctx, off))) return rc;
            ra = replay_args(ctx, n_words, 0, true);
        } else
        {
            if ((rc = ensure_wave(b, wave))) return rc;
            ra.seed        = rng->seed;
            ra.offset      = philox_off;
            ra.stream_base = (unsigned long long)attempts;
        }
        if ((rc = clear_flag(ctx))) return rc;
        if (b->delta_cap > 0)
        {
            LAUNCH_DELTA(ctx, k_rs_attempt_delta, rng->mode == FBA_RNG_REPLAY, blocks_for(wave), kThreads, D, b->base,
                         b->lstride, b->counts[b->cur], b->stride, b->state[b->cur], b->sid[b->cur], b->N, a, o,
                         wave, ra, b->att_src, b->att_state, b->att_accept, b->att_rec, ctx->d_flag);
        } else
            LAUNCH_RL(ctx, k_rs_attempt, rng->mode == FBA_RNG_REPLAY, b->m->long_rows, blocks_for(wave), kThreads,
                      D, b->counts[b->cur], b->stride, b->state[b->cur], b->sid[b->cur], b->N, a, o, wave, ra,
                      b->att_src, b->att_state, b->att_accept, b->att_rec, ctx->d_flag);
        {
            int const n_ft = (int)((wave + kFlagTile - 1) / kFlagTile);
            LAUNCH(ctx, k_flag_tile_counts, n_ft, kThreads, b->att_accept, wave, b->att_tiles);
            LAUNCH(ctx, k_flag_scan_tiles, 1, kThreads, b->att_tiles, n_ft, b->d_total);
            LAUNCH(ctx, k_flag_scan_apply, n_ft, kThreads, b->att_accept, wave, b->att_tiles, b->att_pos);
        }
        if (inplace)
            LAUNCH(ctx, k_rs_collect, blocks_for(wave), kThreads, b->N, D.J, wave, b->att_src, b->att_state,
                   b->att_accept, b->att_pos, b->att_rec, accepted, b->rs_src, b->rs_state, b->rs_rec);
        else
            LAUNCH(ctx, k_rs_commit, stream_grid(ctx, wave), kThreads, b->counts[b->cur], b->counts[nx], b->stride,
                   b->sid[b->cur], b->sid[nx], b->state[nx], b->m->d_sizes, b->N, D.J, wave, b->att_src,
                   b->att_state, b->att_accept, b->att_pos, b->att_rec, accepted, b->delta_cap > 0 ? 1 : 0,
                   b->delta_cap, ctx->d_flag);
        int got = 0;
        CU(ctx, cudaMemcpyAsync(&got, b->d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = check_flag(ctx))) return rc; // synchronises

        long long used = wave; // attempts of this wave that count
        if (accepted + got >= b->N)
        { // the N-th acceptance happened inside this wave: find the attempt that produced it
            int const need = (int)(b->N - accepted);
            LAUNCH(ctx, k_find_nth_accept, blocks_for(wave), kThreads, b->att_accept, b->att_pos, wave, need,
                   b->d_total);
            int nth = 0;
            CU(ctx, cudaMemcpyAsync(&nth, b->d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            used     = nth;
            accepted = b->N;
        } else
            accepted += got;
        attempts += used;
        if (rng->mode == FBA_RNG_REPLAY)
        {
            rng->cursor += off[used];
            if (accepted < b->N && rng->cursor >= rng->n_words)
                return ctx->err = "replay stream ended before N particles were accepted", FBA_ERR_RNG_UNDERRUN;
        }
        rate = std::max(1e-4, (double)std::max(got, 1) / (double)wave);
        if (got == 0 && ++empty_waves > 64)
            return ctx->err = "rejection sampling: observation has (near) zero probability under the belief",
                   FBA_ERR_INVALID;
    }
    if (inplace)
    {
        long long const N = b->N;
        int const n_ft    = (int)((N + kFlagTile - 1) / kFlagTile);
        int* const first  = b->noff;   // N ints
        int* const extra  = b->escan;  // N ints
        int* const xpos   = b->src_of; // >= N ints
        LAUNCH(ctx, k_fill_int, blocks_for(N), kThreads, first, N, 0x7fffffff);
        LAUNCH(ctx, k_rs_first, blocks_for(N), kThreads, b->rs_src, N, first);
        LAUNCH(ctx, k_rs_flags, blocks_for(N), kThreads, b->rs_src, first, N, extra, b->rs_flag);
        LAUNCH(ctx, k_flag_tile_counts, n_ft, kThreads, extra, N, b->rs_tiles);
        LAUNCH(ctx, k_flag_scan_tiles, 1, kThreads, b->rs_tiles, n_ft, b->d_total);
        LAUNCH(ctx, k_flag_scan_apply, n_ft, kThreads, extra, N, b->rs_tiles, xpos);
        LAUNCH(ctx, k_flag_tile_counts, n_ft, kThreads, b->rs_flag, N, b->rs_tiles);
        LAUNCH(ctx, k_flag_scan_tiles, 1, kThreads, b->rs_tiles, n_ft, b->d_total);
        LAUNCH(ctx, k_flag_scan_apply, n_ft, kThreads, b->rs_flag, N, b->rs_tiles, b->rs_pos);
        LAUNCH(ctx, k_rs_empty_list, blocks_for(N), kThreads, b->rs_flag, b->rs_pos, N, b->dead);
        LAUNCH(ctx, k_rs_place_extras, stream_grid(ctx, N), kThreads, b->counts[b->cur], b->stride, b->sid[b->cur],
               b->state[b->cur], b->m->d_sizes, N, D.J, b->rs_src, b->rs_state, b->rs_rec, extra, xpos, b->dead);
        LAUNCH(ctx, k_rs_apply_first, blocks_for(N), kThreads, b->counts[b->cur], b->stride, b->state[b->cur], N,
               D.J, first, b->rs_state, b->rs_rec);
    } else
        flip(b);
    if (attempts_out) *attempts_out = attempts;
    return FBA_OK;
}

// ---- reinvigoration ---------------------------------------------------------------------------

namespace {
// host-side view of either random source, for the few sequential draws breeding needs
struct HostDraws
{
    fba_rng* rng;
    ReplayRng replay;
    PhiloxRng philox;
    explicit HostDraws(fba_rng* r) :
            rng(r), replay(r->words, r->cursor, r->n_words), philox(r->seed, 0, r->offset)
    {
        if (r->mode == FBA_RNG_PHILOX) ++r->offset;
    }
    int k(uint32_t range) { return rng->mode == FBA_RNG_REPLAY ? draw_k(replay, range) : draw_k(philox, range); }
    int slow(int max) { return rng->mode == FBA_RNG_REPLAY ? draw_slow_int(replay, max) : draw_slow_int(philox, max); }
    bool overrun() const { return rng->mode == FBA_RNG_REPLAY && replay.overrun; }
    void commit()
    {
        if (rng->mode == FBA_RNG_REPLAY) rng->cursor = replay.pos;
    }
};
} // namespace

// WeightedFilter::replace(i, s, dealloc) (WeightedFilter.cpp:71-90) for the slots in `order`, one after the
// other: the new particle's weight is _total_weight / N and _total_weight moves by the difference. The
// weights are few and the rule is sequential, so it runs on the host over a copy of the weight array.
static int host_replace_weights(fba_belief* b, const std::vector<int>& order)
{
    fba_ctx* ctx = b->ctx;
    if (!b->weighted || order.empty()) return FBA_OK;
    std::vector<double> w((size_t)b->N);
    CU(ctx, cudaMemcpyAsync(w.data(), b->w, (size_t)b->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    volatile double total = b->total_weight; // volatile: separately rounded operations, in order
    for (int slot : order)
    {
        double const w_new = total / (double)b->N;
        volatile double diff = w_new - w[(size_t)slot];
        w[(size_t)slot] = w_new;
        total           = total + diff;
    }
    CU(ctx, cudaMemcpyAsync(b->w, w.data(), (size_t)b->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    b->total_weight = total;
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now = false;
    return FBA_OK;
}

// `amount` x breed (ReinvigoratingRejectionSampling.cpp:24-35): a structure donor from `b`, a counts donor
// from `fc`, the domain's mutate, marginalizeOut; the bred particle goes to slot dst_slots[k] of `dst`, or,
// with dst_slots == NULL, to a uniformly drawn slot of dst = b (FlatFilter::replace).
static int breed_core(fba_belief* dst, const int64_t* dst_slots, fba_belief* b, fba_belief* fc, int64_t amount,
                      int32_t mutate_kind, fba_rng* rng)
{
    fba_ctx* ctx      = b->ctx;
    fba_model* m      = b->m;
    DevModel const& D = m->dev;
    REQUIRE(ctx, fc->m == m && fc->ctx == ctx && dst->m == m && dst->ctx == ctx,
            "reinvigorate: the beliefs must share one model and context");
    REQUIRE(ctx, amount >= 1, "reinvigorate: resample size of < 1 (" + std::to_string(amount) + ")");
    REQUIRE(ctx, !D.tabular && b->delta_cap == 0 && fc->delta_cap == 0 && dst->delta_cap == 0,
            "reinvigorate: factored models (dense storage) only");
    REQUIRE(ctx, !b->weighted && !fc->weighted, "reinvigorate: the donor beliefs are flat filters");
    REQUIRE(ctx, dst_slots || dst == b, "reinvigorate: destination slots missing");
    if (dst_slots)
        for (int64_t k = 0; k < amount; ++k)
            REQUIRE(ctx, dst_slots[k] >= 0 && dst_slots[k] < dst->N, "breed_into: destination slot out of range");
    CU(ctx, cudaSetDevice(ctx->device));

    // host mirrors of the small per-particle arrays the sequential part reads (pinned: one DMA each)
    for (fba_belief* x : {b, fc})
    {
        if (!x->h_sid) CU(ctx, cudaMallocHost(&x->h_sid, (size_t)x->N * sizeof(int)));
        if (!x->h_state) CU(ctx, cudaMallocHost(&x->h_state, (size_t)x->N * sizeof(int)));
    }
    int* const sid    = b->h_sid;
    int* const st     = b->h_state;
    int* const fc_sid = fc->h_sid;
    CU(ctx, cudaMemcpyAsync(sid, b->sid[b->cur], b->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(st, b->state[b->cur], b->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(fc_sid, fc->sid[fc->cur], fc->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));

    int const nT = D.A * D.FS, nO = D.A * D.FO;
    std::vector<uint32_t> tp(nT), op(nO);
    std::map<int, BreedJob> last_writer; // slot -> job; a later breed overwrites an earlier one
    std::vector<int> order;              // destination slots in breeding order (weights of a weighted dst)
    HostDraws g(rng);
    int const first_new_struct = m->n_structs;
    struct UploadGuard // whatever the exit path, structures registered on the host reach the device
    {
        fba_model* m;
        int first;
        bool done = false;
        int run()
        {
            done = true;
            return upload_structures(m, first);
        }
        ~UploadGuard()
        {
            if (!done) upload_structures(m, first);
        }
    } guard{m, first_new_struct};
    for (int64_t k = 0; k < amount; ++k)
    {
        // ReinvigoratingRejectionSampling.cpp:128: breed(fbapomdp, _belief.sample(),
        // _fully_connected_belief.sample()) — g++ -std=c++11 evaluates right to left
        int const counts_donor = g.k((uint32_t)fc->N);
        int const struct_donor = g.k((uint32_t)b->N);
        memcpy(tp.data(), m->t_par.data() + (size_t)sid[struct_donor] * nT, nT * sizeof(uint32_t));
        memcpy(op.data(), m->o_par.data() + (size_t)sid[struct_donor] * nO, nO * sizeof(uint32_t));
        switch (mutate_kind)
        {
            case FBA_MUT_FACTORED_TIGER: // FactoredTigerPriors.cpp:374-375: O[listen = 2][0]
                REQUIRE(ctx, D.A >= 3, "mutate: factored tiger needs 3 actions");
                op[2 * D.FO] ^= 1u << g.slow(D.FS); // Structure::flip_random_edge, BABNModel.cpp:16-31
                break;
            case FBA_MUT_COLLISION_AVOIDANCE: { // CollisionAvoidancePriors.cpp:478-486
                int const a    = g.k((uint32_t)D.A);
                int const obst = 2 + g.k((uint32_t)D.dom_ip[2]);
                tp[a * D.FS + obst] ^= 1u << g.slow(D.FS);
                break;
            }
            case FBA_MUT_SYSADMIN: { // SysAdminFactoredPrior.cpp:51-52: T[action][computer]; the
                                     // subscripts' draws run right to left under g++ -std=c++11
                int const comp = g.k((uint32_t)D.FS);
                int const a    = g.k((uint32_t)D.A);
                tp[a * D.FS + comp] ^= 1u << g.slow(D.FS);
                break;
            }
            case FBA_MUT_GRIDWORLD: { // GridWorldBAPriors.cpp:200-225: toggle the goal feature
                int const a = g.slow(D.A);
                int const f = g.slow(2);
                tp[a * D.FS + f] ^= 1u << (D.FS - 1);
                break;
            }
            default: ctx->err = "reinvigorate: unknown mutate kind"; return FBA_ERR_INVALID;
        }
        int32_t id = -1;
        int rc     = add_structures(m, 1, tp.data(), op.data(), &id, false); // uploaded once, below
        if (rc) return rc;
        if (m->sizes[id] > dst->stride)
        {
            ctx->err = "reinvigorate: mutated structure needs " + std::to_string(m->sizes[id])
                       + " cells, the destination's stride is " + std::to_string(dst->stride);
            return FBA_ERR_CAPACITY;
        }
        int const slot = dst_slots ? (int)dst_slots[k] : g.k((uint32_t)b->N); // FlatFilter::replace, FlatFilter.cpp:39-46
        if (g.overrun()) return ctx->err = "replay stream underrun in reinvigoration", FBA_ERR_RNG_UNDERRUN;
        BreedJob job{counts_donor, fc_sid[counts_donor], id, slot, st[struct_donor]};
        last_writer[slot] = job;
        order.push_back(slot);
        if (dst == b)
        { // a later breed may draw this slot as its structure donor
            sid[slot] = id;
            st[slot]  = job.state;
        }
    }
    g.commit();
    {
        int const rc = guard.run();
        if (rc) return rc;
    }

    std::vector<BreedJob> jobs;
    for (auto const& kv : last_writer) jobs.push_back(kv.second);
    if ((long long)jobs.size() > b->jobs_cap)
    {
        cudaFree(b->d_jobs);
        b->d_jobs   = nullptr;
        b->jobs_cap = 0;
        long long const cap = (long long)jobs.size() * 2 + 64;
        CU(ctx, cudaMalloc(&b->d_jobs, (size_t)cap * sizeof(BreedJob)));
        b->jobs_cap = cap;
    }
    BreedJob* const d_jobs = b->d_jobs;
    CU(ctx, cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(BreedJob), cudaMemcpyHostToDevice, ctx->stream));
    // structures may have been added: the node table pointer is unchanged, its contents were copied
    LAUNCH(ctx, k_breed, (int)jobs.size(), kThreads, D, fc->counts[fc->cur], fc->stride, dst->counts[dst->cur],
           dst->stride, dst->state[dst->cur], dst->sid[dst->cur], d_jobs);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return host_replace_weights(dst, order);
}

extern "C" int fba_belief_reinvigorate(fba_belief* b, fba_belief* fc, int64_t amount, int32_t mutate_kind,
                                       fba_rng* rng)
{
    if (!b || !fc || !rng) return FBA_ERR_INVALID;
    return breed_core(b, nullptr, b, fc, amount, mutate_kind, rng);
}

extern "C" int fba_belief_breed_into(fba_belief* dst, const int64_t* dst_slot, int64_t n, fba_belief* structure_donors,
                                     fba_belief* fully_connected, int32_t mutate_kind, fba_rng* rng)
{
    if (!dst || !dst_slot || !structure_donors || !fully_connected || !rng) return FBA_ERR_INVALID;
    return breed_core(dst, dst_slot, structure_donors, fully_connected, n, mutate_kind, rng);
}

// grow-only device staging of n elements of T
template<class T>
static int stage_buf(fba_ctx* ctx, T** buf, long long* cap, long long n)
{
    if (n <= *cap) return FBA_OK;
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    long long const c = n + n / 2 + 256;
    CU(ctx, cudaMalloc(buf, (size_t)c * sizeof(T)));
    *cap = c;
    return FBA_OK;
}

// ---- rollouts -----------------------------------------------------------------------------------

extern "C" int fba_rollouts(fba_belief* b, int64_t n, const int64_t* particle, const int32_t* start_state,
                            const int32_t* depth, double discount, fba_rng* rng, const int64_t* word_offset,
                            double* returns)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, n >= 0 && particle && start_state && depth && returns, "rollouts: null argument");
    REQUIRE(ctx, discount > 0 && discount <= 1, "rollouts: discount must be in (0, 1]");
    if (n == 0) return FBA_OK;
    for (int64_t i = 0; i < n; ++i)
    {
        REQUIRE(ctx, particle[i] >= 0 && particle[i] < b->N, "rollouts: particle index out of range");
        REQUIRE(ctx, start_state[i] >= 0 && start_state[i] < D.S, "rollouts: start state out of range");
        REQUIRE(ctx, depth[i] >= 0, "rollouts: negative depth");
    }
    CU(ctx, cudaSetDevice(ctx->device));
    {
        int const rcp = stage_buf(ctx, &b->roll_p, &b->roll_cap_p, (long long)n);
        if (rcp) return rcp;
    }
    if (n > b->roll_cap)
    {
        cudaFree(b->roll_s), cudaFree(b->roll_d), cudaFree(b->roll_r);
        b->roll_s = b->roll_d = nullptr, b->roll_r = nullptr;
        b->roll_cap = 0;
        long long const cap = n + n / 2 + 1024;
        CU(ctx, cudaMalloc(&b->roll_s, cap * sizeof(int)));
        CU(ctx, cudaMalloc(&b->roll_d, cap * sizeof(int)));
        CU(ctx, cudaMalloc(&b->roll_r, cap * sizeof(double)));
        b->roll_cap = cap;
    }
    long long* d_p = b->roll_p;
    int *d_s = b->roll_s, *d_d = b->roll_d;
    double* d_r = b->roll_r;
    CU(ctx, cudaMemcpyAsync(d_p, particle, n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_s, start_state, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_d, depth, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    int rc = FBA_OK;
    // warp-per-rollout pays when rows are long and the batch cannot fill the GPU with threads
    int max_range = 1;
    for (int f = 0; f < D.FS; ++f) max_range = std::max(max_range, D.feat_s[f]);
    for (int f = 0; f < D.FO; ++f) max_range = std::max(max_range, D.feat_o[f]);
    // measured (tools/exp_rollouts.py, profiles/r1f_rollouts_*): the cooperative kernel executes 27x more
    // warp instructions and is issue-bound; it is kept selectable but never chosen automatically
    bool const coop = ctx->rollout_coop > 0;
    // thread-per-rollout: spread a small batch over as many SMs as possible (one warp per CTA until
    // every SM has work), so that its uncoalesced row loads do not queue on a few SMs' LSUs
    int tpb = kThreads;
    while (tpb > 32 && (n + tpb - 1) / tpb < 2ll * ctx->sm_count) tpb >>= 1;
    if (rng->mode == FBA_RNG_REPLAY)
    {
        REQUIRE(ctx, word_offset, "rollouts: REPLAY mode needs word_offset");
        long long const avail = rng->n_words - rng->cursor;
        for (int64_t i = 0; i < n; ++i)
            REQUIRE(ctx, word_offset[i] >= 0 && word_offset[i] <= std::max(0ll, avail),
                    "rollouts: word_offset outside the replay stream");
        std::vector<long long> off(word_offset, word_offset + n);
        if ((rc = stage_words(ctx, rng, std::max(0ll, avail)))) return rc;
        if ((rc = stage_offsets(ctx, off))) return rc;
        if ((rc = clear_flag(ctx))) return rc;
        RngArgs const ra = replay_args(ctx, avail, 0, true);
        if (b->delta_cap > 0)
            LAUNCH_DELTA(ctx, k_rollouts_delta, true, blocks_for(n, tpb), tpb, D, b->base, b->lstride,
                         b->counts[b->cur], b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r,
                         ctx->d_flag, ctx->d_counters, (const int*)b->d_proto_sid);
        else if (coop)
            LAUNCH(ctx, (k_rollouts<true, true, false, false>), blocks_for(n * 32), kThreads, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
        else if (b->m->long_rows)
            LAUNCH(ctx, (k_rollouts<true, false, true, false>), blocks_for(n, tpb), tpb, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
        else
            LAUNCH(ctx, (k_rollouts<true, false, false, false>), blocks_for(n, tpb), tpb, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
    } else
    {
        RngArgs const ra = philox_args(rng);
        bool const lr = b->m->long_rows;
        if (b->delta_cap > 0)
            LAUNCH_DELTA(ctx, k_rollouts_delta, false, blocks_for(n, tpb), tpb, D, b->base, b->lstride,
                         b->counts[b->cur], b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r,
                         ctx->d_flag, ctx->d_counters, (const int*)b->d_proto_sid);
        else if (coop && !D.sampled)
            LAUNCH(ctx, (k_rollouts<false, true, false, false>), blocks_for(n * 32), kThreads, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
        else if (D.sampled && lr)
            LAUNCH(ctx, (k_rollouts<false, false, true, true>), blocks_for(n, tpb), tpb, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
        else if (D.sampled)
            LAUNCH(ctx, (k_rollouts<false, false, false, true>), blocks_for(n, tpb), tpb, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
        else if (lr)
            LAUNCH(ctx, (k_rollouts<false, false, true, false>), blocks_for(n, tpb), tpb, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
        else
            LAUNCH(ctx, (k_rollouts<false, false, false, false>), blocks_for(n, tpb), tpb, D, b->counts[b->cur],
                   b->stride, b->sid[b->cur], n, d_p, d_s, d_d, discount, ra, d_r, ctx->d_flag, ctx->d_counters);
    }
    CU(ctx, cudaMemcpyAsync(returns, d_r, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (rng->mode == FBA_RNG_REPLAY)
    {
        if ((rc = check_flag(ctx))) return rc;
        rng->cursor = rng->n_words; // the batch owns the rest of the stream it was handed
    }
    return FBA_OK;
}

// ---- planner support ----------------------------------------------------------------------------

extern "C" int fba_belief_sample_batch(fba_belief* b, fba_rng* rng, int64_t n, int64_t* indices)
{
    if (!b || !rng || !indices || n < 0) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "sample_batch: PHILOX mode only (REPLAY: fba_belief_sample)");
    if (n == 0) return FBA_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = stage_buf(ctx, &b->roll_p, &b->roll_cap_p, n))) return rc;
    bool const by_cdf = b->weighted && !b->uniform_now; // uniform weights: uniform picks, no normalisation pass
    if (by_cdf && !b->cdf_valid)
        if ((rc = native_normalize(b, false, 1.0))) return rc;
    LAUNCH(ctx, k_sample_batch, blocks_for(n), kThreads, by_cdf ? b->aux : (const double*)nullptr, b->N,
           (long long)n, philox_args(rng), b->roll_p);
    CU(ctx, cudaMemcpyAsync(indices, b->roll_p, n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

// domain states of the given particles (host indices in, host states out)
extern "C" int fba_belief_gather_states(fba_belief* b, int64_t n, const int64_t* indices, int32_t* states)
{
    if (!b || n < 0 || (n && (!indices || !states))) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    if (n == 0) return FBA_OK;
    for (int64_t i = 0; i < n; ++i)
        REQUIRE(ctx, indices[i] >= 0 && indices[i] < b->N, "gather_states: particle index out of range");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = stage_buf(ctx, &b->roll_p, &b->roll_cap_p, n))) return rc;
    if ((rc = stage_buf(ctx, &b->step_i, &b->step_cap, n))) return rc;
    CU(ctx, cudaMemcpyAsync(b->roll_p, indices, n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_gather_states, blocks_for(n), kThreads, b->state[b->cur], b->roll_p, (long long)n, b->step_i);
    CU(ctx, cudaMemcpyAsync(states, b->step_i, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

extern "C" int fba_step_batch(fba_belief* b, int64_t n, const int64_t* particle, const int32_t* state,
                              const int32_t* action, fba_rng* rng, int32_t* new_state, int32_t* observation,
                              double* reward, int32_t* terminal)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, n >= 0 && particle && state && action && new_state && observation && reward && terminal,
            "step_batch: null argument");
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "step_batch: PHILOX mode only");
    if (n == 0) return FBA_OK;
    for (int64_t i = 0; i < n; ++i)
    {
        REQUIRE(ctx, particle[i] >= 0 && particle[i] < b->N, "step_batch: particle index out of range");
        REQUIRE(ctx, state[i] >= 0 && state[i] < D.S, "step_batch: state out of range");
        REQUIRE(ctx, action[i] >= 0 && action[i] < D.A, "step_batch: action out of range");
    }
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = stage_buf(ctx, &b->roll_p, &b->roll_cap_p, n))) return rc;
    if ((rc = stage_buf(ctx, &b->step_i, &b->step_cap, 5 * n))) return rc; // state, action, s', o, terminal
    if ((rc = stage_buf(ctx, &b->step_r, &b->step_cap_r, n))) return rc;
    int *d_s = b->step_i, *d_a = b->step_i + n, *d_s2 = b->step_i + 2 * n, *d_o = b->step_i + 3 * n,
        *d_t = b->step_i + 4 * n;
    CU(ctx, cudaMemcpyAsync(b->roll_p, particle, n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_s, state, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_a, action, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    int tpb = kThreads;
    while (tpb > 32 && (n + tpb - 1) / tpb < 2ll * ctx->sm_count) tpb >>= 1;
    if (b->delta_cap > 0)
        LAUNCH_DELTA(ctx, k_step_batch_delta, false, blocks_for(n, tpb), tpb, D, b->base, b->lstride,
                     b->counts[b->cur], b->stride, b->sid[b->cur], (long long)n, b->roll_p, d_s, d_a, philox_args(rng),
                     d_s2, d_o, b->step_r, d_t, ctx->d_flag, (const int*)b->d_proto_sid);
    else
        LAUNCH_RL(ctx, k_step_batch, false, b->m->long_rows, blocks_for(n, tpb), tpb, D, b->counts[b->cur],
                  b->stride, b->sid[b->cur], (long long)n, b->roll_p, d_s, d_a, philox_args(rng), d_s2, d_o,
                  b->step_r, d_t, ctx->d_flag);
    CU(ctx, cudaMemcpyAsync(new_state, d_s2, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(observation, d_o, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(terminal, d_t, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(reward, b->step_r, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

// ---- Bayesian-Dirichlet score ---------------------------------------------------------------------

extern "C" int fba_belief_log_bd_score(fba_belief* b, fba_belief* prior, double* scores)
{
    if (!b || !prior || !scores) return FBA_ERR_INVALID;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, prior->m == b->m && prior->ctx == ctx, "log_bd_score: both beliefs must share one model and context");
    REQUIRE(ctx, !D.tabular && b->delta_cap == 0, "log_bd_score: factored models with dense storage only");
    REQUIRE(ctx, prior->N == b->N || prior->N == 1, "log_bd_score: the prior belief has N particles or one");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = clear_flag(ctx))) return rc;
    LAUNCH(ctx, k_log_bd_score, stream_grid(ctx, b->N), kThreads, D, b->counts[b->cur], b->stride, b->sid[b->cur], b->N,
           prior->counts[prior->cur], prior->stride, prior->sid[prior->cur], prior->N, b->aux, ctx->d_flag);
    b->suffix_valid = b->cdf_valid = false; // aux was used as the result buffer
    b->uniform_now = false;
    CU(ctx, cudaMemcpyAsync(scores, b->aux, (size_t)b->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    REQUIRE(ctx, *ctx->h_flag == 0, "log_bd_score: a particle and its prior have different structures");
    return FBA_OK;
}

// ---- MH structure beliefs: history replay on proposal particles, particle transfer between beliefs ----

extern "C" int fba_belief_replay_history(fba_belief* b, int32_t n_episodes, const int32_t* episode_len,
                                         const int32_t* actions, const int32_t* observations, fba_rng* rng,
                                         int64_t max_attempts)
{
    if (!b || !rng || n_episodes < 0 || (n_episodes && (!episode_len || !actions || !observations))) return FBA_ERR_INVALID;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, b->delta_cap == 0, "replay_history: dense storage only");
    REQUIRE(ctx, max_attempts >= 1, "replay_history: max_attempts must be at least 1");
    long long total = 0;
    int max_len     = 1;
    for (int e = 0; e < n_episodes; ++e)
    {
        REQUIRE(ctx, episode_len[e] >= 0, "replay_history: negative episode length");
        total += episode_len[e];
        max_len = std::max(max_len, (int)episode_len[e]);
    }
    for (long long k = 0; k < total; ++k)
    {
        REQUIRE(ctx, actions[k] >= 0 && actions[k] < D.A, "replay_history: action out of range");
        REQUIRE(ctx, observations[k] >= 0 && observations[k] < D.O, "replay_history: observation out of range");
    }
    CU(ctx, cudaSetDevice(ctx->device));
    DevTmp<int> d_len, d_act, d_obs, d_rec, d_failed;
    DevTmp<long long> d_words_used;
    CU(ctx, cudaMalloc(&d_len, std::max(1, (int)n_episodes) * sizeof(int)));
    CU(ctx, cudaMalloc(&d_act, std::max(1ll, total) * sizeof(int)));
    CU(ctx, cudaMalloc(&d_obs, std::max(1ll, total) * sizeof(int)));
    CU(ctx, cudaMalloc(&d_rec, (size_t)b->N * max_len * D.J * sizeof(int)));
    CU(ctx, cudaMalloc(&d_failed, sizeof(int)));
    CU(ctx, cudaMemsetAsync(d_failed, 0, sizeof(int), ctx->stream));
    if (n_episodes)
        CU(ctx, cudaMemcpyAsync(d_len, episode_len, n_episodes * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (total)
    {
        CU(ctx, cudaMemcpyAsync(d_act, actions, total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(d_obs, observations, total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    HistoryArgs H{};
    H.n_episodes = n_episodes, H.max_len = max_len;
    H.episode_len = d_len, H.actions = d_act, H.observations = d_obs;
    H.max_attempts = max_attempts;
    int rc;
    bool const lr = b->m->long_rows;
    if (rng->mode == FBA_RNG_REPLAY)
    { // every particle draws from its own equal slice of the remaining words
        long long const need = rng->n_words - rng->cursor, per = need / b->N;
        REQUIRE(ctx, per >= 1, "replay_history: the replay stream is shorter than one word per particle");
        if ((rc = stage_words(ctx, rng, per * b->N))) return rc;
        if ((rc = clear_flag(ctx))) return rc;
        CU(ctx, cudaMalloc(&d_words_used, (size_t)b->N * sizeof(long long)));
        if (lr)
            LAUNCH(ctx, (k_mh_replay<true, true>), blocks_for(b->N), kThreads, D, b->counts[b->cur], b->stride,
                   b->state[b->cur], b->sid[b->cur], b->N, H, replay_args(ctx, per * b->N, per, false), (int*)d_rec,
                   (int*)d_failed, ctx->d_flag, (long long*)d_words_used);
        else
            LAUNCH(ctx, (k_mh_replay<true, false>), blocks_for(b->N), kThreads, D, b->counts[b->cur], b->stride,
                   b->state[b->cur], b->sid[b->cur], b->N, H, replay_args(ctx, per * b->N, per, false), (int*)d_rec,
                   (int*)d_failed, ctx->d_flag, (long long*)d_words_used);
        if (b->N == 1)
        { // a single proposal consumes the stream the way the reference does: exactly the words it drew
            long long used = 0;
            CU(ctx, cudaMemcpyAsync(&used, (long long*)d_words_used, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            rng->cursor += used;
        } else
            rng->cursor += per * b->N;
    } else
    {
        if (lr)
            LAUNCH(ctx, (k_mh_replay<false, true>), blocks_for(b->N), kThreads, D, b->counts[b->cur], b->stride,
                   b->state[b->cur], b->sid[b->cur], b->N, H, philox_args(rng), (int*)d_rec, (int*)d_failed,
                   ctx->d_flag, (long long*)nullptr);
        else
            LAUNCH(ctx, (k_mh_replay<false, false>), blocks_for(b->N), kThreads, D, b->counts[b->cur], b->stride,
                   b->state[b->cur], b->sid[b->cur], b->N, H, philox_args(rng), (int*)d_rec, (int*)d_failed,
                   ctx->d_flag, (long long*)nullptr);
    }
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, d_failed, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag)
    {
        ctx->err = "replay_history: a particle needed more than max_attempts episode attempts (its model gives the "
                   "observed history almost no probability)";
        return FBA_ERR_CAPACITY;
    }
    if (rng->mode == FBA_RNG_REPLAY && (rc = check_flag(ctx))) return rc;
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now = false;
    return FBA_OK;
}

extern "C" int fba_belief_assign_from(fba_belief* dst, int64_t first, fba_belief* src, int64_t n,
                                      const int64_t* src_index)
{
    if (!dst || !src || n < 0 || (n && !src_index)) return FBA_ERR_INVALID;
    fba_ctx* ctx = dst->ctx;
    REQUIRE(ctx, src->ctx == ctx && src->m == dst->m && src->stride == dst->stride && dst->delta_cap == 0
                     && src->delta_cap == 0,
            "assign_from: both beliefs must share context, model and stride (dense storage)");
    REQUIRE(ctx, first >= 0 && first + n <= dst->N, "assign_from: destination range out of bounds");
    REQUIRE(ctx, n <= dst->anc_cap, "assign_from: too many particles for one call");
    if (n == 0) return FBA_OK;
    std::vector<int> idx((size_t)n);
    for (int64_t j = 0; j < n; ++j)
    {
        REQUIRE(ctx, src_index[j] >= 0 && src_index[j] < src->N, "assign_from: source index out of range");
        idx[(size_t)j] = (int)src_index[j];
    }
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(dst->anc, idx.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_gather, stream_grid(ctx, n), kThreads, src->counts[src->cur], dst->counts[dst->cur] + first * dst->stride,
           dst->stride, src->state[src->cur], dst->state[dst->cur] + first, src->sid[src->cur], dst->sid[dst->cur] + first,
           dst->m->d_sizes, (double*)nullptr, 0.0, dst->anc, (long long)n, 0);
    CU(ctx, cudaStreamSynchronize(ctx->stream)); // idx is a host temporary
    return FBA_OK;
}

// ---- MHwithinGibbs: state histories and their posterior counts -------------------------------------------

namespace {
// the (action, observation) history on the device + its checks, shared by the calls below
struct DeviceHistory
{
    PoolTmp<int> len, act, obs;
    HistoryArgs H{};
    long long total = 0;
    int stage(fba_ctx* ctx, DevModel const& D, int32_t n_episodes, const int32_t* episode_len, const int32_t* actions,
              const int32_t* observations, bool need_steps)
    {
        int max_len = 1;
        for (int e = 0; e < n_episodes; ++e)
        {
            REQUIRE(ctx, episode_len[e] >= (need_steps ? 1 : 0), "history: an episode without steps");
            total += episode_len[e];
            max_len = std::max(max_len, (int)episode_len[e]);
        }
        for (long long k = 0; k < total; ++k)
        {
            REQUIRE(ctx, actions[k] >= 0 && actions[k] < D.A, "history: action out of range");
            REQUIRE(ctx, observations[k] >= 0 && observations[k] < D.O, "history: observation out of range");
        }
        int rc;
        if ((rc = len.alloc(ctx, (size_t)n_episodes)) || (rc = act.alloc(ctx, (size_t)total)) || (rc = obs.alloc(ctx, (size_t)total)))
            return rc;
        if (n_episodes)
            CU(ctx, cudaMemcpyAsync(len, episode_len, n_episodes * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (total)
        {
            CU(ctx, cudaMemcpyAsync(act, actions, total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, cudaMemcpyAsync(obs, observations, total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        }
        H.n_episodes = n_episodes, H.max_len = max_len;
        H.episode_len = len, H.actions = act, H.observations = obs;
        H.max_attempts = 1;
        return FBA_OK;
    }
};
} // namespace

extern "C" int fba_belief_sample_state_history(fba_belief* b, int32_t method, int32_t n_episodes,
                                               const int32_t* episode_len, const int32_t* actions,
                                               const int32_t* observations, const float* state_prior, fba_rng* rng,
                                               int64_t max_attempts, int32_t* states)
{
    if (!b || !rng || !states || n_episodes < 1 || !episode_len || !actions || !observations) return FBA_ERR_INVALID;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, b->delta_cap == 0, "sample_state_history: dense storage only");
    REQUIRE(ctx, method == 0 || method == 1, "sample_state_history: method 0 (messages) or 1 (rejection sampling)");
    REQUIRE(ctx, !D.sampled, "sample_state_history: expected-Dirichlet models only (MHwithinGibbs.cpp:223-225)");
    REQUIRE(ctx, method == 1 || state_prior, "sample_state_history: the message method needs the domain state prior");
    REQUIRE(ctx, max_attempts >= 1, "sample_state_history: max_attempts must be at least 1");
    CU(ctx, cudaSetDevice(ctx->device));
    DeviceHistory h;
    int rc;
    if ((rc = h.stage(ctx, D, n_episodes, episode_len, actions, observations, true))) return rc;
    h.H.max_attempts         = max_attempts;
    long long const out_len  = h.total + n_episodes;
    PoolTmp<int> d_out, d_failed;
    PoolTmp<long long> d_words_used; // REPLAY: words every particle's slice actually handed out
    if ((rc = d_out.alloc(ctx, (size_t)b->N * out_len)) || (rc = d_failed.alloc(ctx, 1))
        || (rc = d_words_used.alloc(ctx, (size_t)b->N)))
        return rc;
    CU(ctx, cudaMemsetAsync(d_failed, 0, sizeof(int), ctx->stream));
    bool const replay   = rng->mode == FBA_RNG_REPLAY;
    long long const per = replay ? (rng->n_words - rng->cursor) / b->N : 0;
    RngArgs ra;
    if (replay)
    { // every particle draws from its own equal slice of the remaining words
        REQUIRE(ctx, per >= 1, "sample_state_history: the replay stream is shorter than one word per particle");
        if ((rc = stage_words(ctx, rng, per * b->N))) return rc;
        ra = replay_args(ctx, per * b->N, per, false);
    } else
        ra = philox_args(rng);
    if ((rc = clear_flag(ctx))) return rc;
    if (method == 1)
    {
        bool const lr = b->m->long_rows;
        if (replay && lr)
            LAUNCH(ctx, (k_state_history_rs<true, true>), blocks_for(b->N, 32), 32, D, b->counts[b->cur], b->stride,
                   b->sid[b->cur], b->N, h.H, ra, (int*)d_out, out_len, (int*)d_failed, ctx->d_flag,
                   (long long*)d_words_used);
        else if (replay)
            LAUNCH(ctx, (k_state_history_rs<true, false>), blocks_for(b->N, 32), 32, D, b->counts[b->cur], b->stride,
                   b->sid[b->cur], b->N, h.H, ra, (int*)d_out, out_len, (int*)d_failed, ctx->d_flag,
                   (long long*)d_words_used);
        else
        { // PHILOX: one warp per particle, 32 attempts per round (the thread-per-particle kernel is the REPLAY form)
            PoolTmp<int> d_scratch;
            if ((rc = d_scratch.alloc(ctx, (size_t)b->N * 32 * (h.H.max_len + 1)))) return rc;
            if (lr)
                LAUNCH(ctx, (k_state_history_rs_warp<true>), blocks_for(b->N * 32, 64), 64, D, b->counts[b->cur], b->stride,
                       b->sid[b->cur], b->N, h.H, ra, (int*)d_out, out_len, (int*)d_scratch, (int*)d_failed);
            else
                LAUNCH(ctx, (k_state_history_rs_warp<false>), blocks_for(b->N * 32, 64), 64, D, b->counts[b->cur], b->stride,
                       b->sid[b->cur], b->N, h.H, ra, (int*)d_out, out_len, (int*)d_scratch, (int*)d_failed);
        }
    } else
    {
        {
            int cells = 0;
            for (int f = 0; f < D.FS; ++f) cells += D.feat_s[f];
            for (int f = 0; f < D.FO; ++f) cells += D.feat_o[f];
            REQUIRE(ctx, cells <= kFlattenCells, "sample_state_history: the feature ranges sum to more than 192");
        }
        // flattenT / flattenO of every particle: A S S + A O S floats each, and (max_len + 2) S doubles of messages
        double const bytes = (double)b->N * ((double)D.A * D.S * ((double)D.S + D.O) * 4.0 + (h.H.max_len + 2.0) * D.S * 8.0);
        REQUIRE(ctx, bytes < 64e9, "sample_state_history: the flattened models of these particles do not fit (N A S (S + O) floats)");
        std::vector<unsigned char> used((size_t)D.A, 0);
        for (long long k = 0; k < h.total; ++k) used[(size_t)actions[k]] = 1;
        PoolTmp<unsigned char> d_used;
        PoolTmp<float> d_prior;
        if ((rc = d_used.alloc(ctx, (size_t)D.A)) || (rc = d_prior.alloc(ctx, (size_t)D.S))) return rc;
        // the big pieces live in one grow-only buffer of the context: gigabytes for large S, and allocating them
        // per call cost several times the kernels
        size_t const nT = (size_t)b->N * D.A * D.S * D.S, nO = (size_t)b->N * D.A * D.O * D.S,
                     nM = (size_t)b->N * (h.H.max_len + 2) * D.S;
        size_t const bytes_needed = nM * sizeof(double) + (nT + nO) * sizeof(float);
        if (bytes_needed > ctx->big_scratch_bytes)
        {
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->big_scratch);
            ctx->big_scratch = nullptr, ctx->big_scratch_bytes = 0;
            CU(ctx, cudaMalloc(&ctx->big_scratch, bytes_needed));
            ctx->big_scratch_bytes = bytes_needed;
        }
        double* const d_msg = (double*)ctx->big_scratch;
        float* const d_T    = (float*)(d_msg + nM);
        float* const d_O    = d_T + nT;
        CU(ctx, cudaMemcpyAsync(d_used, used.data(), used.size(), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(d_prior, state_prior, (size_t)D.S * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        dim3 const grid((unsigned)std::min<long long>(blocks_for((long long)D.A * D.S), 4 * ctx->sm_count), (unsigned)b->N);
        int const msg_threads = (int)std::min<long long>(kMsgThreads, std::max<long long>(64, ((long long)D.S + 31) / 32 * 32));
        LAUNCH(ctx, k_flatten_model, grid, kThreads, D, b->counts[b->cur], b->stride, b->sid[b->cur],
               (const unsigned char*)d_used, (float*)d_T, (float*)d_O);
        bool const sh          = (size_t)D.S * 2 * sizeof(double) <= 48 * 1024; // two message rows in shared memory
        size_t const sh_bytes  = sh ? (size_t)D.S * 2 * sizeof(double) : 0;
#define LAUNCH_MSG(R, SHM)                                                                                         \
    do {                                                                                                           \
        if (ctx->profiling) profile_mark(ctx, "k_state_history_msg", true);                                        \
        k_state_history_msg<R, SHM><<<(int)b->N, msg_threads, sh_bytes, ctx->stream>>>(                            \
            D, b->N, h.H, (const float*)d_T, (const float*)d_O, (const float*)d_prior, (double*)d_msg, ra,         \
            (int*)d_out, out_len, ctx->d_flag, (long long*)d_words_used);                                          \
        if (ctx->profiling) profile_mark(ctx, "k_state_history_msg", false);                                       \
        ++ctx->launches;                                                                                           \
        CU(ctx, cudaGetLastError());                                                                               \
    } while (0)
        // one model over a cluster of CTAs when its states split evenly and the rows fit shared memory
        bool const clustered = sh && ctx->msg_cluster && D.S % kMsgCluster == 0 && D.S / kMsgCluster >= 32;
        if (clustered)
        {
            int const cl_threads = (int)std::min<long long>(kMsgThreads, ((long long)D.S / kMsgCluster + 31) / 32 * 32);
            if (ctx->profiling) profile_mark(ctx, "k_state_history_msg_cluster", true);
            if (replay)
                k_state_history_msg_cluster<true><<<(int)b->N * kMsgCluster, cl_threads, sh_bytes, ctx->stream>>>(
                    D, b->N, h.H, (const float*)d_T, (const float*)d_O, (const float*)d_prior, (double*)d_msg, ra,
                    (int*)d_out, out_len, ctx->d_flag, (long long*)d_words_used);
            else
                k_state_history_msg_cluster<false><<<(int)b->N * kMsgCluster, cl_threads, sh_bytes, ctx->stream>>>(
                    D, b->N, h.H, (const float*)d_T, (const float*)d_O, (const float*)d_prior, (double*)d_msg, ra,
                    (int*)d_out, out_len, ctx->d_flag, (long long*)nullptr);
            if (ctx->profiling) profile_mark(ctx, "k_state_history_msg_cluster", false);
            ++ctx->launches;
            CU(ctx, cudaGetLastError());
        } else if (replay && sh)
            LAUNCH_MSG(true, true);
        else if (replay)
            LAUNCH_MSG(true, false);
        else if (sh)
            LAUNCH_MSG(false, true);
        else
            LAUNCH_MSG(false, false);
#undef LAUNCH_MSG
    }
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, d_failed, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(states, d_out, (size_t)b->N * out_len * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag)
    {
        ctx->err = "sample_state_history: a particle needed more than max_attempts episode attempts (or ran out of "
                   "replay words): its model gives the observed history almost no probability";
        return replay ? FBA_ERR_RNG_UNDERRUN : FBA_ERR_CAPACITY;
    }
    if (replay)
    {
        if ((rc = check_flag(ctx))) return rc;
        if (b->N == 1)
        { // a single model consumes the stream the way the reference does: exactly the words it drew
            long long used = 0;
            CU(ctx, cudaMemcpyAsync(&used, (long long*)d_words_used, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            rng->cursor += used;
        } else
            rng->cursor += per * b->N;
    }
    return FBA_OK;
}

extern "C" int fba_belief_add_history_counts(fba_belief* b, int32_t n_episodes, const int32_t* episode_len,
                                             const int32_t* actions, const int32_t* observations,
                                             const int32_t* states, int32_t shared)
{
    if (!b || !states || n_episodes < 1 || !episode_len || !actions || !observations) return FBA_ERR_INVALID;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, b->delta_cap == 0, "add_history_counts: dense storage only");
    CU(ctx, cudaSetDevice(ctx->device));
    DeviceHistory h;
    int rc;
    if ((rc = h.stage(ctx, D, n_episodes, episode_len, actions, observations, false))) return rc;
    if (h.total == 0) return FBA_OK;
    long long const seq_len = h.total + n_episodes, n_seq = shared ? 1 : b->N;
    for (long long k = 0; k < n_seq * seq_len; ++k)
        REQUIRE(ctx, states[k] >= 0 && states[k] < D.S, "add_history_counts: domain state out of range");
    std::vector<int> pos((size_t)h.total);
    { // step t of the flattened history sits at states[pos[t]], its successor right after
        long long t = 0, p = 0;
        for (int e = 0; e < n_episodes; ++e)
        {
            for (int k = 0; k < episode_len[e]; ++k) pos[(size_t)t++] = (int)(p + k);
            p += episode_len[e] + 1;
        }
    }
    PoolTmp<int> d_pos, d_states;
    if ((rc = d_pos.alloc(ctx, (size_t)h.total)) || (rc = d_states.alloc(ctx, (size_t)n_seq * seq_len))) return rc;
    CU(ctx, cudaMemcpyAsync(d_pos, pos.data(), (size_t)h.total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_states, states, (size_t)n_seq * seq_len * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_add_history_counts, blocks_for(b->N * h.total), kThreads, D, b->counts[b->cur], b->stride, b->sid[b->cur],
           b->N, (int)h.total, (const int*)h.act, (const int*)h.obs, (const int*)d_pos, (const int*)d_states,
           shared ? 0ll : seq_len);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

// ---- single particles between the filters of the composite beliefs -------------------------------------

static int replace_core(fba_belief* dst, const std::vector<int>& dst_index, fba_belief* src,
                        const std::vector<int>& src_index)
{
    fba_ctx* ctx = dst->ctx;
    size_t const n = dst_index.size();
    if (n == 0) return FBA_OK;
    std::map<int, int> last; // destination slot -> source of its last writer
    for (size_t j = 0; j < n; ++j) last[dst_index[j]] = src_index[j];
    std::vector<int2> jobs;
    jobs.reserve(last.size());
    for (auto const& kv : last) jobs.push_back(make_int2(kv.second, kv.first));
    PoolTmp<int2> d_jobs;
    CU(ctx, cudaSetDevice(ctx->device));
    {
        int const rc = d_jobs.alloc(ctx, jobs.size());
        if (rc) return rc;
    }
    CU(ctx, cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_replace_from, stream_grid(ctx, (long long)jobs.size()), kThreads, src->counts[src->cur],
           dst->counts[dst->cur], dst->stride, src->state[src->cur], dst->state[dst->cur], src->sid[src->cur],
           dst->sid[dst->cur], dst->m->d_sizes, (const int2*)d_jobs.p, (long long)jobs.size());
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return host_replace_weights(dst, dst_index);
}

extern "C" int fba_belief_replace_from(fba_belief* dst, const int64_t* dst_index, fba_belief* src,
                                       const int64_t* src_index, int64_t n)
{
    if (!dst || !src || n < 0 || (n && (!dst_index || !src_index))) return FBA_ERR_INVALID;
    fba_ctx* ctx = dst->ctx;
    REQUIRE(ctx, src != dst && src->ctx == ctx && src->m == dst->m && src->stride == dst->stride
                     && dst->delta_cap == 0 && src->delta_cap == 0,
            "replace_from: two different beliefs of one context, model and stride (dense storage)");
    std::vector<int> di((size_t)n), si((size_t)n);
    for (int64_t j = 0; j < n; ++j)
    {
        REQUIRE(ctx, src_index[j] >= 0 && src_index[j] < src->N, "replace_from: source index out of range");
        REQUIRE(ctx, dst_index[j] >= 0 && dst_index[j] < dst->N, "replace_from: destination index out of range");
        di[(size_t)j] = (int)dst_index[j], si[(size_t)j] = (int)src_index[j];
    }
    return replace_core(dst, di, src, si);
}

// CheatingReinvigoration::cheat (prototypes/CheatingReinvigoration.cpp:136-147)
extern "C" int fba_belief_cheat(fba_belief* belief, fba_belief* correct, int64_t amount, fba_rng* rng)
{
    if (!belief || !correct || !rng || amount < 0) return FBA_ERR_INVALID;
    fba_ctx* ctx = belief->ctx;
    REQUIRE(ctx, belief != correct && correct->ctx == ctx && correct->m == belief->m
                     && correct->stride == belief->stride && belief->delta_cap == 0 && correct->delta_cap == 0,
            "cheat: two different beliefs of one context, model and stride (dense storage)");
    REQUIRE(ctx, belief->weighted && !correct->weighted, "cheat: a weighted belief and a flat correct-structure filter");
    std::vector<int> di, si;
    HostDraws g(rng);
    for (int64_t k = 0; k < amount; ++k)
    { // _belief.replace(rnd::slowRandomInt(0, size), copyState(_correct_structured_belief.sample()), ...):
      // g++ evaluates the arguments right to left
        si.push_back(g.k((uint32_t)correct->N));
        di.push_back(g.slow((int)belief->N));
        if (g.overrun()) return ctx->err = "replay stream underrun in cheat", FBA_ERR_RNG_UNDERRUN;
    }
    g.commit();
    return replace_core(belief, di, correct, si);
}

// WeightedFilter::leastLikely (src/beliefs/particle_filters/WeightedFilter.cpp:206-243), with the same
// container (std::priority_queue over (weight, index) pairs compared by weight) driven the same way, so
// that ties come out in the reference's order: seeded with the first n particles, then EVERY particle
// (the first n again) replaces the current largest if its weight is strictly smaller.
extern "C" int fba_belief_least_likely(fba_belief* b, int64_t n, int64_t* index)
{
    if (!b || !index || n < 0) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, b->weighted, "least_likely: weighted beliefs only");
    REQUIRE(ctx, n < b->N, "least_likely: n must be smaller than the belief");
    if (n == 0) return FBA_OK;
    std::vector<double> w((size_t)b->N);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(w.data(), b->w, (size_t)b->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    typedef std::pair<double, int> El;
    struct Less
    {
        bool operator()(El l, El r) const { return l.first < r.first; }
    };
    std::priority_queue<El, std::vector<El>, Less> q;
    for (int64_t i = 0; i < n; ++i) q.push({w[(size_t)i], (int)i});
    for (int64_t i = 0; i < b->N; ++i)
        if (w[(size_t)i] < q.top().first)
        {
            q.pop();
            q.push({w[(size_t)i], (int)i});
        }
    for (int64_t i = 0; i < n; ++i)
    {
        index[i] = q.top().second;
        q.pop();
    }
    return FBA_OK;
}

// StructureIncubatorSampling::reinvigorateBelief (factored/StructureIncubatorSampling.cpp:155-187): every
// shadow particle whose normalised weight exceeds the threshold is copied over a uniformly drawn particle
// of the flat belief (FlatFilter::replace, one draw each, in particle order) and its weight set to zero;
// if any moved, the shadow weights are normalised (WeightedFilter::normalize, sequential sums).
extern "C" int fba_belief_promote(fba_belief* shadow, fba_belief* belief, double threshold, fba_rng* rng,
                                  int64_t* n_promoted)
{
    if (!shadow || !belief || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = shadow->ctx;
    REQUIRE(ctx, belief != shadow && belief->ctx == ctx && belief->m == shadow->m && belief->stride == shadow->stride
                     && belief->delta_cap == 0 && shadow->delta_cap == 0,
            "promote: two different beliefs of one context, model and stride (dense storage)");
    REQUIRE(ctx, shadow->weighted && !belief->weighted, "promote: a weighted shadow belief and a flat belief");
    std::vector<double> w((size_t)shadow->N);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(w.data(), shadow->w, (size_t)shadow->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int> di, si;
    HostDraws g(rng);
    for (int64_t i = 0; i < shadow->N; ++i)
        if (w[(size_t)i] / shadow->total_weight > threshold) // WeightedFilter::normalizedWeight
        {
            si.push_back((int)i);
            di.push_back(g.k((uint32_t)belief->N));
            if (g.overrun()) return ctx->err = "replay stream underrun in promote", FBA_ERR_RNG_UNDERRUN;
            w[(size_t)i] = 0.0;
        }
    g.commit();
    if (n_promoted) *n_promoted = (int64_t)si.size();
    if (si.empty()) return FBA_OK;
    int rc = replace_core(belief, di, shadow, si);
    if (rc) return rc;
    volatile double total = 0.0, acc = 0.0; // normalize(): total, then divide and re-accumulate, in order
    for (double x : w) total = total + x;
    REQUIRE(ctx, total > 0.0, "promote: every shadow particle passed the threshold (the reference divides by zero here; "
                              "choose a threshold above 1 / size)");
    for (double& x : w)
    {
        x   = x / total;
        acc = acc + x;
    }
    CU(ctx, cudaMemcpyAsync(shadow->w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    shadow->total_weight = acc;
    shadow->suffix_valid = shadow->cdf_valid = false;
    shadow->uniform_now = false;
    return FBA_OK;
}

// ---- NestedBelief ----------------------------------------------------------------------------------------

struct fba_nested
{
    fba_ctx* ctx    = nullptr;
    fba_model* m    = nullptr;
    fba_belief* top = nullptr; // the weighted top filter: count blocks, structure ids, weights
    long long n_top = 0;
    int n_bottom    = 0;
    int* states[2]  = {nullptr, nullptr}; // [n_top][n_bottom] domain states, double-buffered per update
    int cur         = 0;
    long long* d_attempts = nullptr;
    int* d_failed   = nullptr;
};

extern "C" void fba_nested_destroy(fba_nested* n)
{
    if (!n) return;
    cudaSetDevice(n->ctx->device);
    fba_belief_destroy(n->top);
    cudaFree(n->states[0]), cudaFree(n->states[1]), cudaFree(n->d_attempts), cudaFree(n->d_failed);
    delete n;
}

extern "C" int fba_nested_create(fba_ctx* ctx, fba_model* m, int64_t n_top, int64_t n_bottom, int64_t stride,
                                 fba_nested** out)
{
    if (!ctx || !m || !out) return FBA_ERR_INVALID;
    *out = nullptr;
    // NestedBelief.cpp:19-26
    REQUIRE(ctx, n_top >= 1 && n_bottom >= 1,
            "NestedBelief: cannot initiate with filter size < 1 (top: " + std::to_string(n_top) + ", bottom: "
                + std::to_string(n_bottom) + ")");
    REQUIRE(ctx, n_bottom < (1ll << 31) && n_top * n_bottom < (1ll << 40), "NestedBelief: filters too large");
    REQUIRE(ctx, m->delta_cap == 0, "NestedBelief: dense storage only");
    auto n      = new fba_nested();
    n->ctx      = ctx;
    n->m        = m;
    n->n_top    = n_top;
    n->n_bottom = (int)n_bottom;
    int rc      = fba_belief_create(ctx, m, n_top, stride, 1, &n->top);
    if (rc)
    {
        delete n;
        return rc;
    }
    cudaError_t e = cudaMalloc(&n->states[0], (size_t)n_top * n_bottom * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&n->states[1], (size_t)n_top * n_bottom * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&n->d_attempts, (size_t)n_top * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&n->d_failed, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(n->states[0], 0, (size_t)n_top * n_bottom * sizeof(int));
    if (e != cudaSuccess)
    {
        ctx->err = std::string("nested alloc: ") + cudaGetErrorString(e);
        fba_nested_destroy(n);
        return FBA_ERR_CUDA;
    }
    *out = n;
    return FBA_OK;
}

extern "C" fba_belief* fba_nested_top(fba_nested* n)
{
    return n ? n->top : nullptr;
}
extern "C" int64_t fba_nested_bottom_size(const fba_nested* n)
{
    return n ? n->n_bottom : 0;
}

extern "C" int fba_nested_upload_states(fba_nested* n, int64_t first_top, int64_t count, const int32_t* states)
{
    if (!n || !states) return FBA_ERR_INVALID;
    fba_ctx* ctx = n->ctx;
    REQUIRE(ctx, first_top >= 0 && count >= 0 && first_top + count <= n->n_top, "nested upload: range out of bounds");
    for (int64_t k = 0; k < count * n->n_bottom; ++k)
        REQUIRE(ctx, states[k] >= 0 && states[k] < n->m->dev.S, "nested upload: domain state out of range");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(n->states[n->cur] + first_top * n->n_bottom, states, (size_t)count * n->n_bottom * sizeof(int),
                            cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

extern "C" int fba_nested_download_states(fba_nested* n, int64_t first_top, int64_t count, int32_t* states)
{
    if (!n || !states) return FBA_ERR_INVALID;
    fba_ctx* ctx = n->ctx;
    REQUIRE(ctx, first_top >= 0 && count >= 0 && first_top + count <= n->n_top, "nested download: range out of bounds");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(states, n->states[n->cur] + first_top * n->n_bottom, (size_t)count * n->n_bottom * sizeof(int),
                            cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

// NestedBelief::resetDomainStateDistribution (NestedBelief.cpp:33-61): n_bottom fresh domain start states
// for every top particle, in particle order — the flat reset kernel over all n_top * n_bottom states
extern "C" int fba_nested_reset_domain_states(fba_nested* n, fba_rng* rng)
{
    if (!n || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx      = n->ctx;
    DevModel const& D = n->m->dev;
    long long const total = n->n_top * n->n_bottom;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (rng->mode == FBA_RNG_REPLAY)
    {
        std::vector<long long> off((size_t)total);
        long long pos = rng->cursor;
        for (long long j = 0; j < total; ++j)
        {
            off[(size_t)j] = pos - rng->cursor;
            pos += start_state_words(D, rng, pos);
        }
        long long const need = pos - rng->cursor;
        if ((rc = stage_words(ctx, rng, need))) return rc;
        if ((rc = stage_offsets(ctx, off))) return rc;
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if ((rc = clear_flag(ctx))) return rc;
        LAUNCH(ctx, k_reset_states<true>, blocks_for(total), kThreads, D, n->states[n->cur], total,
               replay_args(ctx, need, 0, true), 0, ctx->d_flag);
        rng->cursor += need;
        if ((rc = check_flag(ctx))) return rc;
    } else
        LAUNCH(ctx, k_reset_states<false>, blocks_for(total), kThreads, D, n->states[n->cur], total, philox_args(rng),
               0, ctx->d_flag);
    return FBA_OK;
}

// WeightedFilter::normalize() on the top filter (WeightedFilter.cpp:118-143), sequential sums on the host
static int nested_normalize(fba_nested* n)
{
    fba_ctx* ctx  = n->ctx;
    fba_belief* b = n->top;
    std::vector<double> w((size_t)b->N);
    CU(ctx, cudaMemcpyAsync(w.data(), b->w, w.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    volatile double total = 0.0, acc = 0.0;
    for (double x : w) total = total + x;
    for (double& x : w)
    {
        x   = x / total;
        acc = acc + x;
    }
    CU(ctx, cudaMemcpyAsync(b->w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    b->total_weight = acc;
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now = false;
    return FBA_OK;
}

#define LAUNCH_NESTED(R, L, S)                                                                                     \
    LAUNCH(ctx, (k_nested_update<R, L, S>), blocks_for(n->n_top, 32), 32, D, b->counts[b->cur], b->stride,         \
           b->sid[b->cur], b->w, n->n_top, n->n_bottom, n->states[n->cur], n->states[n->cur ^ 1], (int)action,     \
           (int)observation, amount, (long long)max_attempts, ra, n->d_attempts, n->d_failed, ctx->d_flag)

extern "C" int fba_nested_update(fba_nested* n, int32_t action, int32_t observation, fba_rng* rng, int64_t max_attempts,
                                 int64_t* attempts)
{
    if (!n || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx      = n->ctx;
    fba_belief* b     = n->top;
    DevModel const& D = n->m->dev;
    REQUIRE(ctx, action >= 0 && action < D.A && observation >= 0 && observation < D.O, "nested update: action / observation out of range");
    REQUIRE(ctx, max_attempts >= 1, "nested update: max_attempts must be positive");
    REQUIRE(ctx, !(D.sampled && rng->mode == FBA_RNG_REPLAY), "nested update: sampled Dirichlets run in PHILOX mode only");
    CU(ctx, cudaSetDevice(ctx->device));
    // update_step = static_cast<float>(1.0 / static_cast<float>(_bottom_filter_size))  (NestedBelief.cpp:133)
    float const amount = (float)(1.0 / (double)(float)n->n_bottom);
    bool const lr      = n->m->long_rows;
    int rc;
    CU(ctx, cudaMemsetAsync(n->d_failed, 0, sizeof(int), ctx->stream));
    if (rng->mode == FBA_RNG_REPLAY)
    { // top particle i draws from the i-th equal slice of the remaining words
        long long const per = (rng->n_words - rng->cursor) / n->n_top;
        REQUIRE(ctx, per >= 1, "nested update: the replay stream is shorter than one word per top particle");
        if ((rc = stage_words(ctx, rng, per * n->n_top))) return rc;
        if ((rc = clear_flag(ctx))) return rc;
        RngArgs const ra = replay_args(ctx, per * n->n_top, per, false);
        if (lr) LAUNCH_NESTED(true, true, false);
        else
            LAUNCH_NESTED(true, false, false);
        rng->cursor += per * n->n_top;
    } else if (ctx->nested_exact)
    {
        RngArgs const ra = philox_args(rng);
        if (D.sampled)
        {
            if (lr) LAUNCH_NESTED(false, true, true);
            else
                LAUNCH_NESTED(false, false, true);
        } else if (lr)
            LAUNCH_NESTED(false, true, false);
        else
            LAUNCH_NESTED(false, false, false);
    } else
    { // one warp per top particle, 32 attempts per round (k_nested_update_warp)
        RngArgs const ra = philox_args(rng);
#define LAUNCH_NESTED_WARP(L, S)                                                                                   \
    LAUNCH(ctx, (k_nested_update_warp<L, S>), blocks_for(n->n_top * 32, 64), 64, D, b->counts[b->cur], b->stride,    \
           b->sid[b->cur], b->w, n->n_top, n->n_bottom, n->states[n->cur], n->states[n->cur ^ 1], (int)action,     \
           (int)observation, amount, (long long)max_attempts, ra, n->d_attempts, n->d_failed)
        if (D.sampled)
        {
            if (lr) LAUNCH_NESTED_WARP(true, true);
            else
                LAUNCH_NESTED_WARP(false, true);
        } else if (lr)
            LAUNCH_NESTED_WARP(true, false);
        else
            LAUNCH_NESTED_WARP(false, false);
#undef LAUNCH_NESTED_WARP
    }
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, n->d_failed, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (attempts)
        CU(ctx, cudaMemcpyAsync(attempts, n->d_attempts, (size_t)n->n_top * sizeof(long long), cudaMemcpyDeviceToHost,
                                ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag)
    {
        ctx->err = "nested update: a top particle ran out of attempts (max_attempts) or of replay words before "
                   "its bottom filter was full";
        return rng->mode == FBA_RNG_REPLAY ? FBA_ERR_RNG_UNDERRUN : FBA_ERR_CAPACITY;
    }
    if (rng->mode == FBA_RNG_REPLAY && (rc = check_flag(ctx))) return rc; // a slice that ran out inside its last attempt
    n->cur ^= 1;
    return nested_normalize(n);
}
#undef LAUNCH_NESTED

// NestedBelief::sample (NestedBelief.cpp:117-127): a weighted draw of a top particle, then a uniform draw
// from its bottom filter
extern "C" int fba_nested_sample(fba_nested* n, fba_rng* rng, int64_t* top_index, int32_t* state)
{
    if (!n || !rng || !top_index || !state) return FBA_ERR_INVALID;
    fba_ctx* ctx = n->ctx;
    int rc       = fba_belief_sample(n->top, rng, top_index);
    if (rc) return rc;
    HostDraws g(rng);
    int const j = g.k((uint32_t)n->n_bottom);
    if (g.overrun()) return ctx->err = "replay stream underrun in nested sample", FBA_ERR_RNG_UNDERRUN;
    g.commit();
    CU(ctx, cudaMemcpyAsync(state, n->states[n->cur] + *top_index * n->n_bottom + j, sizeof(int), cudaMemcpyDeviceToHost,
                            ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}

// ---- POMCP, tree on the device -------------------------------------------------------------------

struct fba_tree
{
    fba_ctx* ctx   = nullptr;
    fba_model* m   = nullptr;
    long long max_sims = 0;
    int max_depth  = 0;
    unsigned table = 0; // slots (power of two); node `table` is the root
    unsigned long long* keys = nullptr;
    int* visits    = nullptr;
    TreeStat* stat = nullptr;
    long long path_cap = 0; // simulations per wave the path scratch holds
    int *path_node = nullptr, *path_action = nullptr;
    double* path_reward = nullptr;
    int* d_overflow = nullptr;
};

extern "C" void fba_tree_destroy(fba_tree* t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->keys), cudaFree(t->visits), cudaFree(t->stat);
    cudaFree(t->path_node), cudaFree(t->path_action), cudaFree(t->path_reward), cudaFree(t->d_overflow);
    delete t;
}

extern "C" int fba_tree_create(fba_ctx* ctx, fba_model* m, int64_t max_simulations, int32_t max_depth,
                               fba_tree** out)
{
    if (!ctx || !m || !out) return FBA_ERR_INVALID;
    *out = nullptr;
    REQUIRE(ctx, max_simulations >= 1 && max_simulations <= (1ll << 26), "tree: max_simulations must be in [1, 2^26]");
    REQUIRE(ctx, max_depth >= 1 && max_depth <= 4096, "tree: max_depth must be in [1, 4096]");
    CU(ctx, cudaSetDevice(ctx->device));
    auto t       = new fba_tree();
    t->ctx       = ctx;
    t->m         = m;
    t->max_sims  = max_simulations;
    t->max_depth = max_depth;
    t->table     = 1024;
    while ((long long)t->table < 2 * max_simulations) t->table <<= 1; // load factor <= 1/2
    size_t const nodes = (size_t)t->table + 1, A = (size_t)m->dev.A;
    cudaError_t e = cudaMalloc(&t->keys, (size_t)t->table * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&t->visits, nodes * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&t->stat, nodes * A * sizeof(TreeStat));
    if (e == cudaSuccess) e = cudaMalloc(&t->d_overflow, sizeof(int));
    if (e != cudaSuccess)
    {
        ctx->err = std::string("tree alloc: ") + cudaGetErrorString(e);
        fba_tree_destroy(t);
        return FBA_ERR_CUDA;
    }
    *out = t;
    return FBA_OK;
}

// The final choice (RBAPOUCT.cpp:112: selectChanceNodeUCB without the exploration term): best mean
// return, random among ties (RBAPOUCT.cpp:204). An action no simulation has returned from has no
// estimate: the reference gives it q = 0, which only matters there when n_simulations < A because its
// sequential search tries every root action first — the ticket scheme of tree_ucb guarantees the same
// here — so such actions are left out unless nothing was visited at all.
template<class G>
static int tree_final_choice(const TreeStat* root, int A, G& g, double* q_out, int64_t* visits_out)
{
    bool any = false;
    for (int a = 0; a < A; ++a) any = any || root[a].n_done > 0;
    double best = -1.7976931348623157e308;
    int pick = 0, ties = 0;
    for (int a = 0; a < A; ++a)
    {
        int const n    = root[a].n_done;
        double const v = n > 0 ? root[a].q_sum / (double)n : 0.0;
        if (q_out) q_out[a] = v;
        if (visits_out) visits_out[a] = n;
        if (any && n == 0) continue;
        if (v > best) best = v, pick = a, ties = 1;
        else if (v == best && draw_k(g, (uint32_t)++ties) == 0)
            pick = a;
    }
    return pick;
}

extern "C" int fba_tree_search(fba_tree* t, fba_belief* b, int64_t n_sims, int32_t depth, double u, double discount,
                               int32_t wave, fba_rng* rng, int32_t* action, double* q_out, int64_t* visits_out)
{
    if (!t || !b || !rng || !action) return FBA_ERR_INVALID;
    fba_ctx* ctx      = t->ctx;
    DevModel const& D = t->m->dev;
    REQUIRE(ctx, b->m == t->m && b->ctx == ctx, "tree_search: the belief must use the tree's model and context");
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "tree_search: PHILOX mode only");
    REQUIRE(ctx, n_sims >= 1 && n_sims <= t->max_sims, "tree_search: n_simulations must be in [1, max_simulations]");
    REQUIRE(ctx, depth >= 0 && depth <= t->max_depth, "tree_search: depth must be in [0, max_depth]");
    REQUIRE(ctx, wave >= 1, "tree_search: wave must be at least 1");
    REQUIRE(ctx, discount > 0 && discount <= 1 && u >= 0, "tree_search: discount in (0, 1], u >= 0");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc;
    long long const W = std::min<long long>(wave, n_sims);
    if (W > t->path_cap)
    {
        cudaFree(t->path_node), cudaFree(t->path_action), cudaFree(t->path_reward);
        t->path_node = t->path_action = nullptr, t->path_reward = nullptr;
        t->path_cap = 0;
        size_t const cells = (size_t)W * t->max_depth;
        CU(ctx, cudaMalloc(&t->path_node, cells * sizeof(int)));
        CU(ctx, cudaMalloc(&t->path_action, cells * sizeof(int)));
        CU(ctx, cudaMalloc(&t->path_reward, cells * sizeof(double)));
        t->path_cap = W;
    }
    size_t const nodes = (size_t)t->table + 1, A = (size_t)D.A;
    CU(ctx, cudaMemsetAsync(t->keys, 0xFF, (size_t)t->table * sizeof(unsigned long long), ctx->stream));
    CU(ctx, cudaMemsetAsync(t->visits, 0, nodes * sizeof(int), ctx->stream));
    CU(ctx, cudaMemsetAsync(t->stat, 0, nodes * A * sizeof(TreeStat), ctx->stream));
    CU(ctx, cudaMemsetAsync(t->d_overflow, 0, sizeof(int), ctx->stream));
    if (b->weighted && !b->cdf_valid)
        if ((rc = native_normalize(b, false, 1.0))) return rc;

    TreeArgs T{};
    T.keys = t->keys, T.visits = t->visits, T.stat = t->stat;
    T.mask = t->table - 1, T.root = (int)t->table;
    T.counts = b->counts[b->cur], T.stride = b->stride, T.sid = b->sid[b->cur], T.state = b->state[b->cur];
    T.cdf = b->weighted ? b->aux : nullptr, T.N = b->N;
    T.base = b->base, T.base_stride = b->lstride, T.proto_sid = b->d_proto_sid;
    T.depth = depth, T.u = u, T.discount = discount;
    T.path_node = t->path_node, T.path_action = t->path_action, T.path_reward = t->path_reward;
    T.overflow = t->d_overflow;
    RngArgs ra{};
    ra.seed   = rng->seed;
    ra.offset = rng->offset++;
    for (long long first = 0; first < n_sims; first += W)
    {
        T.first_sim = first;
        T.n_wave    = std::min(W, n_sims - first);
        // spread a small wave over as many SMs as possible (latency-bound: see fba_rollouts)
        int tpb = kThreads;
        while (tpb > 32 && (T.n_wave + tpb - 1) / tpb < 2ll * ctx->sm_count) tpb >>= 1;
        int const grid = blocks_for(T.n_wave, tpb);
        if (b->delta_cap > 0)
        {
            if (D.sampled) LAUNCH(ctx, (k_pomcp_wave<true, true, true>), grid, tpb, D, T, ra);
            else
                LAUNCH(ctx, (k_pomcp_wave<true, true, false>), grid, tpb, D, T, ra);
        } else if (D.sampled)
        {
            if (b->m->long_rows) LAUNCH(ctx, (k_pomcp_wave<false, true, true>), grid, tpb, D, T, ra);
            else
                LAUNCH(ctx, (k_pomcp_wave<false, false, true>), grid, tpb, D, T, ra);
        } else
        {
            if (b->m->long_rows) LAUNCH(ctx, (k_pomcp_wave<false, true, false>), grid, tpb, D, T, ra);
            else
                LAUNCH(ctx, (k_pomcp_wave<false, false, false>), grid, tpb, D, T, ra);
        }
    }
    std::vector<TreeStat> root(A);
    CU(ctx, cudaMemcpyAsync(root.data(), t->stat + (size_t)t->table * A, A * sizeof(TreeStat), cudaMemcpyDeviceToHost,
                            ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, t->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    REQUIRE(ctx, *ctx->h_flag == 0, "tree_search: the node table overflowed");
    PhiloxRng g(rng->seed, 0, rng->offset++);
    *action = tree_final_choice(root.data(), D.A, g, q_out, visits_out);
    return FBA_OK;
}

// ---- many independent runs ------------------------------------------------------------------------

struct fba_runs
{
    fba_belief* b = nullptr;
    int R         = 0;
    long long n   = 0;
    int n_tiles   = 0;
    double* tile  = nullptr;
    int2* pairs   = nullptr;
    int* totals   = nullptr;
    double* scal  = nullptr;
    long long* picked = nullptr;
    int *d_a = nullptr, *d_o = nullptr;
    unsigned char* d_active       = nullptr;
    unsigned long long* d_copies = nullptr;
    // fba_runs_plan: one search tree per run in one hash table (allocated on first use)
    unsigned long long table = 0; // slots; node table + r is the root of run r
    unsigned long long* keys = nullptr;
    int* visits    = nullptr;
    TreeStat* stat = nullptr;
    long long path_cells = 0;
    int *path_node = nullptr, *path_action = nullptr, *d_depth = nullptr, *d_overflow = nullptr;
    double* path_reward = nullptr;
};

extern "C" void fba_runs_destroy(fba_runs* r)
{
    if (!r) return;
    if (r->b)
    {
        cudaSetDevice(r->b->ctx->device);
        cudaStreamSynchronize(r->b->ctx->stream);
    }
    cudaFree(r->tile), cudaFree(r->pairs), cudaFree(r->totals), cudaFree(r->scal), cudaFree(r->picked);
    cudaFree(r->d_a), cudaFree(r->d_o), cudaFree(r->d_active), cudaFree(r->d_copies);
    cudaFree(r->keys), cudaFree(r->visits), cudaFree(r->stat);
    cudaFree(r->path_node), cudaFree(r->path_action), cudaFree(r->path_reward), cudaFree(r->d_depth);
    cudaFree(r->d_overflow);
    fba_belief_destroy(r->b);
    delete r;
}

extern "C" int fba_runs_create(fba_ctx* ctx, fba_model* m, int32_t n_runs, int64_t particles_per_run,
                               int64_t stride, fba_runs** out)
{
    if (!ctx || !m || !out) return FBA_ERR_INVALID;
    *out = nullptr;
    REQUIRE(ctx, n_runs >= 1 && particles_per_run >= 1, "runs: n_runs and particles_per_run must be >= 1");
    REQUIRE(ctx, m->delta_cap == 0, "runs: dense storage only (small beliefs are what gets batched)");
    REQUIRE(ctx, (long long)n_runs * particles_per_run < (1ll << 31), "runs: at most 2^31-1 particles in total");
    auto r = new fba_runs();
    int rc = fba_belief_create(ctx, m, (long long)n_runs * particles_per_run, stride, 1, &r->b);
    if (rc)
    {
        delete r;
        return rc;
    }
    r->R       = n_runs;
    r->n       = particles_per_run;
    r->n_tiles = (int)((particles_per_run + kTile - 1) / kTile);
    size_t const nt = (size_t)n_runs * r->n_tiles;
    cudaError_t e   = cudaMalloc(&r->tile, nt * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&r->pairs, nt * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc(&r->totals, (size_t)n_runs * 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&r->scal, (size_t)n_runs * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&r->picked, (size_t)n_runs * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&r->d_a, (size_t)n_runs * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&r->d_o, (size_t)n_runs * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&r->d_active, (size_t)n_runs);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_copies, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(r->d_copies, 0, sizeof(unsigned long long));
    if (e != cudaSuccess)
    {
        ctx->err = std::string("runs alloc: ") + cudaGetErrorString(e);
        fba_runs_destroy(r);
        return FBA_ERR_CUDA;
    }
    *out = r;
    return FBA_OK;
}

extern "C" fba_belief* fba_runs_belief(fba_runs* r)
{
    return r ? r->b : nullptr;
}

extern "C" int fba_runs_init_sampled(fba_runs* r, int32_t n_protos, const int32_t* proto_struct_id,
                                     const float* proto_counts, const double* proto_probs, fba_rng* rng)
{
    if (!r || !rng) return FBA_ERR_INVALID;
    fba_belief* b = r->b;
    fba_ctx* ctx  = b->ctx;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "runs: PHILOX mode only");
    REQUIRE(ctx, n_protos >= 1 && proto_struct_id && proto_counts, "runs_init_sampled: prototypes required");
    for (int p = 0; p < n_protos; ++p)
        REQUIRE(ctx, proto_struct_id[p] >= 0 && proto_struct_id[p] < b->m->n_structs,
                "runs_init_sampled: unknown structure id");
    CU(ctx, cudaSetDevice(ctx->device));
    DevTmp<float> d_protos;
    DevTmp<int> d_psid, d_pp, d_ps;
    DevTmp<double> d_cdf;
    size_t const pc = (size_t)n_protos * b->lstride;
    CU(ctx, cudaMalloc(&d_protos, pc * sizeof(float)));
    CU(ctx, cudaMalloc(&d_psid, n_protos * sizeof(int)));
    CU(ctx, cudaMalloc(&d_pp, (size_t)b->N * sizeof(int)));
    CU(ctx, cudaMalloc(&d_ps, (size_t)b->N * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(d_protos, proto_counts, pc * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_psid, proto_struct_id, n_protos * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double> cdf;
    if (proto_probs)
    { // the same host arithmetic as fba_belief_init_sampled
        double acc = 0, tot = 0;
        for (int p = 0; p < n_protos; ++p) tot += proto_probs[p];
        for (int p = 0; p < n_protos; ++p) cdf.push_back((acc += proto_probs[p]) / tot);
        CU(ctx, cudaMalloc(&d_cdf, n_protos * sizeof(double)));
        CU(ctx, cudaMemcpyAsync(d_cdf, cdf.data(), n_protos * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    LAUNCH(ctx, k_runs_draw_init, blocks_for(b->N), kThreads, b->m->dev, r->n, b->N, n_protos,
           (const double*)d_cdf, (int*)d_pp, (int*)d_ps, philox_args(rng));
    LAUNCH(ctx, k_init_from_protos, stream_grid(ctx, b->N), kThreads, b->counts[b->cur], b->stride,
           b->state[b->cur], b->sid[b->cur], b->w, b->N, (const float*)d_protos, (const int*)d_psid,
           (const int*)d_pp, (const int*)d_ps);
    LAUNCH(ctx, k_fill, blocks_for(b->N), kThreads, b->w, b->N, 1.0 / (double)r->n);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now = false;
    return FBA_OK;
}

extern "C" int fba_runs_init(fba_runs* r, int32_t n_protos, const int32_t* proto_struct_id, const float* proto_counts,
                             const int32_t* particle_proto, const int32_t* particle_state)
{
    if (!r) return FBA_ERR_INVALID;
    fba_belief* b = r->b;
    fba_ctx* ctx  = b->ctx;
    int const rc  = fba_belief_init(b, n_protos, proto_struct_id, proto_counts, particle_proto, particle_state);
    if (rc) return rc;
    LAUNCH(ctx, k_fill, blocks_for(b->N), kThreads, b->w, b->N, 1.0 / (double)r->n); // uniform PER RUN
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    b->suffix_valid = b->cdf_valid = false;
    b->uniform_now = false;
    return FBA_OK;
}

static RunsArgs runs_args(fba_runs* r)
{
    fba_belief* b = r->b;
    RunsArgs A{};
    A.counts = b->counts[b->cur], A.stride = b->stride, A.state = b->state[b->cur], A.sid = b->sid[b->cur];
    A.w = b->w, A.cdf = b->aux, A.noff = b->noff, A.escan = b->escan, A.dead = b->dead, A.src_of = b->src_of;
    A.tile = r->tile, A.tile_pairs = r->pairs, A.totals = r->totals, A.scal = r->scal, A.picked = r->picked;
    A.n = r->n, A.n_tiles = r->n_tiles;
    A.action = r->d_a, A.observation = r->d_o, A.active = nullptr;
    A.struct_size = b->m->d_sizes;
    A.copies      = r->d_copies;
    return A;
}

static int runs_stage_active(fba_runs* r, const uint8_t* active, RunsArgs& A)
{
    fba_ctx* ctx = r->b->ctx;
    if (!active) return FBA_OK;
    CU(ctx, cudaMemcpyAsync(r->d_active, active, (size_t)r->R, cudaMemcpyHostToDevice, ctx->stream));
    A.active = r->d_active;
    return FBA_OK;
}

// k_runs_step<LONG, SAMPLED, MODE> chosen at run time
#define LAUNCH_RUNS(ctx, mode, longrows, sampled, grid, ...)                                       \
    do {                                                                                           \
        if (sampled)                                                                               \
        {                                                                                          \
            if (longrows) LAUNCH(ctx, (k_runs_step<true, true, mode>), grid, kThreads, __VA_ARGS__);   \
            else                                                                                   \
                LAUNCH(ctx, (k_runs_step<false, true, mode>), grid, kThreads, __VA_ARGS__);        \
        } else                                                                                     \
        {                                                                                          \
            if (longrows) LAUNCH(ctx, (k_runs_step<true, false, mode>), grid, kThreads, __VA_ARGS__);  \
            else                                                                                   \
                LAUNCH(ctx, (k_runs_step<false, false, mode>), grid, kThreads, __VA_ARGS__);       \
        }                                                                                          \
    } while (0)

extern "C" int fba_runs_update_estimation(fba_runs* r, const int32_t* action, const int32_t* observation,
                                          const uint8_t* active, fba_rng* rng, double* likelihood)
{
    if (!r || !rng) return FBA_ERR_INVALID;
    fba_belief* b     = r->b;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "runs: PHILOX mode only");
    REQUIRE(ctx, action && observation, "runs_update_estimation: action and observation are required");
    for (int k = 0; k < r->R; ++k)
    {
        if (active && !active[k]) continue;
        REQUIRE(ctx, action[k] >= 0 && action[k] < D.A, "runs_update_estimation: action out of range");
        REQUIRE(ctx, observation[k] >= 0 && observation[k] < D.O, "runs_update_estimation: observation out of range");
    }
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(r->d_a, action, (size_t)r->R * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(r->d_o, observation, (size_t)r->R * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RunsArgs A = runs_args(r);
    int rc     = runs_stage_active(r, active, A);
    if (rc) return rc;
    RngArgs ra{};
    ra.seed   = rng->seed;
    ra.offset = rng->offset;
    rng->offset += 2; // update + resample, as fba_belief_update_estimation
    LAUNCH_RUNS(ctx, 0, b->m->long_rows, D.sampled != 0, r->R, D, A, ra);
    if (likelihood)
    {
        std::vector<double> tmp((size_t)r->R);
        CU(ctx, cudaMemcpyAsync(tmp.data(), r->scal, (size_t)r->R * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < r->R; ++k)
            if (!active || active[k]) likelihood[k] = tmp[k];
    }
    return FBA_OK;
}

extern "C" int fba_runs_reset_domain_states(fba_runs* r, const uint8_t* active, fba_rng* rng)
{
    if (!r || !rng) return FBA_ERR_INVALID;
    fba_belief* b     = r->b;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "runs: PHILOX mode only");
    CU(ctx, cudaSetDevice(ctx->device));
    RunsArgs A = runs_args(r);
    int rc     = runs_stage_active(r, active, A);
    if (rc) return rc;
    RngArgs ra{};
    ra.seed   = rng->seed;
    ra.offset = rng->offset;
    rng->offset += 2; // resample + start states, as fba_belief_reset_domain_states
    LAUNCH(ctx, (k_runs_step<false, false, 1>), r->R, kThreads, D, A, ra); // no categorical draws: one variant
    return FBA_OK;
}

extern "C" int fba_runs_sample(fba_runs* r, const uint8_t* active, fba_rng* rng, int64_t* index)
{
    if (!r || !rng || !index) return FBA_ERR_INVALID;
    fba_belief* b     = r->b;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "runs: PHILOX mode only");
    CU(ctx, cudaSetDevice(ctx->device));
    RunsArgs A = runs_args(r);
    int rc     = runs_stage_active(r, active, A);
    if (rc) return rc;
    RngArgs ra{};
    ra.seed   = rng->seed;
    ra.offset = rng->offset;
    rng->offset += 1; // as fba_belief_sample
    LAUNCH(ctx, (k_runs_step<false, false, 2>), r->R, kThreads, D, A, ra);
    std::vector<long long> tmp((size_t)r->R);
    CU(ctx, cudaMemcpyAsync(tmp.data(), r->picked, (size_t)r->R * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < r->R; ++k)
        if (!active || active[k]) index[k] = tmp[k];
    return FBA_OK;
}

extern "C" int fba_runs_plan(fba_runs* r, int64_t n_sims, const int32_t* depth, double u, double discount,
                             int32_t sims_per_wave, const uint8_t* active, fba_rng* rng, int32_t* action,
                             double* q_out, int64_t* visits_out)
{
    if (!r || !rng || !depth || !action) return FBA_ERR_INVALID;
    fba_belief* b     = r->b;
    fba_ctx* ctx      = b->ctx;
    DevModel const& D = b->m->dev;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "runs: PHILOX mode only");
    REQUIRE(ctx, n_sims >= 1 && sims_per_wave >= 1, "runs_plan: n_simulations and sims_per_wave must be >= 1");
    REQUIRE(ctx, discount > 0 && discount <= 1 && u >= 0, "runs_plan: discount in (0, 1], u >= 0");
    int max_depth = 1;
    for (int k = 0; k < r->R; ++k)
    {
        if (active && !active[k]) continue;
        REQUIRE(ctx, depth[k] >= 0 && depth[k] <= 4096, "runs_plan: depth must be in [0, 4096]");
        max_depth = std::max(max_depth, (int)depth[k]);
    }
    CU(ctx, cudaSetDevice(ctx->device));
    size_t const A = (size_t)D.A;
    // one node per simulation per run, load factor <= 1/2
    unsigned long long table = 1024;
    while (table < 2ull * (unsigned long long)r->R * (unsigned long long)n_sims) table <<= 1;
    REQUIRE(ctx, table <= (1ull << 30), "runs_plan: n_runs x n_simulations exceeds 2^29 tree nodes (node ids are 32-bit)");
    if (table > r->table)
    {
        cudaFree(r->keys), cudaFree(r->visits), cudaFree(r->stat);
        r->keys = nullptr, r->visits = nullptr, r->stat = nullptr;
        r->table = 0;
        size_t const nodes = (size_t)table + (size_t)r->R;
        cudaError_t e = cudaMalloc(&r->keys, (size_t)table * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&r->visits, nodes * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&r->stat, nodes * A * sizeof(TreeStat));
        if (e != cudaSuccess)
        {
            cudaFree(r->keys), cudaFree(r->visits), cudaFree(r->stat);
            r->keys = nullptr, r->visits = nullptr, r->stat = nullptr;
            ctx->err = std::string("runs_plan: tree tables: ") + cudaGetErrorString(e);
            return FBA_ERR_CUDA;
        }
        r->table = table;
    }
    table = r->table;
    long long const W     = std::min<long long>(sims_per_wave, n_sims);
    long long const cells = (long long)r->R * W * max_depth;
    if (cells > r->path_cells)
    {
        cudaFree(r->path_node), cudaFree(r->path_action), cudaFree(r->path_reward);
        r->path_node = r->path_action = nullptr, r->path_reward = nullptr;
        r->path_cells = 0;
        CU(ctx, cudaMalloc(&r->path_node, (size_t)cells * sizeof(int)));
        CU(ctx, cudaMalloc(&r->path_action, (size_t)cells * sizeof(int)));
        CU(ctx, cudaMalloc(&r->path_reward, (size_t)cells * sizeof(double)));
        r->path_cells = cells;
    }
    if (!r->d_depth) CU(ctx, cudaMalloc(&r->d_depth, (size_t)r->R * sizeof(int)));
    if (!r->d_overflow) CU(ctx, cudaMalloc(&r->d_overflow, sizeof(int)));
    size_t const nodes = (size_t)table + (size_t)r->R;
    CU(ctx, cudaMemsetAsync(r->keys, 0xFF, (size_t)table * sizeof(unsigned long long), ctx->stream));
    CU(ctx, cudaMemsetAsync(r->visits, 0, nodes * sizeof(int), ctx->stream));
    CU(ctx, cudaMemsetAsync(r->stat, 0, nodes * A * sizeof(TreeStat), ctx->stream));
    CU(ctx, cudaMemsetAsync(r->d_overflow, 0, sizeof(int), ctx->stream));
    CU(ctx, cudaMemcpyAsync(r->d_depth, depth, (size_t)r->R * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RunsArgs RA = runs_args(r);
    int rc      = runs_stage_active(r, active, RA);
    if (rc) return rc;
    // the cdf of each run's (uniform) weights, as fba_tree_search's native_normalize on a stand-alone belief
    LAUNCH(ctx, (k_runs_step<false, false, 3>), r->R, kThreads, D, RA, RngArgs{});

    TreeArgs T{};
    T.keys = r->keys, T.visits = r->visits, T.stat = r->stat;
    T.mask = (unsigned int)(table - 1), T.root = (int)table;
    T.counts = b->counts[b->cur], T.stride = b->stride, T.sid = b->sid[b->cur], T.state = b->state[b->cur];
    T.cdf = b->aux, T.N = b->N;
    T.depth = 0, T.u = u, T.discount = discount;
    T.path_node = r->path_node, T.path_action = r->path_action, T.path_reward = r->path_reward;
    T.overflow = r->d_overflow;
    T.R = r->R, T.w = (int)W, T.run_n = r->n, T.n_sims = n_sims;
    T.depth_r = r->d_depth, T.active = RA.active;
    T.n_wave  = (long long)r->R * W;
    RngArgs ra{};
    ra.seed   = rng->seed;
    ra.offset = rng->offset++;
    int tpb = kThreads;
    while (tpb > 32 && (T.n_wave + tpb - 1) / tpb < 2ll * ctx->sm_count) tpb >>= 1;
    int const grid = blocks_for(T.n_wave, tpb);
    for (long long first = 0; first < n_sims; first += W)
    {
        T.first_sim = first;
        if (D.sampled)
        {
            if (b->m->long_rows) LAUNCH(ctx, (k_pomcp_wave<false, true, true>), grid, tpb, D, T, ra);
            else
                LAUNCH(ctx, (k_pomcp_wave<false, false, true>), grid, tpb, D, T, ra);
        } else
        {
            if (b->m->long_rows) LAUNCH(ctx, (k_pomcp_wave<false, true, false>), grid, tpb, D, T, ra);
            else
                LAUNCH(ctx, (k_pomcp_wave<false, false, false>), grid, tpb, D, T, ra);
        }
    }
    std::vector<TreeStat> roots((size_t)r->R * A);
    CU(ctx, cudaMemcpyAsync(roots.data(), r->stat + (size_t)table * A, roots.size() * sizeof(TreeStat),
                            cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, r->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    REQUIRE(ctx, *ctx->h_flag == 0, "runs_plan: the node table overflowed");
    unsigned long long const pick_offset = rng->offset++;
    for (int k = 0; k < r->R; ++k)
    {
        if (active && !active[k]) continue;
        PhiloxRng g(rng->seed + (unsigned long long)k, 0, pick_offset);
        action[k] = tree_final_choice(roots.data() + (size_t)k * A, D.A, g, q_out ? q_out + (size_t)k * A : nullptr,
                                      visits_out ? visits_out + (size_t)k * A : nullptr);
    }
    return FBA_OK;
}

extern "C" int64_t fba_runs_copies(fba_runs* r)
{
    if (!r) return -1;
    unsigned long long h = 0;
    cudaSetDevice(r->b->ctx->device);
    if (cudaMemcpyAsync(&h, r->d_copies, sizeof(h), cudaMemcpyDeviceToHost, r->b->ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(r->b->ctx->stream) != cudaSuccess) return -1;
    return (int64_t)h;
}

// ---- multi-GPU phases ---------------------------------------------------------------------------

extern "C" int fba_belief_propose(fba_belief* b, int32_t a, int32_t o, fba_rng* rng, double* local_total)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "sharded beliefs run in PHILOX mode");
    int rc = propose(b, a, o, rng, 0);
    if (rc) return rc;
    int const n_tiles = (int)((b->N + kTile - 1) / kTile);
    LAUNCH(ctx, k_tile_sums, n_tiles, kThreads, b->w, b->N, b->tile);
    LAUNCH(ctx, k_scan_tile_sums, 1, kThreads, b->tile, n_tiles, b->scal);
    if (local_total)
    { // NULL: stay asynchronous; the total is at fba_belief_scalars_ptr()[0] in stream order
        if ((rc = read_scal(b))) return rc;
        *local_total = ctx->h_scal[0];
    }
    return FBA_OK;
}

extern "C" int fba_belief_normalize(fba_belief* b, double global_total)
{
    if (!b) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, global_total > 0, "normalize: total weight must be positive");
    CU(ctx, cudaSetDevice(ctx->device));
    // tile sums from fba_belief_propose are still in b->tile (exclusive-scanned)
    int const n_tiles = (int)((b->N + kTile - 1) / kTile);
    LAUNCH(ctx, k_scale_and_scan, n_tiles, kThreads, b->w, b->N, b->tile, (const double*)nullptr, global_total,
           b->aux);
    b->cdf_valid    = true;
    b->suffix_valid = false;
    return FBA_OK;
}

// record layout of the export / import staging area: [stride floats][state][structure id][pad 8]
// quotas + exchange plan from the shard totals (host): shared by the sync and async entry points
static int shard_plan(fba_belief* b, const double* totals, int n_ranks, double u, std::vector<long long>& quota,
                      int64_t* send_plan, double* global_total)
{
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, u >= 0.0 && u < 1.0, "shard plan: u must be in [0,1)");
    double W = 0.0;
    for (int g = 0; g < n_ranks; ++g)
    {
        REQUIRE(ctx, totals[g] >= 0.0 && std::isfinite(totals[g]), "shard plan: bad shard total");
        W += totals[g];
    }
    REQUIRE(ctx, W > 0.0, "shard plan: total weight must be positive");
    long long const n_total = b->N * n_ranks;
    quota.assign(n_ranks, 0);
    long long prev = 0;
    double acc     = 0.0;
    for (int g = 0; g < n_ranks; ++g)
    {
        acc += totals[g];
        long long edge = (g == n_ranks - 1) ? n_total : (long long)std::floor(acc / W * (double)n_total - u + 1.0);
        edge     = std::min(std::max(edge, prev), n_total);
        quota[g] = edge - prev;
        prev     = edge;
    }
    if (send_plan)
    { // greedy matching of surplus to deficit in rank order (identical on every rank)
        std::fill(send_plan, send_plan + (size_t)n_ranks * n_ranks, 0);
        std::vector<long long> surplus(n_ranks), deficit(n_ranks);
        for (int g = 0; g < n_ranks; ++g)
        {
            surplus[g] = std::max(0ll, quota[g] - b->N);
            deficit[g] = std::max(0ll, b->N - quota[g]);
        }
        int h = 0;
        for (int g = 0; g < n_ranks; ++g)
            while (surplus[g] > 0)
            {
                while (deficit[h] == 0) ++h;
                long long const k = std::min(surplus[g], deficit[h]);
                send_plan[(size_t)g * n_ranks + h] += k;
                surplus[g] -= k, deficit[h] -= k;
            }
    }
    if (global_total) *global_total = W;
    return FBA_OK;
}

// Asynchronous phases 2+3: the shard totals are read from DEVICE memory (where the all-gather left
// them), the quota is computed on device, nothing here waits for the GPU. The export buffer must
// already be large enough (fba_belief_reserve_export); surplus beyond it is dropped and reported by
// fba_belief_shard_plan as FBA_ERR_CAPACITY.
extern "C" int fba_belief_shard_resample_async(fba_belief* b, const double* totals_device, int32_t n_ranks,
                                               int32_t rank, double u, fba_rng* rng)
{
    if (!b || !totals_device || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "sharded beliefs run in PHILOX mode");
    REQUIRE(ctx, ctx->inplace_resample, "shard_resample_async needs the in-place resampler");
    REQUIRE(ctx, n_ranks >= 1 && rank >= 0 && rank < n_ranks, "shard_resample_async: bad rank");
    REQUIRE(ctx, u >= 0.0 && u < 1.0, "shard_resample_async: u must be in [0,1)");
    CU(ctx, cudaSetDevice(ctx->device));
    int const n_tiles = (int)((b->N + kTile - 1) / kTile);
    LAUNCH(ctx, k_shard_quota, 1, 1, totals_device, n_ranks, rank, u, b->N, b->scal + 2, b->d_quota);
    // tile sums from fba_belief_propose are still in b->tile (exclusive-scanned)
    LAUNCH(ctx, k_scale_and_scan, n_tiles, kThreads, b->w, b->N, b->tile, (const double*)(b->scal + 2), 1.0,
           b->aux);
    b->cdf_valid    = true;
    b->suffix_valid = false;
    return resample_inplace(b, rng, 0, b->d_quota);
}

// Host half of the asynchronous update: the plan for the totals (now on the host), the export /
// import bookkeeping of this rank.
extern "C" int fba_belief_shard_plan(fba_belief* b, const double* totals, int32_t n_ranks, int32_t rank,
                                     double u, int64_t* send_plan, double* global_total)
{
    if (!b || !totals) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, n_ranks >= 1 && rank >= 0 && rank < n_ranks, "shard_plan: bad rank");
    std::vector<long long> quota;
    int rc = shard_plan(b, totals, n_ranks, u, quota, send_plan, global_total);
    if (rc) return rc;
    b->local_kept  = std::min<long long>(quota[rank], b->N);
    b->xport_count = std::max(0ll, quota[rank] - b->N);
    b->imported    = 0;
    if (b->xport_count > b->xport_cap)
    {
        ctx->err = "shard surplus of " + std::to_string(b->xport_count) + " records exceeds the export buffer ("
                   + std::to_string(b->xport_cap) + "); call fba_belief_reserve_export with a larger count";
        return FBA_ERR_CAPACITY;
    }
    return FBA_OK;
}

// Phases 2-3 in one synchronous-input call: totals[G] are the all-gathered shard totals on the HOST.
extern "C" int fba_belief_shard_resample(fba_belief* b, const double* totals, int32_t n_ranks, int32_t rank,
                                         double u, fba_rng* rng, int64_t* send_plan, double* global_total)
{
    if (!b || !totals || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, n_ranks >= 1 && rank >= 0 && rank < n_ranks, "shard_resample: bad rank");
    std::vector<long long> quota;
    double W = 0.0;
    int rc   = shard_plan(b, totals, n_ranks, u, quota, send_plan, &W);
    if (rc) return rc;
    if (global_total) *global_total = W;
    if ((rc = fba_belief_normalize(b, W))) return rc;
    return fba_belief_resample_shard(b, quota[rank], rng);
}

// Imports n_records records lying at an arbitrary DEVICE address (e.g. inside an all-gathered
// window) into the dead slots [already_imported, already_imported + n_records) the local resample
// left. Asynchronous. Only after an in-place shard resample.
extern "C" int fba_belief_import_from(fba_belief* b, const void* records_device, int64_t n_records)
{
    if (!b) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    if (n_records == 0) return FBA_OK;
    REQUIRE(ctx, records_device && n_records > 0, "import_from: null records");
    REQUIRE(ctx, b->inplace_last, "import_from: needs a preceding in-place shard resample");
    REQUIRE(ctx, b->local_kept + n_records <= b->N, "import_from: more records than empty slots");
    CU(ctx, cudaSetDevice(ctx->device));
    LAUNCH(ctx, k_import_inplace, stream_grid(ctx, n_records), kThreads, b->counts[b->cur], b->stride,
           b->state[b->cur], b->sid[b->cur], b->dead, b->totals, b->imported, (long long)n_records,
           (const char*)records_device, fba_belief_record_bytes(b));
    b->imported += n_records;
    b->local_kept += n_records;
    return FBA_OK;
}

// ---- peer-to-peer sharded update over NVLink / NVSwitch (one process per GPU, one box) -----------
// No host, no NCCL on the update path: shard totals, "dead list ready" and "records landed" travel as
// step-stamped flags through peer-mapped control blocks (fba_kernels.cuh: P2PCtrl), and the surplus
// blocks of an over-quota shard are stored by k_copy_inplace straight into dead slots of the
// destination GPU's particle array.

// base pointer + offset of a device pointer inside its allocation (IPC handles name allocations)
static int ipc_describe(fba_ctx* ctx, const void* p, unsigned char* out72)
{
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    static range_fn get_range = nullptr;
    if (!get_range)
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CU(ctx, cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
        REQUIRE(ctx, fn && qr == cudaDriverEntryPointSuccess, "p2p: cuMemGetAddressRange is not available");
        get_range = (range_fn)fn;
    }
    unsigned long long base = 0;
    size_t size             = 0;
    REQUIRE(ctx, get_range(&base, &size, (unsigned long long)(uintptr_t)p) == 0, "p2p: cuMemGetAddressRange failed");
    cudaIpcMemHandle_t hnd;
    CU(ctx, cudaIpcGetMemHandle(&hnd, (void*)(uintptr_t)base));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out72, &hnd, 64);
    long long const off = (long long)((unsigned long long)(uintptr_t)p - base);
    memcpy(out72 + 64, &off, 8);
    return FBA_OK;
}

// Allocates this rank's control block (the in-place resampler's dead list and totals move into it)
// and writes the FBA_P2P_BLOB_BYTES blob peers need to map it and the particle arrays.
extern "C" int fba_belief_p2p_export(fba_belief* b, void* blob)
{
    if (!b || !blob) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, b->weighted, "p2p: importance-sampling (weighted) beliefs only");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (!b->p2p_block)
    {
        size_t const bytes = (size_t)kP2PDeadOffset + (size_t)b->N * sizeof(int);
        char* blk          = nullptr;
        CU(ctx, cudaMalloc(&blk, bytes));
        CU(ctx, cudaMemset(blk, 0, bytes));
        cudaFree(b->dead), cudaFree(b->totals);
        b->p2p_block = blk;
        b->totals    = reinterpret_cast<P2PCtrl*>(blk)->totals;
        b->dead      = reinterpret_cast<int*>(blk + kP2PDeadOffset);
    }
    if (!b->d_step)
    {
        CU(ctx, cudaMalloc(&b->d_step, sizeof(unsigned long long)));
        CU(ctx, cudaMemset(b->d_step, 0, sizeof(unsigned long long)));
    }
    unsigned char* out = (unsigned char*)blob;
    memset(out, 0, FBA_P2P_BLOB_BYTES);
    int rc;
    if ((rc = ipc_describe(ctx, b->p2p_block, out))) return rc;
    if ((rc = ipc_describe(ctx, b->counts[b->cur], out + 72))) return rc;
    if ((rc = ipc_describe(ctx, b->state[b->cur], out + 144))) return rc;
    if ((rc = ipc_describe(ctx, b->sid[b->cur], out + 216))) return rc;
    long long const shape[2] = {b->N, b->stride};
    memcpy(out + 288, shape, 16);
    return FBA_OK;
}

// blobs: n_ranks x FBA_P2P_BLOB_BYTES (all-gathered by the host, once); maps every peer's control
// block and particle arrays into this process
extern "C" int fba_belief_p2p_open(fba_belief* b, const void* blobs, int32_t n_ranks, int32_t rank)
{
    if (!b || !blobs) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, n_ranks >= 1 && n_ranks <= kMaxRanks && rank >= 0 && rank < n_ranks, "p2p_open: bad rank");
    REQUIRE(ctx, b->p2p_block, "p2p_open: call fba_belief_p2p_export first");
    CU(ctx, cudaSetDevice(ctx->device));
    for (auto p : b->opened) cudaIpcCloseMemHandle(p); // a second call re-maps: drop the old mappings
    b->opened.clear();
    b->peers_open = false;
    std::map<std::string, void*> mapped; // one mapping per peer allocation
    for (int g = 0; g < n_ranks; ++g)
    {
        const unsigned char* blob = (const unsigned char*)blobs + (size_t)g * FBA_P2P_BLOB_BYTES;
        long long shape[2];
        memcpy(shape, blob + 288, 16);
        REQUIRE(ctx, shape[0] == b->N && shape[1] == b->stride, "p2p_open: every shard must have the same size and stride");
        if (g == rank)
        {
            b->peers.ctrl[g]   = reinterpret_cast<P2PCtrl*>(b->p2p_block);
            b->peers.counts[g] = b->counts[b->cur];
            b->peers.state[g]  = b->state[b->cur];
            b->peers.sid[g]    = b->sid[b->cur];
            continue;
        }
        void* ptr[4];
        for (int k = 0; k < 4; ++k)
        {
            std::string const key((const char*)blob + 72 * k, 64);
            auto it = mapped.find(key);
            void* base = nullptr;
            if (it != mapped.end()) base = it->second;
            else
            {
                cudaIpcMemHandle_t hnd;
                memcpy(&hnd, blob + 72 * k, 64);
                CU(ctx, cudaIpcOpenMemHandle(&base, hnd, cudaIpcMemLazyEnablePeerAccess));
                b->opened.push_back(base);
                mapped[key] = base;
            }
            long long off;
            memcpy(&off, blob + 72 * k + 64, 8);
            ptr[k] = (char*)base + off;
        }
        b->peers.ctrl[g]   = (P2PCtrl*)ptr[0];
        b->peers.counts[g] = (float*)ptr[1];
        b->peers.state[g]  = (int*)ptr[2];
        b->peers.sid[g]    = (int*)ptr[3];
    }
    b->peers.timeout_ns = 20ll * 1000 * 1000 * 1000;
    if (!b->d_plan) CU(ctx, cudaMalloc(&b->d_plan, (size_t)kMaxRanks * kMaxRanks * sizeof(long long)));
    b->p2p_ranks  = n_ranks;
    b->p2p_rank   = rank;
    b->peers_open = true;
    return FBA_OK;
}

// One global importance-sampling update + resample of a sharded belief, everything enqueued on the
// context's stream by this one call: propose, shard total, [totals to peers / plan], normalise by the
// global total, in-place systematic resample to this shard's quota, surplus blocks stored into the
// destination GPUs' dead slots, [landed stamps]. u in [0,1) must be the same on every rank.
// likelihood != NULL: the GLOBAL un-normalised weight total (sum over all shards) is copied back (8
// bytes, one stream synchronisation); NULL: fully asynchronous.
extern "C" int fba_belief_sharded_update(fba_belief* b, int32_t a, int32_t o, fba_rng* rng, double u, double* likelihood)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, b->peers_open, "sharded_update: call fba_belief_p2p_open first");
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "sharded beliefs run in PHILOX mode");
    REQUIRE(ctx, ctx->inplace_resample, "sharded_update needs the in-place resampler");
    REQUIRE(ctx, u >= 0.0 && u < 1.0, "sharded_update: u must be in [0,1)");
    REQUIRE(ctx, b->peers.counts[b->p2p_rank] == b->counts[b->cur], "sharded_update: the particle buffer changed since p2p_export");
    int rc = propose(b, a, o, rng, 0);
    if (rc) return rc;
    int const n_tiles = (int)((b->N + kTile - 1) / kTile);
    LAUNCH(ctx, k_tile_sums, n_tiles, kThreads, b->w, b->N, b->tile);
    LAUNCH(ctx, k_scan_tile_sums, 1, kThreads, b->tile, n_tiles, b->scal);
    LAUNCH(ctx, k_p2p_plan, 1, 32, reinterpret_cast<P2PCtrl*>(b->p2p_block), b->peers, b->p2p_ranks, b->p2p_rank,
           (const double*)b->scal, u, b->N, b->scal + 2, b->d_quota, b->d_plan, b->d_step, b->stats);
    LAUNCH(ctx, k_scale_and_scan, n_tiles, kThreads, b->w, b->N, b->tile, (const double*)(b->scal + 2), 1.0,
           b->aux);
    b->cdf_valid    = true;
    b->suffix_valid = false;
    if ((rc = resample_inplace(b, rng, 0, b->d_quota, true))) return rc;
    if (likelihood)
    {
        CU(ctx, cudaMemcpyAsync(ctx->h_scal, b->scal + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        *likelihood = ctx->h_scal[0];
    }
    return FBA_OK;
}

// waits (cross-rank) that timed out since creation: 0 in a healthy run. Synchronises the stream.
extern "C" int64_t fba_belief_p2p_timeouts(fba_belief* b)
{
    if (!b) return -1;
    long long h[4] = {0, 0, 0, 0};
    cudaSetDevice(b->ctx->device);
    if (cudaMemcpyAsync(h, b->stats, sizeof(h), cudaMemcpyDeviceToHost, b->ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(b->ctx->stream) != cudaSuccess) return -1;
    return h[3];
}

extern "C" int fba_belief_p2p_set_timeout(fba_belief* b, double seconds)
{
    if (!b || !(seconds > 0)) return FBA_ERR_INVALID;
    b->peers.timeout_ns = (long long)(seconds * 1e9);
    return FBA_OK;
}

extern "C" int fba_belief_reserve_export(fba_belief* b, int64_t records)
{
    if (!b || records < 0) return FBA_ERR_INVALID;
    cudaSetDevice(b->ctx->device);
    return ensure_export(b, records);
}

extern "C" int64_t fba_belief_record_bytes(const fba_belief* b)
{
    if (!b) return -1;
    return b->stride * (int64_t)sizeof(float) + 16;
}

namespace fba {
__global__ void __launch_bounds__(kThreads)
    k_export(const float* __restrict__ src, long long stride, const int* __restrict__ state,
             const int* __restrict__ sid, const int* __restrict__ anc, long long first, long long n,
             char* __restrict__ out, long long rec_bytes)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < n; r += nwarp)
    {
        long long const i = anc[first + r];
        char* rec         = out + r * rec_bytes;
        warp_copy_block(src + i * stride, reinterpret_cast<float*>(rec), (int)(stride >> 2), lane);
        if (lane == 0)
        {
            int* tail = reinterpret_cast<int*>(rec + stride * sizeof(float));
            tail[0]   = state[i];
            tail[1]   = sid[i];
        }
    }
}

__global__ void __launch_bounds__(kThreads)
    k_import(float* __restrict__ dst, long long stride, int* __restrict__ state, int* __restrict__ sid,
             double* __restrict__ w, double w_new, long long first_slot, long long n,
             const char* __restrict__ in, long long rec_bytes)
{
    int const lane        = threadIdx.x & 31;
    long long const warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long const nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < n; r += nwarp)
    {
        const char* rec      = in + r * rec_bytes;
        long long const slot = first_slot + r;
        warp_copy_block(reinterpret_cast<const float*>(rec), dst + slot * stride, (int)(stride >> 2), lane);
        if (lane == 0)
        {
            const int* tail = reinterpret_cast<const int*>(rec + stride * sizeof(float));
            state[slot]     = tail[0];
            sid[slot]       = tail[1];
            if (w) w[slot] = w_new;
        }
    }
}
} // namespace fba

extern "C" int fba_belief_resample_shard(fba_belief* b, int64_t n_offspring, fba_rng* rng)
{
    if (!b || !rng) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, rng->mode == FBA_RNG_PHILOX, "sharded beliefs run in PHILOX mode");
    REQUIRE(ctx, n_offspring >= 0 && n_offspring < (1ll << 31), "resample_shard: bad offspring count");
    REQUIRE(ctx, b->cdf_valid, "resample_shard: call fba_belief_normalize first");
    CU(ctx, cudaSetDevice(ctx->device));
    long long const kept = std::min<long long>(n_offspring, b->N);
    long long const surplus = n_offspring - kept;
    b->local_kept  = kept;
    b->xport_count = surplus;
    b->imported    = 0;
    if (ctx->inplace_resample)
    {
        return resample_inplace(b, rng, n_offspring); // asynchronous: export buffer valid in stream order
    }
    b->inplace_last = false;
    if (n_offspring == 0)
    {
        flip(b);
        b->cdf_valid = false;
        b->uniform_now = false;
        return FBA_OK;
    }
    if (n_offspring > b->anc_cap)
    {
        cudaFree(b->anc);
        b->anc     = nullptr;
        b->anc_cap = n_offspring + n_offspring / 8;
        CU(ctx, cudaMalloc(&b->anc, (size_t)b->anc_cap * sizeof(int)));
    }
    int* anc = b->anc;
    LAUNCH(ctx, k_pick_native, blocks_for(n_offspring), kThreads, b->aux, b->N, (long long)n_offspring, 1,
           philox_args(rng), anc);
    int const nx = b->cur ^ 1;
    if (int const rc = ensure_next(b)) return rc;
    LAUNCH(ctx, k_gather, stream_grid(ctx, kept), kThreads, b->counts[b->cur], b->counts[nx], b->stride,
           b->state[b->cur], b->state[nx], b->sid[b->cur], b->sid[nx], b->m->d_sizes, b->w,
           1.0 / (double)b->N, anc, kept, b->delta_cap > 0 ? 1 : 0);
    if (surplus > 0)
    {
        long long const rb = fba_belief_record_bytes(b);
        if (surplus > b->xport_cap)
        {
            cudaFree(b->xport);
            b->xport_cap = surplus + surplus / 4 + 16;
            CU(ctx, cudaMalloc(&b->xport, b->xport_cap * rb));
        }
        LAUNCH(ctx, k_export, stream_grid(ctx, surplus), kThreads, b->counts[b->cur], b->stride,
               b->state[b->cur], b->sid[b->cur], anc, kept, surplus, b->xport, rb);
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    flip(b);
    b->total_weight = 1.0;
    b->cdf_valid = b->suffix_valid = false;
    b->uniform_now = false;
    return FBA_OK;
}

// copies made / resamples run by the in-place resampler since the belief was created
extern "C" int64_t fba_belief_dropped_records(fba_belief* b)
{
    if (!b) return -1;
    long long h = 0;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaMemcpy(&h, b->stats + 2, sizeof(h), cudaMemcpyDeviceToHost);
    return h;
}

// REPLAY, large beliefs: segments the serial pass had to recompute since creation (the rest were
// shifted); -1 if the parallel evaluation never ran
extern "C" int64_t fba_belief_chain_recomputed(fba_belief* b)
{
    if (!b) return -1;
    if (!b->chain_stats) return -1;
    long long h = 0;
    cudaSetDevice(b->ctx->device);
    if (cudaMemcpyAsync(&h, b->chain_stats, sizeof(h), cudaMemcpyDeviceToHost, b->ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(b->ctx->stream) != cudaSuccess) return -1;
    return h;
}

extern "C" int fba_belief_resample_stats(fba_belief* b, int64_t* copies, int64_t* resamples)
{
    if (!b) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    long long h[2] = {0, 0};
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaMemcpy(h, b->stats, sizeof(h), cudaMemcpyDeviceToHost));
    if (copies) *copies = h[0];
    if (resamples) *resamples = h[1];
    return FBA_OK;
}

extern "C" int64_t fba_belief_export_count(const fba_belief* b)
{
    return b ? b->xport_count : 0;
}
extern "C" void* fba_belief_export_ptr(fba_belief* b)
{
    return b ? b->xport : nullptr;
}
extern "C" void* fba_belief_import_ptr(fba_belief* b, int64_t n_records)
{
    if (!b || n_records <= 0) return nullptr;
    cudaSetDevice(b->ctx->device);
    if (n_records > b->import_cap)
    {
        cudaFree(b->import_buf);
        b->import_buf = nullptr;
        b->import_cap = n_records + n_records / 4 + 16;
        if (cudaMalloc(&b->import_buf, b->import_cap * fba_belief_record_bytes(b)) != cudaSuccess)
        {
            b->import_cap = 0;
            return nullptr;
        }
    }
    return b->import_buf;
}

extern "C" int fba_belief_import(fba_belief* b, int64_t n_records)
{
    if (!b) return FBA_ERR_INVALID;
    fba_ctx* ctx = b->ctx;
    REQUIRE(ctx, n_records >= 0 && b->local_kept + n_records <= b->N, "import: more records than empty slots");
    if (n_records == 0) return FBA_OK;
    REQUIRE(ctx, b->import_buf && n_records <= b->import_cap, "import: call fba_belief_import_ptr first");
    CU(ctx, cudaSetDevice(ctx->device));
    if (b->inplace_last)
    {
        LAUNCH(ctx, k_import_inplace, stream_grid(ctx, n_records), kThreads, b->counts[b->cur], b->stride,
               b->state[b->cur], b->sid[b->cur], b->dead, b->totals, 0ll, (long long)n_records, b->import_buf,
               fba_belief_record_bytes(b));
        b->local_kept += n_records;
        return FBA_OK; // asynchronous
    }
    LAUNCH(ctx, k_import, stream_grid(ctx, n_records), kThreads, b->counts[b->cur], b->stride, b->state[b->cur],
           b->sid[b->cur], b->w, 1.0 / (double)b->N, b->local_kept, (long long)n_records, b->import_buf,
           fba_belief_record_bytes(b));
    b->local_kept += n_records;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return FBA_OK;
}
