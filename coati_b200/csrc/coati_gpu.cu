// libcoati_gpu.so -- C ABI (include/coati_gpu.h) over the sm_100a marginal Gotoh kernels.
//
// Host side of the hot path: model upload, batch planning (length-ordered work list, direction
// buffer chunks), kernel launches on one stream per context, result gathering.  No CPU fallback:
// if CUDA is unavailable every entry point returns COATI_GPU_E_CUDA.
#include "../../include/coati_gpu.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"
#include "forward.cuh"
#include "forward_band.cuh"
#include "sample_spec.cuh"
#include "traceback.cuh"
#include "viterbi_generic.cuh"
#include "viterbi_pipe.cuh"
#include "viterbi_pipe1.cuh"
#include "viterbi_pipe3.cuh"
#include "viterbi_wave1.cuh"

using namespace coati_gpu;

// COATI_GPU_TRACE=1: host-side timeline on stderr (ms since the first mark of the process)
static bool trace_on() {
    static const bool on = std::getenv("COATI_GPU_TRACE") != nullptr;
    return on;
}
static void trace_mark(const char* what, size_t id) {
    if(!trace_on()) return;
    static const auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "[coati_gpu trace] %9.2f ms  %-22s %zu\n", ms, what, id);
}


// ---------------------------------------------------------------------------------------------
// Grow-only device memory pool: batches borrow blocks and give them back, so repeated calls of the
// public batch entry point do not pay cudaMalloc/cudaFree (which synchronise the device).
struct DevPool {
    struct Block {
        void* p;
        size_t bytes;
        bool used;
    };
    std::vector<Block> blocks;
    void* take(size_t bytes, cudaError_t* err) {
        *err = cudaSuccess;
        if(bytes == 0) return nullptr;
        Block* best = nullptr;
        for(Block& b : blocks)
            if(!b.used && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
        if(best && best->bytes <= 2 * bytes + (4u << 20)) {
            best->used = true;
            return best->p;
        }
        // new block: some headroom, so the slightly larger sibling sub-batches of a pipelined call
        // find it big enough (cudaMalloc synchronises the device: a miss stalls the pipeline)
        void* p = nullptr;
        trace_mark("    pool miss (MiB)", bytes >> 20);
        size_t padded = (bytes + std::min<size_t>(bytes / 8, size_t(1) << 30) + 511) & ~size_t(511);
        *err = cudaMalloc(&p, padded);
        if(*err != cudaSuccess) {
            cudaGetLastError();
            padded = bytes;
            *err = cudaMalloc(&p, padded);
        }
        if(*err != cudaSuccess) {  // release idle blocks and retry once
            cudaGetLastError();
            trim();
            *err = cudaMalloc(&p, padded);
            if(*err != cudaSuccess) return nullptr;
        }
        blocks.push_back(Block{p, padded, true});
        return p;
    }
    void give(void* p) {
        for(Block& b : blocks)
            if(b.p == p) b.used = false;
    }
    void trim() {
        for(size_t i = 0; i < blocks.size();) {
            if(!blocks[i].used) {
                cudaFree(blocks[i].p);
                blocks.erase(blocks.begin() + i);
            } else {
                ++i;
            }
        }
    }
    size_t idle_bytes() const {
        size_t n = 0;
        for(const Block& b : blocks)
            if(!b.used) n += b.bytes;
        return n;
    }
};

// same idea for pinned host staging (a D2H copy into pageable memory blocks the host until the whole
// stream has drained, which would serialise the sub-batch pipeline)
struct HostPool {
    struct Block {
        void* p;
        size_t bytes;
        bool used;
    };
    std::vector<Block> blocks;
    void* take(size_t bytes) {
        if(bytes == 0) return nullptr;
        Block* best = nullptr;
        for(Block& b : blocks)
            if(!b.used && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
        if(best) {
            best->used = true;
            return best->p;
        }
        void* p = nullptr;
        const size_t padded = (bytes + bytes / 8 + 4095) & ~size_t(4095);
        if(cudaMallocHost(&p, padded) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        blocks.push_back(Block{p, padded, true});
        return p;
    }
    void give(void* p) {
        for(Block& b : blocks)
            if(b.p == p) b.used = false;
    }
    void clear() {
        for(Block& b : blocks) cudaFreeHost(b.p);
        blocks.clear();
    }
};

struct coati_gpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;   // all work of the public staged API
    // lanes of the pipelined batch calls: everything of a sub-batch but its fills (H2D, encode, traceback,
    // expand, D2H) runs on the lane's high-priority stream lane_hi[i]; the fills of all lanes alternate between
    // two low-priority fill streams.  A fill's grid is persistent and takes every CTA slot, so the next fill
    // (other stream) moves in as the CTAs of the running one retire: the tail of a fill -- a few warps still on
    // their 2400-nt pairs, 2 ms of a 12 ms fill at 32 k pairs -- is covered, and there are never three fills
    // competing from the start.  Measured on C5, 1 M pairs cut into 31 sub-batches: one stream per lane (round 1)
    // 377-490 ms with convoys, one shared stream 410-490 ms (tails exposed), two alternating streams 326 ms.
    // Four lanes: two fills share the GPU and end together, the third follows at once, so with three lanes all
    // three sub-batches in flight completed within a few ms of each other and the GPU idled while the host
    // planned and uploaded the next one (host timeline, COATI_GPU_TRACE: 76 ms per three 23 ms sub-batches);
    // a fourth lane keeps one sub-batch queued behind them: 1 M pairs 330-341 -> 319 ms (3 / 4 / 5 lanes: 341 /
    // 319 / 327 ms in one session).
#ifndef COATI_GPU_NLANE
#define COATI_GPU_NLANE 4
#endif
    static constexpr int NLANE = COATI_GPU_NLANE;
    cudaStream_t fill_stream = nullptr, fill_stream2 = nullptr;
    cudaStream_t lane_hi[NLANE] = {};
    cudaDeviceProp prop{};
    bool model_set = false;
    GapConsts gap{};
    float* d_table = nullptr;  // n_models x TABLE_ROWS x TABLE_LD
    uint32_t n_models = 1, table_cap = 1;
    uint64_t launches = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;  // moved by the batch calls since creation (coati_gpu_transfer_bytes)
    std::string last_error;
    size_t dir_budget = 0;  // 0 = derive from free memory
    bool force_generic = false, no_wave = false;
    bool tb_serial = false;  // COATI_GPU_TB_SERIAL=1: long pairs walked one column at a time (A/B)
    // Rows into page-locked caller arenas: written by rows_to_host_kernel (used bytes only: half the volume, but
    // SM-issued writes) or copied as padded slots by the copy engine.  -1: the kernel when the call is one share of
    // a batch spread over several devices (their host links are shared and the volume decides: section 5 of
    // DESIGN.md), the copy when the device has the link to itself (measured at N = 1: 313 vs 317 ms per 1 M pairs).
    // COATI_GPU_ROWS_DIRECT=0 / 1 forces the copy / the kernel.
    int rows_direct = -1;
    uint32_t rows_ctas = 64;  // COATI_GPU_ROWS_CTAS: grid of rows_to_host_kernel (16: 319.5 ms, 64: 316.7 ms)
    uint32_t force_r = 0;  // COATI_GPU_FORCE_R: rows per lane of every inter-pair fill (tuning)
    uint32_t wave_r = 0;  // COATI_GPU_WAVE_R: force rows-per-lane of the wavefront kernel (tuning)
    bool forward_generic = false;  // COATI_GPU_FORWARD_GENERIC=1: the any-k Forward kernel also for k = 1 (A/B)
    int fwd_ctas_per_sm[2] = {1, 1};  // forward_band_kernel<false>, <true>
    bool sample_serial = false;  // COATI_GPU_SAMPLE_SERIAL=1: one thread draws all samples (A/B)
    bool pipe_scalar = false;  // COATI_GPU_PIPE_SCALAR=1: scalar template instead of the FADD2 kernels (A/B)
    int ctas_per_sm[32] = {};  // per entry of the kernel registry, on this device
    uint64_t model_gen = 0;    // bumped by every coati_gpu_set_model(s): Forward handles remember theirs
    DevPool pool;
    HostPool hpool;
    // Free device memory, asked afresh by every top-level batch call before its first kernel (other
    // libraries and processes allocate on the same device; a cached answer goes stale).  Never called
    // inside the sub-batch pipeline: cudaMemGetInfo can block for tens of ms beside running kernels.
    cudaError_t free_bytes(size_t* out) {
        size_t total_b = 0;
        return cudaMemGetInfo(out, &total_b);
    }
};

#define CU_TRY(ctx, expr)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if(e_ != cudaSuccess) {                                                             \
            (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e_);         \
            cudaGetLastError();                                                             \
            return e_ == cudaErrorMemoryAllocation ? COATI_GPU_E_NOMEM : COATI_GPU_E_CUDA;  \
        }                                                                                   \
    } while(0)

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevPool* pool = nullptr;
    cudaError_t alloc(size_t count, DevPool* from = nullptr) {
        release();
        n = count;
        pool = from;
        if(count == 0) return cudaSuccess;
        if(pool) {
            cudaError_t e;
            p = static_cast<T*>(pool->take(count * sizeof(T), &e));
            return e;
        }
        return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
    }
    void release() {
        if(p) {
            if(pool) pool->give(p);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct Chunk {
    uint32_t first, last;  // range in sorted order
    uint64_t dir_bytes;
    uint32_t max_la;
    uint32_t long_first = 0, long_count = 0;  // its slice of the batch's list of warp-walked long pairs
};

// a run of pairs inside a chunk that share one kernel configuration
struct Run {
    uint32_t first, last, cfg, chunk;
    uint64_t ops_off = 0;  // long pairs: op bytes and checkpoints of the segment-parallel expansion
    uint32_t ck_off = 0;
};

// ---- pipelined-kernel registry ------------------------------------------------------------------
typedef void (*pipe_kernel_t)(const PairDesc*, uint32_t, uint32_t, unsigned int*, const uint8_t*,
                              const uint8_t*, const float*, GapConsts, float4*, uint32_t, uint8_t*,
                              PairResult*);
typedef void (*pipe1_kernel_t)(const PairDesc*, uint32_t, uint32_t, unsigned int*, const uint8_t*,
                               const uint8_t*, const float*, GapConsts, float4*, uint32_t, uint8_t*,
                               PairResult*, const unsigned int*);
struct PipeCfg {
    uint32_t k, R;
    bool wave;
    uint32_t nc;  // substitution columns held per lane (16, or 4 for ACGT-only descendants)
    bool ab;      // A/B form, only eligible under COATI_GPU_PIPE_SCALAR=1 (scalar template in place of FADD2)
    pipe_kernel_t fn;    // generic-K pipelined kernel (viterbi_pipe.cuh)
    pipe1_kernel_t fn1;  // FADD2 specialisations (viterbi_pipe1.cuh, viterbi_pipe3.cuh); takes precedence
    size_t smem;
    const void* entry() const { return fn1 ? (const void*)fn1 : (const void*)fn; }
};
template <int K, int R>
PipeCfg make_cfg(bool ab) {
    return PipeCfg{(uint32_t)K, (uint32_t)R, false, 16, ab, viterbi_pipe_kernel<K, R>, nullptr,
                   (size_t)PIPE_WARPS * ((R + 3) / 4) * 16 * 32 * sizeof(float4)};
}
template <int R, bool WAVE, int NC>
PipeCfg make_cfg1() {  // K = 1: FADD2 specialisation, inter-pair (WAVE = false) or intra-pair wavefront
    pipe1_kernel_t fn = nullptr;
    if constexpr(WAVE) fn = viterbi_wave1_kernel<R, NC>;
    else fn = viterbi_pipe1_kernel<R, NC>;
    return PipeCfg{1u, (uint32_t)R, WAVE, (uint32_t)NC, false, nullptr, fn,
                   (size_t)PIPE_WARPS * ((R + 3) / 4) * NC * 32 * sizeof(float4)};
}
template <int R, int NC>
PipeCfg make_cfg3() {  // K = 3: FADD2 specialisation (viterbi_pipe3.cuh)
    return PipeCfg{3u, (uint32_t)R, false, (uint32_t)NC, false, nullptr, viterbi_pipe3_kernel<R, NC>,
                   (size_t)PIPE_WARPS * ((R + 3) / 4) * NC * 32 * sizeof(float4)};
}
// Immutable after static initialisation (contexts on several host threads read it concurrently); what
// depends on the device -- resident CTAs per SM -- lives in the context.
const PipeCfg g_pipe_cfgs[] = {
    make_cfg1<4, false, 16>(), make_cfg1<4, false, 4>(),  make_cfg1<8, false, 16>(), make_cfg1<8, false, 4>(),
    make_cfg1<10, false, 16>(), make_cfg1<10, false, 4>(), make_cfg<3, 3>(false),    make_cfg3<6, 16>(),
    make_cfg3<6, 4>(),         make_cfg1<2, true, 16>(),  make_cfg1<2, true, 4>(),   make_cfg1<4, true, 16>(),
    make_cfg1<4, true, 4>(),   make_cfg1<8, true, 16>(),  make_cfg1<8, true, 4>(),   make_cfg1<10, true, 16>(),
    make_cfg1<10, true, 4>(),
    // A/B: the scalar template where a FADD2 specialisation exists
    make_cfg<1, 4>(true),      make_cfg<1, 8>(true),      make_cfg<3, 6>(true)};
constexpr int N_PIPE_CFG = (int)(sizeof(g_pipe_cfgs) / sizeof(g_pipe_cfgs[0]));

// configurations the planner may pick for inter-pair fills: the specialisations, or (A/B run) only the
// scalar template's instances
bool cfg_eligible(const PipeCfg& pc, bool scalar_ab) {
    return scalar_ab ? (pc.fn != nullptr && pc.fn1 == nullptr) : !pc.ab;
}
// nc = 4 picks the ACGT-only variant when it exists, else falls back to the 16-column kernel
const PipeCfg* find_cfg(uint32_t k, uint32_t cfg, uint32_t nc, bool scalar_ab) {
    const PipeCfg* fallback = nullptr;
    for(const PipeCfg& pc : g_pipe_cfgs)
        if(pc.k == k && pc.R == (cfg & 0xffu) && pc.wave == ((cfg & CFG_WAVE) != 0) &&
           (pc.wave || cfg_eligible(pc, scalar_ab))) {
            if(pc.nc == nc) return &pc;
            if(pc.nc == 16) fallback = &pc;
        }
    return fallback;
}

// float4 entries of the wavefront's boundary rows: (bands + 1) rows of (lb + 4) / 2 float4 = lb + 4 float2
uint64_t wave_bnd_f4(uint32_t la, uint32_t lb, uint32_t R) {
    const uint64_t nb = (la + 32 * R - 1) / (32 * R);
    return (nb + 1) * ((lb + 4) / 2);
}

// issue-slot model of one pair on one warp: bands x steps x instructions per step (SASS counts of the interior
// step loops, tools/sass_count.py: 85 / 156.5 / 192.5 at R = 4 / 8 / 10, i.e. 14 + 17.9 R), weighted by the issue
// efficiency of that configuration; checked against whole-workload runs of each configuration alone
// (COATI_GPU_FORCE_R: 1007 / 1042 / 1137 GCUPS on C5, which the model reproduces with R = 4 at 0.965 of the others)
double pipe_cost(uint32_t la, uint32_t lb, uint32_t R) {
    const double nbands = (la + 32 * R - 1) / (32 * R);
    const double eff = R <= 4 ? 0.965 : 1.0;
    return nbands * (lb + 31.0) * (R * 17.9 + 14.0) / eff;
}

}  // namespace

struct coati_gpu_batch {
    coati_gpu_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;  // every operation of this batch is ordered on this stream ...
    cudaStream_t fill_stream = nullptr, fill_stream2 = nullptr;  // ... except the fills, fenced by events when it is another stream
    size_t npairs = 0;
    uint64_t a_total = 0, b_total = 0, out_total = 0;
    std::vector<PairDesc> descs;  // sorted (largest lattice first)
    std::vector<Chunk> chunks;
    std::vector<Run> runs;
    std::vector<int32_t> host_status;  // validation done at create time (caller order)
    uint64_t cells = 0, dir_bytes = 0, launches = 0;
    DevBuf<uint8_t> d_a, d_b, d_dirs;
    DevBuf<char> d_anc, d_des, d_out_a, d_out_b;
    DevBuf<PairDesc> d_pairs;
    DevBuf<PairResult> d_results;
    DevBuf<unsigned int> d_counters;
    DevBuf<float> d_ring;
    DevBuf<float4> d_bnd;
    DevBuf<uint32_t> d_prog;
    DevBuf<uint32_t> d_long_list;
    std::vector<uint32_t> long_list;  // sorted indices of the batch pairs walked by traceback_burst_list_kernel
    DevBuf<char> d_long_ops;
    DevBuf<uint4> d_long_ck;
    uint32_t ring_stride = 0, ring_ctas = 0, bnd_stride = 0, bnd_ctas = 0;
    uint32_t nc = 16;  // 4 when every descendant symbol of the batch is A/C/G/T (set at upload)
    bool raw = false;  // raw-sequence batch: symbols are encoded on the device
    PairResult* h_results = nullptr;    // D2H landing zone (pinned, from ctx->hpool)
    std::vector<PairResult> h_init;     // records of the pairs rejected by host-side validation
    std::vector<uint32_t> rejected;     // their indices (caller order)
    bool rejected_known = false;
    // device-visible addresses of the caller's page-locked row arenas (this batch's slice), set by the pipelined
    // calls: rows_to_host_kernel then writes the used bytes of the rows there and no D2H copy of the rows
    // follows
    char *h_out_a = nullptr, *h_out_b = nullptr;
    std::vector<cudaEvent_t> events;  // 4 per run: fill start, fill end, traceback end, compact end
    ~coati_gpu_batch() {
        for(cudaEvent_t e : events) cudaEventDestroy(e);
        if(h_results) ctx->hpool.give(h_results);
    }
};

// ---------------------------------------------------------------------------------------------
#ifdef COATI_WAVE_TRACE
extern "C" int coati_gpu_debug_wave_trace(unsigned long long* out, size_t n) {
    return (int)cudaMemcpyFromSymbol(out, g_wave_trace, n * sizeof(unsigned long long));
}
#endif
extern "C" const char* coati_gpu_strerror(int code) {
    switch(code) {
    case COATI_GPU_OK: return "success";
    case COATI_GPU_E_CUDA: return "CUDA failure or no usable sm_100 device (there is no CPU fallback).";
    case COATI_GPU_E_ARG: return "Invalid argument.";
    case COATI_GPU_E_NOMEM: return "sequences to align exceed available memory.";
    case COATI_GPU_E_SYMBOL: return "Encoded symbol outside the substitution table.";
    case COATI_GPU_E_LENGTH:
        return "Length of sequence must be multiple of gap unit length.";
    case COATI_GPU_E_AMBIGUOUS: return "Ambiguous nucleotides in ancestor/reference.";
    case COATI_GPU_E_STOP: return "Early stop codon in ancestor/reference.";
    case COATI_GPU_E_INTERNAL: return "Traceback left the lattice.";
    default: return "Unknown error.";
    }
}

extern "C" int coati_gpu_init(int device, coati_gpu_ctx** out) {
    if(!out) return COATI_GPU_E_ARG;
    *out = nullptr;
    int count = 0;
    if(cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return COATI_GPU_E_CUDA;
    }
    auto* ctx = new(std::nothrow) coati_gpu_ctx;
    if(!ctx) return COATI_GPU_E_NOMEM;
    ctx->device = device;
    if(cudaSetDevice(device) != cudaSuccess ||
       cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
       cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
       cudaMalloc(reinterpret_cast<void**>(&ctx->d_table),
                  TABLE_ROWS * TABLE_LD * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return COATI_GPU_E_CUDA;
    }
    {
        int least = 0, greatest = 0;
        bool ok = cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&ctx->fill_stream, cudaStreamNonBlocking, least) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&ctx->fill_stream2, cudaStreamNonBlocking, least) == cudaSuccess;
        for(int i = 0; ok && i < coati_gpu_ctx::NLANE; ++i)
            ok = cudaStreamCreateWithPriority(&ctx->lane_hi[i], cudaStreamNonBlocking, greatest) == cudaSuccess;
        if(!ok) {
            cudaGetLastError();
            coati_gpu_shutdown(ctx);
            return COATI_GPU_E_CUDA;
        }
    }
    if(const char* env = std::getenv("COATI_GPU_DIR_BUDGET_MB")) {
        ctx->dir_budget = static_cast<size_t>(std::strtoull(env, nullptr, 10)) << 20;
    }
    if(const char* env = std::getenv("COATI_GPU_FORCE_GENERIC")) ctx->force_generic = env[0] == '1';
    if(const char* env = std::getenv("COATI_GPU_PIPE_SCALAR")) ctx->pipe_scalar = env[0] == '1';
    if(const char* env = std::getenv("COATI_GPU_SAMPLE_SERIAL")) ctx->sample_serial = env[0] == '1';
    if(const char* env = std::getenv("COATI_GPU_FORWARD_GENERIC")) ctx->forward_generic = env[0] == '1';
    if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->fwd_ctas_per_sm[0], forward_band_kernel<false>,
                                                     FB_WARPS * 32, 0) != cudaSuccess ||
       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->fwd_ctas_per_sm[1], forward_band_kernel<true>,
                                                     FB_WARPS * 32, 0) != cudaSuccess ||
       ctx->fwd_ctas_per_sm[0] < 1 || ctx->fwd_ctas_per_sm[1] < 1) {
        cudaGetLastError();
        coati_gpu_shutdown(ctx);
        return COATI_GPU_E_CUDA;
    }
    if(const char* env = std::getenv("COATI_GPU_NO_WAVE")) ctx->no_wave = env[0] == '1';
    if(const char* env = std::getenv("COATI_GPU_TB_SERIAL")) ctx->tb_serial = env[0] == '1';
    if(const char* env = std::getenv("COATI_GPU_ROWS_DIRECT")) ctx->rows_direct = env[0] == '0' ? 0 : 1;
    if(const char* env = std::getenv("COATI_GPU_ROWS_CTAS")) ctx->rows_ctas = (uint32_t)std::max(1, std::atoi(env));
    if(const char* env = std::getenv("COATI_GPU_FORCE_R")) ctx->force_r = (uint32_t)std::atoi(env);
    if(const char* env = std::getenv("COATI_GPU_WAVE_R")) {
        const uint32_t r = (uint32_t)std::atoi(env);
        if(r == 2 || r == 4 || r == 8 || r == 10) ctx->wave_r = r;
    }
    static_assert(N_PIPE_CFG <= 32, "ctx->ctas_per_sm");
    for(int x = 0; x < N_PIPE_CFG; ++x) {  // function attributes are per device: set for this one
        const PipeCfg& pc = g_pipe_cfgs[x];
        if(cudaFuncSetAttribute(pc.entry(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)pc.smem) != cudaSuccess ||
           cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->ctas_per_sm[x], pc.entry(),
                                                         PIPE_WARPS * 32, pc.smem) != cudaSuccess ||
           ctx->ctas_per_sm[x] < 1) {
            cudaGetLastError();
            coati_gpu_shutdown(ctx);
            return COATI_GPU_E_CUDA;
        }
    }
    *out = ctx;
    return COATI_GPU_OK;
}

extern "C" void coati_gpu_shutdown(coati_gpu_ctx* ctx) {
    if(!ctx) return;
    cudaSetDevice(ctx->device);
    for(int i = 0; i < coati_gpu_ctx::NLANE; ++i) {
        cudaStream_t both[2] = {ctx->lane_hi[i], i == 0 ? ctx->fill_stream : i == 1 ? ctx->fill_stream2 : nullptr};
        for(cudaStream_t st : both)
            if(st) {
                cudaStreamSynchronize(st);
                cudaStreamDestroy(st);
            }
    }
    if(ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    if(ctx->d_table) cudaFree(ctx->d_table);
    ctx->pool.trim();
    ctx->hpool.clear();
    delete ctx;
}

extern "C" const char* coati_gpu_last_cuda_error(coati_gpu_ctx* ctx) {
    return ctx ? ctx->last_error.c_str() : "";
}
extern "C" void* coati_gpu_stream(coati_gpu_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
extern "C" uint64_t coati_gpu_launch_count(coati_gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void coati_gpu_transfer_bytes(coati_gpu_ctx* ctx, uint64_t* h2d, uint64_t* d2h) {
    if(h2d) *h2d = ctx ? ctx->h2d_bytes : 0;
    if(d2h) *d2h = ctx ? ctx->d2h_bytes : 0;
}

extern "C" int coati_gpu_device_info(coati_gpu_ctx* ctx, int* sm_count, int* clock_khz,
                                     size_t* free_bytes, size_t* total_bytes) {
    if(!ctx) return COATI_GPU_E_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if(sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if(clock_khz) {
        int khz = 0;
        CU_TRY(ctx, cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device));
        *clock_khz = khz;
    }
    size_t f = 0, t = 0;
    CU_TRY(ctx, cudaMemGetInfo(&f, &t));
    if(free_bytes) *free_bytes = f;
    if(total_bytes) *total_bytes = t;
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_set_models(coati_gpu_ctx* ctx, uint32_t n_models, const float* tables, float g,
                                    float e, uint32_t k) {
    if(!ctx || !tables || k == 0 || n_models == 0 || n_models > CFG_MAX_MODELS) return COATI_GPU_E_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    // align_pair.cc:66-69 -- host libm, float, same calls as the reference
    GapConsts c;
    c.ng = ::log1pf(-g);
    c.gs = ::log1pf(-e);
    c.go = ::logf(g);
    c.ge = ::logf(e);
    c.gk1 = c.ge * static_cast<float>(static_cast<size_t>(k - 1));  // semiring.hpp:109-111
    c.gk = c.ge * static_cast<float>(static_cast<size_t>(k));
    c.k = k;
    c.stop_gap = ::logf(g * e * e);  // utils.cc:1049
    if(n_models > ctx->table_cap) {
        CU_TRY(ctx, cudaDeviceSynchronize());
        float* nt = nullptr;
        CU_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(&nt),
                               (size_t)n_models * TABLE_ROWS * TABLE_LD * sizeof(float)));
        cudaFree(ctx->d_table);
        ctx->d_table = nt;
        ctx->table_cap = n_models;
    }
    std::vector<float> padded((size_t)n_models * TABLE_ROWS * TABLE_LD, 0.0f);
    for(uint32_t m = 0; m < n_models; ++m)
        for(int r = 0; r < TABLE_ROWS; ++r)
            for(int col = 0; col < TABLE_COLS; ++col)
                padded[((size_t)m * TABLE_ROWS + r) * TABLE_LD + col] =
                    tables[((size_t)m * TABLE_ROWS + r) * TABLE_COLS + col];
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_table, padded.data(), padded.size() * sizeof(float),
                                cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->gap = c;
    ctx->n_models = n_models;
    ctx->model_set = true;
    ++ctx->model_gen;
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_set_model(coati_gpu_ctx* ctx, const float* table, float g, float e,
                                   uint32_t k) {
    return coati_gpu_set_models(ctx, 1, table, g, e, k);
}

// ---------------------------------------------------------------------------------------------
// symbol validation: a < 183, b < 15 (the reference indexes the table unchecked, matrix.hpp:73-76)
__global__ void validate_symbols_kernel(const PairDesc* __restrict__ pairs, uint32_t npairs,
                                        const uint8_t* __restrict__ a_all,
                                        const uint8_t* __restrict__ b_all,
                                        PairResult* __restrict__ results) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if(warp >= npairs) return;
    const PairDesc pd = pairs[warp];
    bool bad = false;
    for(uint32_t x = lane; x < pd.la; x += 32) bad |= a_all[pd.a_off + x] >= TABLE_ROWS;
    for(uint32_t x = lane; x < pd.lb; x += 32) bad |= b_all[pd.b_off + x] >= TABLE_COLS;
    if(__any_sync(0xffffffffu, bad) && lane == 0 && results[pd.orig].status == 0)
        results[pd.orig].status = COATI_GPU_E_SYMBOL;
}

// raw: optional per-pair byte {bit0: ancestor ends with a stop codon, bit1: descendant does, bit7: the
// pair fails the length checks of process_marginal} for the raw-sequence entry point
static int batch_create_on(coati_gpu_ctx* ctx, cudaStream_t stream, cudaStream_t fill_stream,
                           uint64_t lane_budget, size_t npairs, const uint64_t* a_off,
                           const uint64_t* b_off, coati_gpu_batch** out, const uint8_t* raw = nullptr,
                           const uint32_t* model = nullptr);

extern "C" int coati_gpu_batch_create(coati_gpu_ctx* ctx, size_t npairs, const uint64_t* a_off,
                                      const uint64_t* b_off, coati_gpu_batch** out) {
    if(!ctx) return COATI_GPU_E_ARG;
    return batch_create_on(ctx, ctx->stream, ctx->stream, 0, npairs, a_off, b_off, out);
}

// offsets may start anywhere (a sub-range of a larger CSR pack): everything is stored relative to
// a_off[0] / b_off[0]
// lane_budget: direction-stream bytes this batch may use (0 = derive from the free device memory)
static int batch_create_on(coati_gpu_ctx* ctx, cudaStream_t stream, cudaStream_t fill_stream,
                           uint64_t lane_budget, size_t npairs, const uint64_t* a_off,
                           const uint64_t* b_off, coati_gpu_batch** out, const uint8_t* raw,
                           const uint32_t* model) {
    if(!ctx || !out || (npairs && (!a_off || !b_off))) return COATI_GPU_E_ARG;
    if(!ctx->model_set || npairs > 0xfffffff0ull) return COATI_GPU_E_ARG;
    *out = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<coati_gpu_batch> holder(new(std::nothrow) coati_gpu_batch);
    coati_gpu_batch* bt = holder.get();
    if(!bt) return COATI_GPU_E_NOMEM;
    bt->ctx = ctx;
    bt->stream = stream;
    bt->fill_stream = fill_stream;
    bt->npairs = npairs;
    const uint32_t k = ctx->gap.k;
    const uint64_t a0 = npairs ? a_off[0] : 0, b0 = npairs ? b_off[0] : 0;
    try {
        bt->descs.resize(npairs);
        bt->host_status.assign(npairs, COATI_GPU_OK);
        bt->h_results = static_cast<PairResult*>(ctx->hpool.take((npairs + 1) * sizeof(PairResult)));
        if(!bt->h_results) return COATI_GPU_E_NOMEM;
    } catch(const std::bad_alloc&) {
        return COATI_GPU_E_NOMEM;
    }
    bt->a_total = npairs ? a_off[npairs] - a0 : 0;
    bt->b_total = npairs ? b_off[npairs] - b0 : 0;
    bt->out_total = bt->a_total + bt->b_total + npairs;
    trace_mark("  plan: host alloc", npairs);
    uint32_t memo[64][2];
    for(auto& m : memo) m[0] = 0, m[1] = 0;  // la == 0 never looks itself up
    for(size_t p = 0; p < npairs; ++p) {
        PairDesc& d = bt->descs[p];
        uint64_t la = a_off[p + 1] - a_off[p], lb = b_off[p + 1] - b_off[p];
        if(la > 0x7fffffffull || lb > 0x7fffffffull) return COATI_GPU_E_ARG;
        uint32_t stop_bits = 0;
        if(raw) {  // lengths after trim_end_stops; the slots keep their full size
            if(raw[p] & 0x80) bt->host_status[p] = COATI_GPU_E_LENGTH;
            if(raw[p] & 1) la -= 3, stop_bits |= CFG_STOP_A;
            if(raw[p] & 2) lb -= 3, stop_bits |= CFG_STOP_B;
        }
        d.a_off = a_off[p] - a0;
        d.b_off = b_off[p] - b0;
        d.out_off = d.a_off + d.b_off + p;
        d.dir_off = 0;
        d.la = static_cast<uint32_t>(la);
        d.lb = static_cast<uint32_t>(lb);
        d.orig = static_cast<uint32_t>(p);
        d.cfg = 0;
        if(!ctx->force_generic && la > 0 && lb > 0) {
            // rows per lane: the cost model's column factor is the same for every candidate, so the choice
            // depends on la alone; batches repeat lengths (bins), so remember the last few answers
            uint32_t& memo_la = memo[(d.la * 2654435761u) >> 26][0];
            uint32_t& memo_r = memo[(d.la * 2654435761u) >> 26][1];
            if(memo_la != d.la) {
                double best = 0;
                uint32_t r = 0;
                for(const PipeCfg& pc : g_pipe_cfgs) {
                    if(pc.k != k || pc.wave || !cfg_eligible(pc, ctx->pipe_scalar) ||
                       (ctx->force_r && pc.R != ctx->force_r))
                        continue;
                    const double c = pipe_cost(d.la, 1000, pc.R);
                    if(r == 0 || c < best) best = c, r = pc.R;
                }
                memo_la = d.la, memo_r = r;
            }
            d.cfg = memo_r;
            // long pairs (or pairs of a batch too small to fill the GPU) run as an intra-pair wavefront
            const uint64_t cells = la * lb;
            const bool small_batch = npairs < 2048;
            if(k == 1 && !ctx->no_wave && la >= 1024 && lb >= 512 &&
               (cells >= (1ull << 26) || (small_batch && cells >= (1ull << 21)))) {
                // One band = one warp, and the chain of bands runs at the pace of its slowest member, so every band
                // wants a scheduler of its own (4 per SM): the narrowest lane tile whose band count still fits.
                // Fill time = (lb + bands * lag) steps; measured on B200 (tools/wave_lag.py, tools/wave_exp.py):
                // a step costs 109 / 164 / 277 / 332 cycles at R = 2 / 4 / 8 / 10 and the fill grows by 9 000 / 12 000 /
                // 21 400 / 24 600 cycles per band; fills at R = 2 / 4 / 8 / 10: 10k 1.13 / 1.27 / 1.88 / 2.04 ms,
                // 20k 2.35 / 2.55 / 3.69 / 4.07, 40k 6.28 / 5.20 / 7.44 / 8.27, 80k 17.8 / 15.0 / 14.8 / 16.6,
                // 160k 64.9 / 47.3 / 47.0 / 33.2 (too many bands double up on schedulers).
                const uint64_t slots = 4ull * (uint64_t)ctx->prop.multiProcessorCount;
                const uint32_t wr = (la + 63) / 64 <= slots ? 2u : (la + 127) / 128 <= slots ? 4u : (la + 255) / 256 <= slots ? 8u : 10u;
                d.cfg = wr | CFG_WAVE;
                if(ctx->wave_r) d.cfg = ctx->wave_r | CFG_WAVE;
            }
        }
        // the reference checks divisibility before trimming stops (utils.cc:819-837); a lattice
        // whose terminal cell is unreachable is undefined behaviour upstream -> reject here.
        if(la % k != 0 || lb % k != 0) bt->host_status[p] = COATI_GPU_E_LENGTH;
        d.cfg |= stop_bits;
        if(model) {
            if(model[p] >= ctx->n_models) return COATI_GPU_E_ARG;
            d.cfg |= model[p] << CFG_MODEL_SHIFT;
        }
    }
    bt->raw = raw != nullptr;
    trace_mark("  plan: descs", npairs);
    // longest-processing-time order: biggest lattices first
    // (stable counting sort on a bucketed key: kernel config, then lattice size to ~1.6 % -- LPT does
    // not need an exact order and a comparison sort of 1 M descriptors costs ~100 ms of host time)
    {
        auto bucket = [](const PairDesc& d) -> uint32_t {
            const uint64_t cells = (uint64_t)d.la * d.lb + 1;
            const int lz = 63 - __builtin_clzll(cells);
            const uint32_t mant = lz >= 6 ? (uint32_t)((cells >> (lz - 6)) & 63) : (uint32_t)(cells << (6 - lz)) & 63;
            const uint32_t size_rank = 4095u - (uint32_t)(lz * 64 + mant);  // bigger lattice first
            // config rank: wave pairs first, then by R descending (any fixed order works)
            const uint32_t cfg_rank = (d.cfg & CFG_WAVE) ? 0u : 16u - std::min(15u, d.cfg & 0xffu);
            return cfg_rank * 4096u + size_rank;
        };
        const uint32_t NB = 17u * 4096u;
        std::vector<uint32_t> hist(NB + 1, 0);
        std::vector<uint32_t> key(npairs);
        for(size_t p = 0; p < npairs; ++p) ++hist[(key[p] = bucket(bt->descs[p])) + 1];
        for(uint32_t x = 0; x < NB; ++x) hist[x + 1] += hist[x];
        std::vector<PairDesc> sorted(npairs);
        for(size_t p = 0; p < npairs; ++p) sorted[hist[key[p]]++] = bt->descs[p];
        bt->descs.swap(sorted);
    }
    // direction-buffer chunks
    trace_mark("  plan: sorted", npairs);
    uint64_t budget = ctx->dir_budget ? ctx->dir_budget : lane_budget;
    if(budget == 0) {  // (cudaMemGetInfo can block for tens of ms while kernels run: never inside the pipeline)
        size_t free_b = 0;
        CU_TRY(ctx, ctx->free_bytes(&free_b));
        free_b += ctx->pool.idle_bytes();
        const uint64_t fixed = 2 * (bt->a_total + bt->b_total) + 2 * bt->out_total +
                               npairs * (sizeof(PairDesc) + sizeof(PairResult)) + (256ull << 20);
        budget = free_b > fixed ? static_cast<uint64_t>((free_b - fixed) * 0.85) : 0;
    }
    uint64_t need_max = 0;
    {
        Chunk cur{0, 0, 0, 0};
        for(uint32_t s = 0; s < npairs; ++s) {
            PairDesc& d = bt->descs[s];
            const bool live = bt->host_status[d.orig] == COATI_GPU_OK;
            const uint64_t cells = live ? (uint64_t)d.la * d.lb : 0;
            const uint64_t bytes = !live || cells == 0 ? 0
                                   : (d.cfg & 0x1ffu) ? pipe_dir_bytes(d.la, d.lb, d.cfg & 0xffu) : cells;
            const uint64_t padded = (bytes + 127) & ~127ull;
            if(padded > budget) {
                ctx->last_error = "direction stream of one pair exceeds device memory budget";
                return COATI_GPU_E_NOMEM;
            }
            if(cur.dir_bytes + padded > budget && s > cur.first) {
                cur.last = s;
                bt->chunks.push_back(cur);
                cur = Chunk{s, s, 0, 0};
            }
            d.dir_off = cur.dir_bytes;
            cur.dir_bytes += padded;
            cur.max_la = std::max(cur.max_la, d.la);
            bt->cells += cells;
            bt->dir_bytes += bytes;
        }
        cur.last = static_cast<uint32_t>(npairs);
        if(cur.last > cur.first) bt->chunks.push_back(cur);
        for(const Chunk& c : bt->chunks) need_max = std::max(need_max, c.dir_bytes);
    }
    // runs of equal kernel configuration inside each chunk
    uint32_t max_lb_pipe = 0;
    for(size_t ci = 0; ci < bt->chunks.size(); ++ci) {
        const Chunk& c = bt->chunks[ci];
        uint32_t s0 = c.first;
        for(uint32_t s = c.first; s <= c.last; ++s) {
            if(s == c.last || (bt->descs[s].cfg & 0x1ffu) != (bt->descs[s0].cfg & 0x1ffu) ||
               (bt->descs[s0].cfg & CFG_WAVE)) {
                if(s > s0) bt->runs.push_back(Run{s0, s, bt->descs[s0].cfg & 0x1ffu, (uint32_t)ci});
                s0 = s;
            }
            if(s < c.last && (bt->descs[s].cfg & 0x1ffu)) max_lb_pipe = std::max(max_lb_pipe, bt->descs[s].lb);
        }
    }
    bt->bnd_stride = (max_lb_pipe + 2 + 32 + 7) & ~7u;  // + 32: the step loop reads ahead past column lb
    bt->bnd_ctas = 0;
    uint64_t wave_f4 = 0;
    uint32_t wave_bands = 0;
    for(const Run& r : bt->runs) {
        if(!r.cfg) continue;
        const PipeCfg* pc = find_cfg(k, r.cfg, 16, ctx->pipe_scalar);
        if(!pc) return COATI_GPU_E_ARG;
        if(r.cfg & CFG_WAVE) {
            const PairDesc& d = bt->descs[r.first];
            const uint32_t nb = (d.la + 32 * pc->R - 1) / (32 * pc->R);
            wave_bands = std::max(wave_bands, nb);
            wave_f4 = std::max<uint64_t>(wave_f4, wave_bnd_f4(d.la, d.lb, pc->R));
        } else {
            const uint32_t want = (r.last - r.first + PIPE_WARPS - 1) / PIPE_WARPS;
            const PipeCfg* pc4 = find_cfg(k, r.cfg, 4, ctx->pipe_scalar);  // ACGT-only variant may be more resident
            const uint32_t cap = (uint32_t)ctx->prop.multiProcessorCount *
                                 std::max(ctx->ctas_per_sm[pc - g_pipe_cfgs], ctx->ctas_per_sm[pc4 - g_pipe_cfgs]);
            bt->bnd_ctas = std::max(bt->bnd_ctas, std::min(want, cap));
        }
    }
    // device buffers
    uint32_t max_la = 0, n_generic = 0;
    for(const Run& r : bt->runs)
        if(r.cfg == 0) {
            n_generic += r.last - r.first;
            for(uint32_t x = r.first; x < r.last; ++x) max_la = std::max(max_la, bt->descs[x].la);
        }
    bt->ring_stride = (max_la + 1 + 31) & ~31u;
    bt->ring_ctas = static_cast<uint32_t>(
        std::min<size_t>(n_generic, (size_t)ctx->prop.multiProcessorCount * 4));
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if(e == cudaSuccess) e = r;
    };
    trace_mark("  plan: chunks+runs", npairs);
    ok(bt->d_a.alloc(bt->a_total + 1, &ctx->pool));
    ok(bt->d_b.alloc(bt->b_total + 64, &ctx->pool));  // symbol read-ahead of the step loop
    ok(bt->d_anc.alloc(bt->a_total + 1, &ctx->pool));
    ok(bt->d_des.alloc(bt->b_total + 1, &ctx->pool));
    ok(bt->d_out_a.alloc(bt->out_total + 1, &ctx->pool));
    ok(bt->d_out_b.alloc(bt->out_total + 1, &ctx->pool));
    ok(bt->d_pairs.alloc(npairs + 1, &ctx->pool));
    ok(bt->d_results.alloc(npairs + 1, &ctx->pool));
    ok(bt->d_counters.alloc(bt->runs.size() + 1, &ctx->pool));
    ok(bt->d_bnd.alloc(std::max<uint64_t>((uint64_t)bt->bnd_ctas * PIPE_WARPS * 2 * bt->bnd_stride, wave_f4),
                       &ctx->pool));
    ok(bt->d_prog.alloc(wave_bands ? wave_bands + 2 : 0, &ctx->pool));
    {
        uint64_t ops_total = 0, ck_total = 0;
        for(Run& r : bt->runs)
            if(r.cfg & CFG_WAVE) {
                const PairDesc& d = bt->descs[r.first];
                r.ops_off = ops_total, r.ck_off = (uint32_t)ck_total;
                ops_total += ((uint64_t)d.la + d.lb + 64) & ~63ull;
                ck_total += long_ck_capacity(d.la, d.lb);
            }
        ok(bt->d_long_ops.alloc(ops_total, &ctx->pool));
        ok(bt->d_long_ck.alloc(ck_total, &ctx->pool));
        if(!ctx->tb_serial)
            for(Chunk& c : bt->chunks) {
                c.long_first = (uint32_t)bt->long_list.size();
                for(uint32_t x = c.first; x < c.last; ++x)
                    if(bt->host_status[bt->descs[x].orig] == COATI_GPU_OK && burst_in_batch(bt->descs[x], k))
                        bt->long_list.push_back(x);
                c.long_count = (uint32_t)bt->long_list.size() - c.long_first;
            }
        ok(bt->d_long_list.alloc(bt->long_list.size(), &ctx->pool));
    }
    ok(bt->d_dirs.alloc(need_max + 128, &ctx->pool));
    ok(bt->d_ring.alloc((size_t)bt->ring_ctas * 3 * ring_depth(k) * bt->ring_stride, &ctx->pool));
    if(e != cudaSuccess) {
        ctx->last_error = std::string("batch allocation: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return COATI_GPU_E_NOMEM;
    }
    trace_mark("  plan: buffers", npairs);
    if(npairs) {
        CU_TRY(ctx, cudaMemcpyAsync(bt->d_pairs.p, bt->descs.data(), npairs * sizeof(PairDesc),
                                    cudaMemcpyHostToDevice, bt->stream));  // descs outlive the copy (member)
        ctx->h2d_bytes += npairs * sizeof(PairDesc) + bt->long_list.size() * sizeof(uint32_t);
    }
    if(!bt->long_list.empty())
        CU_TRY(ctx, cudaMemcpyAsync(bt->d_long_list.p, bt->long_list.data(), bt->long_list.size() * sizeof(uint32_t),
                                    cudaMemcpyHostToDevice, bt->stream));
    trace_mark("  plan: descs H2D", npairs);
    *out = holder.release();
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_batch_upload(coati_gpu_batch* bt, const uint8_t* a_all,
                                      const uint8_t* b_all, const char* anc_all,
                                      const char* des_all) {
    if(!bt) return COATI_GPU_E_ARG;
    coati_gpu_ctx* ctx = bt->ctx;
    if(bt->raw) {  // raw-sequence batch: only the symbols travel; codes are produced on the device
        if((bt->a_total && !anc_all) || (bt->b_total && !des_all)) return COATI_GPU_E_ARG;
        CU_TRY(ctx, cudaSetDevice(ctx->device));
        cudaStream_t s = bt->stream;
        if(bt->a_total)
            CU_TRY(ctx, cudaMemcpyAsync(bt->d_anc.p, anc_all, bt->a_total, cudaMemcpyHostToDevice, s));
        if(bt->b_total)
            CU_TRY(ctx, cudaMemcpyAsync(bt->d_des.p, des_all, bt->b_total, cudaMemcpyHostToDevice, s));
        bt->nc = 16;
        ctx->h2d_bytes += bt->a_total + bt->b_total;
        return COATI_GPU_OK;
    }
    if((bt->a_total && (!a_all || !anc_all)) || (bt->b_total && (!b_all || !des_all)))
        return COATI_GPU_E_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = bt->stream;
    if(bt->a_total) {
        CU_TRY(ctx, cudaMemcpyAsync(bt->d_a.p, a_all, bt->a_total, cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemcpyAsync(bt->d_anc.p, anc_all, bt->a_total, cudaMemcpyHostToDevice, s));
    }
    if(bt->b_total) {
        // ambiguity codes anywhere in the batch?  (decides the shared-memory footprint of the fill)
        uint64_t acc8 = 0;
        const uint64_t n8 = bt->b_total / 8;
        const uint64_t* w8 = reinterpret_cast<const uint64_t*>(b_all);
        if((reinterpret_cast<uintptr_t>(b_all) & 7) == 0)
            for(uint64_t x = 0; x < n8; ++x) acc8 |= w8[x];
        else
            acc8 = ~0ull;
        for(uint64_t x = n8 * 8; x < bt->b_total; ++x) acc8 |= b_all[x];
        bt->nc = (acc8 & 0xfcfcfcfcfcfcfcfcull) ? 16 : 4;
        CU_TRY(ctx, cudaMemcpyAsync(bt->d_b.p, b_all, bt->b_total, cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemcpyAsync(bt->d_des.p, des_all, bt->b_total, cudaMemcpyHostToDevice, s));
    }
    ctx->h2d_bytes += 2 * (bt->a_total + bt->b_total);
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_batch_run(coati_gpu_batch* bt) {
    if(!bt) return COATI_GPU_E_ARG;
    coati_gpu_ctx* ctx = bt->ctx;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = bt->stream;
    const uint32_t n = static_cast<uint32_t>(bt->npairs);
    bt->launches = 0;
    if(n == 0) return COATI_GPU_OK;
    // reset results (status from host-side validation)
    // reset the per-pair records: all-zero, then the (rare) pairs rejected by host-side validation
    CU_TRY(ctx, cudaMemsetAsync(bt->d_results.p, 0, n * sizeof(PairResult), s));
    if(!bt->rejected_known) {
        for(size_t p = 0; p < bt->npairs; ++p)
            if(bt->host_status[p] != COATI_GPU_OK) {
                PairResult r{};
                r.status = bt->host_status[p];
                bt->h_init.push_back(r);
                bt->rejected.push_back(static_cast<uint32_t>(p));
            }
        bt->rejected_known = true;
    }
    for(size_t x = 0; x < bt->rejected.size(); ++x)
        CU_TRY(ctx, cudaMemcpyAsync(bt->d_results.p + bt->rejected[x], &bt->h_init[x], sizeof(PairResult),
                                    cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemsetAsync(bt->d_counters.p, 0, bt->d_counters.n * sizeof(unsigned int), s));
    {
        const uint32_t warps_per_block = 2;  // small CTAs: co-resident with the other lane's fill
        if(bt->raw) {
            // the counters buffer has one spare word: used as the "any ambiguity code" flag
            unsigned int* flag = bt->d_counters.p + bt->runs.size();
            encode_pairs_kernel<<<(n + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                                  s>>>(bt->d_pairs.p, n, bt->d_anc.p, bt->d_des.p, bt->d_a.p, bt->d_b.p,
                                       bt->d_results.p, flag);
            trace_mark("  run: encode enqueued", n);
            bt->nc = 16;  // decided on the device: see the fill launches below
        } else
            validate_symbols_kernel<<<(n + warps_per_block - 1) / warps_per_block,
                                      warps_per_block * 32, 0, s>>>(bt->d_pairs.p, n, bt->d_a.p,
                                                                     bt->d_b.p, bt->d_results.p);
        ++bt->launches;
    }
    // events: 2 per run (fill start / end), then 5 per chunk (traceback start / end, expand start / end,
    // inputs ready)
    const size_t n_events = 2 * bt->runs.size() + 5 * bt->chunks.size();
    cudaStream_t sf = bt->fill_stream ? bt->fill_stream : s;  // fills; fenced against s when distinct
    if(bt->events.size() != n_events) {
        for(cudaEvent_t e : bt->events) cudaEventDestroy(e);
        bt->events.assign(n_events, nullptr);
        for(cudaEvent_t& e : bt->events) CU_TRY(ctx, cudaEventCreate(&e));
    }
    // Per chunk (one direction buffer): every fill first, then the tracebacks -- serial latency-bound
    // walks, all in flight together -- then one expansion.  The small kernels are sized to fit beside
    // the resident fill CTAs of the other pipeline lane (see traceback_chunk_kernel).
    size_t ri = 0;
    for(size_t ci = 0; ci < bt->chunks.size(); ++ci) {
        const Chunk& ch = bt->chunks[ci];
        const size_t ri0 = ri;
        bool any_inter = false;
        cudaEvent_t* cev = &bt->events[2 * bt->runs.size() + 5 * ci];
        if(sf != s) {  // inputs / the direction buffer are ready: encode done, previous chunk walked
            cudaEvent_t ready = cev[4];
            if(ci == 0) cudaEventRecord(ready, s);
            else ready = bt->events[2 * bt->runs.size() + 5 * (ci - 1) + 1];
            CU_TRY(ctx, cudaStreamWaitEvent(sf, ready, 0));
        }
        for(; ri < bt->runs.size() && bt->runs[ri].chunk == ci; ++ri) {
            const Run& r = bt->runs[ri];
            const uint32_t cnt = r.last - r.first;
            cudaEvent_t* ev = &bt->events[2 * ri];
            cudaEventRecord(ev[0], sf);
            if(r.cfg == 0) {
                const uint32_t grid = std::min(cnt, bt->ring_ctas);
                viterbi_generic_kernel<<<grid, 128, 0, sf>>>(bt->d_pairs.p, r.first, r.last,
                                                            bt->d_counters.p + ri, bt->d_a.p, bt->d_b.p,
                                                            ctx->d_table, ctx->gap, bt->d_ring.p,
                                                            bt->ring_stride, bt->d_dirs.p,
                                                            bt->d_results.p);
                any_inter = true;
            } else {
                const PipeCfg* pc16 = find_cfg(ctx->gap.k, r.cfg, 16, ctx->pipe_scalar);
                const PipeCfg* pc4 = find_cfg(ctx->gap.k, r.cfg, 4, ctx->pipe_scalar);
                if(!pc16) return COATI_GPU_E_ARG;
                // Raw-sequence batches learn on the device whether any descendant carries an ambiguity code
                // (encode_pairs_kernel sets a flag word): both column variants are launched and the one
                // that does not apply returns at once, so the host never synchronises on the flag.
                const PipeCfg* todo[2] = {bt->nc == 4 ? pc4 : pc16, nullptr};
                const unsigned int* nc_flag = nullptr;
                if(bt->raw && pc4 != pc16 && pc16->fn1) {
                    todo[0] = pc16, todo[1] = pc4;
                    nc_flag = bt->d_counters.p + bt->runs.size();
                }
                const PairDesc& d = bt->descs[r.first];
                if(pc16->wave)  // NaN sentinel in every boundary entry: the data is its own ready flag
                    CU_TRY(ctx, cudaMemsetAsync(bt->d_bnd.p, 0xff, wave_bnd_f4(d.la, d.lb, pc16->R) * sizeof(float4), sf));
                for(const PipeCfg* pc : todo) {
                    if(!pc) continue;
                    const uint32_t cap = (uint32_t)ctx->prop.multiProcessorCount * ctx->ctas_per_sm[pc - g_pipe_cfgs];
                    if(pc->wave) {
                        const uint32_t nb = (d.la + 32 * pc->R - 1) / (32 * pc->R);
                        // CTAs of four warps, one band per scheduler (CTAs of one or two warps measured: no gain)
                        const uint32_t grid = std::min((nb + PIPE_WARPS - 1) / PIPE_WARPS, cap);
                        pc->fn1<<<grid, PIPE_WARPS * 32, pc->smem, sf>>>(
                            bt->d_pairs.p, r.first, r.last, bt->d_counters.p + ri, bt->d_a.p, bt->d_b.p,
                            ctx->d_table, ctx->gap, bt->d_bnd.p, (d.lb + 4) / 2, bt->d_dirs.p,
                            bt->d_results.p, nc_flag);
                    } else {
                        const uint32_t want = (cnt + PIPE_WARPS - 1) / PIPE_WARPS;
                        const uint32_t grid = std::min(want, std::min(bt->bnd_ctas, cap));
                        if(pc->fn1)
                            pc->fn1<<<grid, PIPE_WARPS * 32, pc->smem, sf>>>(
                                bt->d_pairs.p, r.first, r.last, bt->d_counters.p + ri, bt->d_a.p, bt->d_b.p,
                                ctx->d_table, ctx->gap, bt->d_bnd.p, bt->bnd_stride, bt->d_dirs.p,
                                bt->d_results.p, nc_flag);
                        else
                            pc->fn<<<grid, PIPE_WARPS * 32, pc->smem, sf>>>(
                                bt->d_pairs.p, r.first, r.last, bt->d_counters.p + ri, bt->d_a.p, bt->d_b.p,
                                ctx->d_table, ctx->gap, bt->d_bnd.p, bt->bnd_stride, bt->d_dirs.p,
                                bt->d_results.p);
                        any_inter = true;
                    }
                    if(pc != todo[0]) ++bt->launches;
                }
            }
            cudaEventRecord(ev[1], sf);
            ++bt->launches;
        }
        if(sf != s && ri > ri0) CU_TRY(ctx, cudaStreamWaitEvent(s, bt->events[2 * (ri - 1) + 1], 0));
        cudaEventRecord(cev[0], s);
        for(size_t rj = ri0; rj < ri; ++rj) {  // long pairs: one warp per pair, with read-ahead
            const Run& r = bt->runs[rj];
            if(!(r.cfg & CFG_WAVE)) continue;
            const uint32_t cnt = r.last - r.first;
#define COATI_TB(RR)                                                                                   \
    if(ctx->tb_serial)                                                                                 \
        traceback_kernel<PipeLayoutR<RR>, true><<<cnt, 32, 0, s>>>(                                    \
            bt->d_pairs.p, r.first, r.last, bt->d_dirs.p, ctx->gap, bt->d_out_b.p, bt->d_results.p);   \
    else                                                                                               \
        traceback_burst_kernel<PipeLayoutR<RR>><<<1, 32, 0, s>>>(                                      \
            bt->d_pairs.p, r.first, bt->d_dirs.p, ctx->gap, bt->d_long_ops.p + r.ops_off,              \
            bt->d_long_ck.p + r.ck_off, bt->d_results.p);
            switch(r.cfg & 0xffu) {
            case 2: COATI_TB(2) break;
            case 4: COATI_TB(4) break;
            case 8: COATI_TB(8) break;
            case 10: COATI_TB(10) break;
            default: return COATI_GPU_E_ARG;
            }
#undef COATI_TB
            ++bt->launches;
            if(!ctx->tb_serial) {  // rows of the long pair, one warp per >= 2048-column segment
                const PairDesc& d = bt->descs[r.first];
                const uint32_t nseg = long_ck_capacity(d.la, d.lb);
                expand_long_kernel<<<(nseg + 1) / 2, 64, 0, s>>>(
                    bt->d_pairs.p, r.first, bt->d_long_ops.p + r.ops_off, bt->d_long_ck.p + r.ck_off,
                    bt->d_anc.p, bt->d_des.p, bt->d_out_a.p, bt->d_out_b.p, bt->d_results.p,
                    ctx->gap.stop_gap);
                ++bt->launches;
            }
        }
        const uint32_t ccnt = ch.last - ch.first;
        if(any_inter) {
            if(ch.long_count) {
                traceback_burst_list_kernel<<<(ch.long_count + 1) / 2, 64, 0, s>>>(
                    bt->d_pairs.p, bt->d_long_list.p + ch.long_first, ch.long_count, bt->d_dirs.p, ctx->gap,
                    bt->d_out_b.p, bt->d_results.p);
                ++bt->launches;
            }
            traceback_chunk_kernel<<<(ccnt + 63) / 64, 64, 0, s>>>(bt->d_pairs.p, ch.first, ch.last,
                                                                   bt->d_dirs.p, ctx->gap, bt->d_out_b.p,
                                                                   bt->d_results.p, ctx->tb_serial ? 0u : 1u);
            ++bt->launches;
        }
        cudaEventRecord(cev[1], s);
        cudaEventRecord(cev[2], s);
        expand_rows_kernel<<<(ccnt + 1) / 2, 64, 0, s>>>(bt->d_pairs.p, ch.first, ch.last, bt->d_anc.p,
                                                         bt->d_des.p, bt->d_out_a.p, bt->d_out_b.p,
                                                         bt->d_results.p, ctx->gap.stop_gap,
                                                         ctx->tb_serial ? 0u : 1u);
        if(bt->h_out_a) {  // the used bytes of every row of the chunk, straight into the caller's arenas
            const uint32_t want = (ccnt + R2H_THREADS / 32 - 1) / (R2H_THREADS / 32);
            rows_to_host_kernel<<<std::min(want, ctx->rows_ctas), R2H_THREADS, 0, s>>>(
                bt->d_pairs.p, ch.first, ch.last, bt->d_out_a.p, bt->d_out_b.p, bt->h_out_a, bt->h_out_b,
                bt->d_results.p);
            ++bt->launches;
        }
        cudaEventRecord(cev[3], s);
        ++bt->launches;
    }
    ctx->launches += bt->launches;
    CU_TRY(ctx, cudaGetLastError());
    return COATI_GPU_OK;
}

static int batch_download_async(coati_gpu_batch* bt, char* out_a, char* out_b) {
    coati_gpu_ctx* ctx = bt->ctx;
    cudaStream_t s = bt->stream;
    if(bt->npairs == 0) return COATI_GPU_OK;
    if(out_a && !bt->h_out_a) {  // (rows written by rows_to_host_kernel are already there)
        CU_TRY(ctx, cudaMemcpyAsync(out_a, bt->d_out_a.p, bt->out_total, cudaMemcpyDeviceToHost, s));
        ctx->d2h_bytes += bt->out_total;
    }
    if(out_b && !bt->h_out_b) {
        CU_TRY(ctx, cudaMemcpyAsync(out_b, bt->d_out_b.p, bt->out_total, cudaMemcpyDeviceToHost, s));
        ctx->d2h_bytes += bt->out_total;
    }
    ctx->d2h_bytes += bt->npairs * sizeof(PairResult);
    CU_TRY(ctx, cudaMemcpyAsync(bt->h_results, bt->d_results.p, bt->npairs * sizeof(PairResult),
                                cudaMemcpyDeviceToHost, s));
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_batch_download(coati_gpu_batch* bt, char* out_a, char* out_b,
                                        uint64_t* out_len, float* score, int32_t* status) {
    if(!bt) return COATI_GPU_E_ARG;
    coati_gpu_ctx* ctx = bt->ctx;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = bt->stream;
    if(bt->npairs == 0) return COATI_GPU_OK;
    if(int rc = batch_download_async(bt, out_a, out_b)) return rc;
    CU_TRY(ctx, cudaStreamSynchronize(s));
    for(size_t p = 0; p < bt->npairs; ++p) {
        const PairResult& r = bt->h_results[p];
        if(out_len) out_len[p] = r.len;
        if(score) score[p] = r.score;
        if(status) status[p] = r.status;
    }
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_batch_stats(coati_gpu_batch* bt, uint64_t* cells, uint64_t* dir_bytes,
                                     uint64_t* launches, uint64_t* chunks) {
    if(!bt) return COATI_GPU_E_ARG;
    if(cells) *cells = bt->cells;
    if(dir_bytes) *dir_bytes = bt->dir_bytes;
    if(launches) *launches = bt->launches;
    if(chunks) *chunks = bt->chunks.size();
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_batch_timing(coati_gpu_batch* bt, double* fill_ms, double* traceback_ms,
                                      double* compact_ms, uint64_t* fill_launches) {
    if(!bt) return COATI_GPU_E_ARG;
    coati_gpu_ctx* ctx = bt->ctx;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(bt->stream));
    double f = 0, t = 0, c = 0;
    const size_t nr = bt->runs.size();
    if(bt->events.size() == 2 * nr + 5 * bt->chunks.size()) {
        float ms = 0;
        for(size_t ri = 0; ri < nr; ++ri) {
            CU_TRY(ctx, cudaEventElapsedTime(&ms, bt->events[2 * ri], bt->events[2 * ri + 1]));
            f += ms;
        }
        for(size_t ci = 0; ci < bt->chunks.size(); ++ci) {
            cudaEvent_t* cev = &bt->events[2 * nr + 5 * ci];
            CU_TRY(ctx, cudaEventElapsedTime(&ms, cev[0], cev[1]));
            t += ms;
            CU_TRY(ctx, cudaEventElapsedTime(&ms, cev[2], cev[3]));
            c += ms;
        }
    }
    if(fill_ms) *fill_ms = f;
    if(traceback_ms) *traceback_ms = t;
    if(compact_ms) *compact_ms = c;
    if(fill_launches) *fill_launches = nr;
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_batch_device_buffers(coati_gpu_batch* bt, void** out_a, void** out_b,
                                              uint64_t* out_bytes, void** results,
                                              uint64_t* result_bytes) {
    if(!bt) return COATI_GPU_E_ARG;
    if(out_a) *out_a = bt->d_out_a.p;
    if(out_b) *out_b = bt->d_out_b.p;
    if(out_bytes) *out_bytes = bt->out_total;
    if(results) *results = bt->d_results.p;
    if(result_bytes) *result_bytes = bt->npairs * sizeof(PairResult);
    return COATI_GPU_OK;
}

extern "C" void coati_gpu_batch_destroy(coati_gpu_batch* bt) {
    if(!bt) return;
    cudaSetDevice(bt->ctx->device);
    cudaStreamSynchronize(bt->stream);
    delete bt;
}

// process_marginal's length checks (before trimming, utils.cc:819-837) and trim_end_stops
// (utils.cc:945-967 via cod_int) for the raw pairs [p0, p1): bit 7 = bad length, bit 0 / 1 = the
// ancestor / descendant ends with a stop codon; raw[p - p0] for pair p.
static void scan_raw_pairs(uint32_t k, const char* anc_all, const uint64_t* anc_off, const char* des_all,
                           const uint64_t* des_off, size_t p0, size_t p1, uint8_t* raw) {
    auto nuc = [](char ch) -> int {
        switch(ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return -1;
        }
    };
    auto ends_with_stop = [&](const char* s, uint64_t len) {
        if(len < 3) return false;
        const int n0 = nuc(s[len - 3]), n1 = nuc(s[len - 2]), n2 = nuc(s[len - 1]);
        if((n0 | n1 | n2) < 0) return false;
        const int cod = (n0 << 4) | (n1 << 2) | n2;
        return cod == 48 || cod == 50 || cod == 56;
    };
    constexpr size_t AHEAD = 16;  // the tails sit hundreds of bytes apart: fetch them ahead of the test
    for(size_t p = p0; p < p1; ++p) {
        if(p + AHEAD < p1) {
            __builtin_prefetch(anc_all + anc_off[p + AHEAD + 1] - 3);
            __builtin_prefetch(des_all + des_off[p + AHEAD + 1] - 3);
        }
        const uint64_t la = anc_off[p + 1] - anc_off[p], lb = des_off[p + 1] - des_off[p];
        uint8_t f = 0;
        if(la % 3 != 0 || la % k != 0 || lb % k != 0) f |= 0x80;
        if(ends_with_stop(anc_all + anc_off[p], la)) f |= 1;
        if(ends_with_stop(des_all + des_off[p], lb)) f |= 2;
        raw[p - p0] = f;
    }
}

// ---------------------------------------------------------------------------------------------
// One CSR batch as a queue of contiguous sub-batches.  A context's worker takes the next range whenever one
// of its three pipeline lanes is free; several workers (one per device) may share one queue, which is the
// whole multi-GPU scheduler: pairs are independent, ranges are handed out heaviest first
// (longest-processing-time order on the sum of La * Lb), and a device that finishes early simply takes more.
struct BatchArgs {
    size_t npairs;
    const uint8_t* a_all;
    const uint64_t* a_off;
    const uint8_t* b_all;
    const uint64_t* b_off;
    const char* anc_all;
    const char* des_all;
    char* out_a;
    char* out_b;
    uint64_t* out_len;
    float* score;
    int32_t* status;
    bool raw_mode;
    const uint32_t* model;
    bool shared_link = false;  // the call is one share of a batch that other devices work on at the same time
};

#ifndef COATI_GPU_TAIL_SPLIT
#define COATI_GPU_TAIL_SPLIT 1
#endif
struct RangeQueue {
    std::vector<std::pair<size_t, size_t>> ranges;  // [first, last) in pair indices
    std::atomic<size_t> next{0};
    std::atomic<bool> abort{false};
    uint64_t max_sym = 0;    // symbols (a + b) of the biggest range
    size_t max_pairs = 0;
    bool pop(size_t& p0, size_t& p1) {
        if(abort.load(std::memory_order_relaxed)) return false;
        const size_t j = next.fetch_add(1, std::memory_order_relaxed);
        if(j >= ranges.size()) return false;
        p0 = ranges[j].first, p1 = ranges[j].second;
        return true;
    }
    bool drained() const { return next.load(std::memory_order_relaxed) >= ranges.size(); }
    void add(size_t p0, size_t p1, const uint64_t* a_off, const uint64_t* b_off) {
        ranges.emplace_back(p0, p1);
        max_sym = std::max<uint64_t>(max_sym, (a_off[p1] - a_off[p0]) + (b_off[p1] - b_off[p0]));
        max_pairs = std::max(max_pairs, p1 - p0);
    }
};

// lattice cells of pairs [p0, p1): the cost the ranges are balanced on
static double range_cells(const uint64_t* a_off, const uint64_t* b_off, size_t p0, size_t p1) {
    double c = 0;
    for(size_t p = p0; p < p1; ++p) c += (double)(a_off[p + 1] - a_off[p]) * (double)(b_off[p + 1] - b_off[p]);
    return c;
}

// Cut [0, npairs) into contiguous chunks of equal WEIGHT (sum of La * Lb), returned heaviest first.
// How heavy: a fill cannot end before its longest pair does -- one warp sweeps a 2400 x 2400 lattice in about
// 12 ms while it shares its scheduler with three others -- so a chunk should hold at least that much work for
// the whole GPU (1.3e10 cells at ~1.06 TCUPS), or its fill ends in a tail of a few busy warps; and about four
// chunks per worker keep the pipeline of H2D / kernels / D2H and the balance between workers.  The count is
// rounded to the nearest multiple of `workers`, and no chunk gets more than 2^17 pairs (a batch sorted by length
// would otherwise put millions of short pairs into one chunk).
constexpr double CHUNK_MIN_CELLS = 1.3e10, CHUNK_MAX_CELLS = 2.6e10;
constexpr size_t CHUNK_MAX_PAIRS = size_t(1) << 17;
static void plan_chunks(size_t npairs, const uint64_t* a_off, const uint64_t* b_off, size_t workers,
                        std::vector<std::pair<size_t, size_t>>& out, std::vector<double>& cost) {
    workers = std::max<size_t>(1, workers);
    const double total = range_cells(a_off, b_off, 0, npairs);
    const double target = std::min(CHUNK_MAX_CELLS, std::max(CHUNK_MIN_CELLS, total / (4.0 * (double)workers)));
    size_t n = std::max<size_t>(1, (size_t)std::llround(total / target));
    if(n > 1 || workers > 1) n = std::max(workers, (n + workers / 2) / workers * workers);  // nearest multiple
    n = std::min(n, std::max<size_t>(npairs, 1));
    std::vector<std::pair<size_t, size_t>> r;
    std::vector<double> c;
    size_t p0 = 0;
    double acc = 0, done = 0;
    for(size_t p = 0; p < npairs; ++p) {
        acc += (double)(a_off[p + 1] - a_off[p]) * (double)(b_off[p + 1] - b_off[p]);
        // the k-th cut sits where the running weight passes k / n of the total
        const bool heavy = done + acc >= total * (double)(r.size() + 1) / (double)n;
        if(p + 1 == npairs || heavy || p + 1 - p0 >= CHUNK_MAX_PAIRS) {
            r.emplace_back(p0, p + 1);
            c.push_back(acc);
            done += acc, acc = 0, p0 = p + 1;
        }
    }
    if(r.empty()) r.emplace_back(0, npairs), c.push_back(0.0);
    std::vector<size_t> order(r.size());
    for(size_t j = 0; j < r.size(); ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return c[x] > c[y]; });
    out.clear(), cost.clear();
    for(size_t j : order) out.push_back(r[j]), cost.push_back(c[j]);
}

// The address the current device may use for host memory [p, p + bytes), or nullptr when the range is not
// page-locked and mapped as a whole (pageable memory, or two registrations side by side).
static char* device_view(char* p, uint64_t bytes) {
    if(!p || !bytes) return nullptr;
    cudaPointerAttributes first{}, last{};
    void* dev = nullptr;
    if(cudaPointerGetAttributes(&first, p) != cudaSuccess ||
       cudaPointerGetAttributes(&last, p + bytes - 1) != cudaSuccess ||
       first.type != cudaMemoryTypeHost || last.type != cudaMemoryTypeHost || !first.devicePointer ||
       !last.devicePointer ||
       static_cast<char*>(last.devicePointer) - static_cast<char*>(first.devicePointer) != (ptrdiff_t)(bytes - 1) ||
       cudaHostGetDevicePointer(&dev, p, 0) != cudaSuccess || dev != first.devicePointer) {
        cudaGetLastError();
        return nullptr;
    }
    return static_cast<char*>(dev);
}

static int viterbi_batch_pipeline(coati_gpu_ctx* ctx, const BatchArgs& A, RangeQueue& q, bool pipelined);

static int viterbi_batch_impl(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all,
                              const uint64_t* a_off, const uint8_t* b_all, const uint64_t* b_off,
                              const char* anc_all, const char* des_all, char* out_a, char* out_b,
                              uint64_t* out_len, float* score, int32_t* status, bool raw_mode,
                              const uint32_t* model = nullptr) {
    if(!ctx || (npairs && (!a_off || !b_off))) return COATI_GPU_E_ARG;
    // Large batches are cut into sub-batches that rotate over the lanes of the context, so that the
    // host-side planning, the H2D copy and the D2H copy of one sub-batch overlap the kernels of its
    // neighbours.
    size_t nsub = 1;
    if(npairs >= 65536) {  // same weight rule as plan_chunks, in input order
        const double total = range_cells(a_off, b_off, 0, npairs);
        nsub = (size_t)std::llround(total / std::min(CHUNK_MAX_CELLS, std::max(CHUNK_MIN_CELLS, total / 4.0)));
        nsub = std::min<size_t>(std::max<size_t>(nsub, 1), npairs / 8192);
    }
    if(const char* env = std::getenv("COATI_GPU_NSUB")) nsub = std::max(1, std::atoi(env));  // tuning
    if(ctx->dir_budget != 0) nsub = 1;  // an explicit direction budget (tests) keeps the simple path
    const BatchArgs A{npairs, a_all, a_off, b_all, b_off, anc_all, des_all, out_a, out_b, out_len, score, status,
                      raw_mode, model};
    const uint64_t zero_off[1] = {0};
    auto run = [&](size_t parts) {
        RangeQueue q;
        for(size_t j = 0; j < parts; ++j)  // in input order: a single device gains nothing from reordering
            q.add(npairs * j / parts, npairs * (j + 1) / parts, npairs ? a_off : zero_off, npairs ? b_off : zero_off);
        return viterbi_batch_pipeline(ctx, A, q, parts > 1);
    };
    int rc = run(nsub);
    if(rc == COATI_GPU_E_NOMEM) {
        // a lane owns its share (1 / NLANE) of the memory: a pair too big for that may still fit the whole device; and
        // idle pool blocks of earlier calls are given back before the plan is redone from the memory
        // that is free now
        CU_TRY(ctx, cudaSetDevice(ctx->device));
        CU_TRY(ctx, cudaDeviceSynchronize());
        ctx->pool.trim();
        rc = run(1);
    }
    return rc;
}

static int viterbi_batch_pipeline(coati_gpu_ctx* ctx, const BatchArgs& A, RangeQueue& q, bool pipelined) {
    const size_t npairs = A.npairs;
    const uint64_t *a_off = A.a_off, *b_off = A.b_off;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    constexpr int NSLOT = coati_gpu_ctx::NLANE;  // sub-batches in flight, one lane per slot
    coati_gpu_batch* bt[NSLOT] = {};
    size_t first[NSLOT] = {};
    // raw pairs: the per-pair scan (length checks, end stops) runs range by range, inside the pipeline
    std::vector<uint8_t> rawbuf;
    // direction-stream budget of a lane, from the memory free now (the GPU is idle: cheap call)
    uint64_t lane_budget = 0;
    if(pipelined) {
        size_t free_b = 0;
        CU_TRY(ctx, ctx->free_bytes(&free_b));
        free_b += ctx->pool.idle_bytes();
        const uint64_t fixed_sub = (4 * q.max_sym + q.max_pairs * (2 + sizeof(PairDesc) + sizeof(PairResult))) * 5 / 4 +
                                   (256ull << 20);
        lane_budget = free_b > NSLOT * fixed_sub
                          ? static_cast<uint64_t>((free_b - NSLOT * fixed_sub) * 0.85 / NSLOT) : (1ull << 20);
    }
    if(pipelined) {
        // one pinned landing zone per slot, each big enough for the biggest range: taken and given back here so
        // that no sub-batch has to allocate pinned memory (cudaMallocHost waits for running kernels) mid-pipeline
        void* zone[NSLOT];
        for(void*& z : zone) z = ctx->hpool.take((q.max_pairs + 1) * sizeof(PairResult));
        for(void* z : zone) ctx->hpool.give(z);
    }
    // Row arenas the device can address (page-locked and mapped, first to last byte): a kernel writes the rows
    // there, used bytes only, instead of a D2H copy of the padded slots (rows_to_host_kernel)
    char *dev_out_a = nullptr, *dev_out_b = nullptr;
    if((ctx->rows_direct > 0 || (ctx->rows_direct < 0 && A.shared_link)) && npairs && A.out_a && A.out_b) {
        const uint64_t total = a_off[npairs] + b_off[npairs] + npairs;
        dev_out_a = device_view(A.out_a, total);
        dev_out_b = dev_out_a ? device_view(A.out_b, total) : nullptr;
        if(!dev_out_b) dev_out_a = nullptr;
    }
    const bool trace = trace_on();
    auto mark = [&](const char* what, size_t j) { trace_mark(what, j); };
    auto finish = [&](int slot) -> int {
        coati_gpu_batch* b = bt[slot];
        if(!b) return COATI_GPU_OK;
        bt[slot] = nullptr;
        int rc = COATI_GPU_OK;
        mark("wait begin", first[slot]);
        if(cudaStreamSynchronize(b->stream) != cudaSuccess) {
            ctx->last_error = cudaGetErrorString(cudaGetLastError());
            rc = COATI_GPU_E_CUDA;
        }
        uint64_t row_bytes = 0;
        for(size_t p = 0; rc == COATI_GPU_OK && p < b->npairs; ++p) {
            const PairResult& r = b->h_results[p];
            if(A.out_len) A.out_len[first[slot] + p] = r.len;
            if(A.score) A.score[first[slot] + p] = r.score;
            if(A.status) A.status[first[slot] + p] = r.status;
            row_bytes += (r.status == 0 ? r.len : 0) + 1;
        }
        if(b->h_out_a) ctx->d2h_bytes += 2 * row_bytes;  // what rows_to_host_kernel wrote over the link
        if(trace && rc == COATI_GPU_OK) {
            double fill = 0, tb = 0, ex = 0;
            coati_gpu_batch_timing(b, &fill, &tb, &ex, nullptr);
            std::fprintf(stderr, "[coati_gpu trace]              sub@%zu device ms: fill %.2f traceback %.2f expand %.2f\n",
                         first[slot], fill, tb, ex);
        }
        mark("wait end", first[slot]);
        delete b;  // stream already drained
        return rc;
    };
    int rc = COATI_GPU_OK;
    size_t p0 = 0, p1 = 0;
    // The first range of a call is split: a small head is planned and on the GPU within a millisecond, and the
    // planning of the rest runs beside its fill.
    constexpr size_t HEAD_PAIRS = 8192;
    size_t rest0 = 0, rest1 = 0;
    auto next_range = [&](size_t j) {
        if(rest1 > rest0) {
            p0 = rest0, p1 = rest1, rest1 = rest0;
            return true;
        }
        if(!q.pop(p0, p1)) return false;
        if(j == 0 && pipelined && p1 - p0 > 2 * HEAD_PAIRS) rest0 = p0 + HEAD_PAIRS, rest1 = p1, p1 = rest0;
        // ... and so is the last one: what is left to do after the last fill -- the traceback, the expansion and
        // the D2H copy of its rows -- then belongs to a third of a chunk (its fill shares the GPU with the fill
        // of the other two thirds)
        else if(COATI_GPU_TAIL_SPLIT && pipelined && q.drained() && p1 - p0 > 3 * HEAD_PAIRS)
            rest0 = p0 + (p1 - p0) / 3 * 2, rest1 = p1, p1 = rest0;
        return true;
    };
    for(size_t j = 0; rc == COATI_GPU_OK && next_range(j); ++j) {
        const int slot = (int)(j % NSLOT);
        rc = finish(slot);
        if(rc != COATI_GPU_OK) break;
        first[slot] = p0;
        mark("plan begin", p0);
        const uint8_t* raw = nullptr;
        if(A.raw_mode) {
            try {
                rawbuf.resize(p1 - p0);
            } catch(const std::bad_alloc&) {
                rc = COATI_GPU_E_NOMEM;
                break;
            }
            scan_raw_pairs(ctx->gap.k, A.anc_all, a_off, A.des_all, b_off, p0, p1, rawbuf.data());
            raw = rawbuf.data();
        }
        rc = batch_create_on(ctx, pipelined ? ctx->lane_hi[slot] : ctx->stream,
                             pipelined ? ((j & 1) ? ctx->fill_stream2 : ctx->fill_stream) : ctx->stream, lane_budget, p1 - p0,
                             a_off + p0, b_off + p0, &bt[slot], raw, A.model ? A.model + p0 : nullptr);
        if(rc != COATI_GPU_OK) break;
        mark("plan end", p0);
        const uint64_t ao = npairs ? a_off[p0] : 0, bo = npairs ? b_off[p0] : 0;
        if(dev_out_a) {
            // not for a sub-batch with wavefront pairs: one warp per pair would send a 100 kB row 512 bytes at a
            // time, and the rows of such pairs are small beside their lattices anyway -- the copy serves them
            bool wave = false;
            for(const Run& r : bt[slot]->runs) wave = wave || (r.cfg & CFG_WAVE);
            if(!wave) bt[slot]->h_out_a = dev_out_a + ao + bo + p0, bt[slot]->h_out_b = dev_out_b + ao + bo + p0;
        }
        rc = coati_gpu_batch_upload(bt[slot], A.a_all ? A.a_all + ao : nullptr, A.b_all ? A.b_all + bo : nullptr,
                                    A.anc_all ? A.anc_all + ao : nullptr, A.des_all ? A.des_all + bo : nullptr);
        mark("upload enqueued", p0);
        if(rc == COATI_GPU_OK) rc = coati_gpu_batch_run(bt[slot]);
        mark("run enqueued", p0);
        if(rc == COATI_GPU_OK)
            rc = batch_download_async(bt[slot], A.out_a ? A.out_a + ao + bo + p0 : nullptr,
                                      A.out_b ? A.out_b + ao + bo + p0 : nullptr);
        mark("download enqueued", p0);
    }
    if(rc != COATI_GPU_OK) q.abort.store(true);
    for(int sl = 0; sl < NSLOT; ++sl) {
        const int r2 = finish(sl);
        if(rc == COATI_GPU_OK) rc = r2;
    }
    return rc;
}

extern "C" int coati_gpu_viterbi_batch(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all,
                                       const uint64_t* a_off, const uint8_t* b_all,
                                       const uint64_t* b_off, const char* anc_all,
                                       const char* des_all, char* out_a, char* out_b,
                                       uint64_t* out_len, float* score, int32_t* status) {
    return viterbi_batch_impl(ctx, npairs, a_all, a_off, b_all, b_off, anc_all, des_all, out_a, out_b,
                              out_len, score, status, false);
}

// The leaf batch of the msa driver (align_msa.cc:285-318): every pair names its own substitution model
// (one table per branch length), gap parameters shared.
extern "C" int coati_gpu_viterbi_batch_models(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all,
                                              const uint64_t* a_off, const uint8_t* b_all,
                                              const uint64_t* b_off, const char* anc_all,
                                              const char* des_all, const uint32_t* model_idx, char* out_a,
                                              char* out_b, uint64_t* out_len, float* score,
                                              int32_t* status) {
    if(npairs && !model_idx) return COATI_GPU_E_ARG;
    return viterbi_batch_impl(ctx, npairs, a_all, a_off, b_all, b_off, anc_all, des_all, out_a, out_b,
                              out_len, score, status, false, model_idx);
}

// marg_alignment (align_marginal.cc:44-88) for a batch of raw pairs: process_marginal's length checks
// (before trimming, utils.cc:819-837), trim_end_stops, marginal_seq_encoding (on the device),
// viterbi_mem + traceback_viterbi, restore_end_stops.
extern "C" int coati_gpu_alignpair_batch(coati_gpu_ctx* ctx, size_t npairs, const char* anc_all,
                                         const uint64_t* anc_off, const char* des_all,
                                         const uint64_t* des_off, char* out_a, char* out_b,
                                         uint64_t* out_len, float* score, int32_t* status) {
    if(!ctx || (npairs && (!anc_off || !des_off || !anc_all || !des_all))) return COATI_GPU_E_ARG;
    if(!ctx->model_set) return COATI_GPU_E_ARG;
    return viterbi_batch_impl(ctx, npairs, nullptr, anc_off, nullptr, des_off, anc_all, des_all, out_a, out_b,
                              out_len, score, status, true);
}

// ---- one batch over several devices ------------------------------------------------------------------
// north_star (4) / SURVEY 8(e): independent pairs sharded over the GPUs of one box by length-binned work
// queues.  The batch is cut into contiguous chunks, ordered heaviest first, and one host thread per context
// runs the four-lane pipeline of the single-device call on whatever chunk it pops next; inside a chunk the
// planner bins the pairs by kernel configuration and lattice size as always.  Results land in the caller's
// arenas in input order -- no collective, no second copy.
extern "C" int coati_gpu_multi_alignpair_batch(coati_gpu_ctx* const* ctxs, int n_ctx, size_t npairs,
                                               const char* anc_all, const uint64_t* anc_off,
                                               const char* des_all, const uint64_t* des_off, char* out_a,
                                               char* out_b, uint64_t* out_len, float* score, int32_t* status) {
    if(!ctxs || n_ctx < 1 || (npairs && (!anc_off || !des_off || !anc_all || !des_all))) return COATI_GPU_E_ARG;
    for(int i = 0; i < n_ctx; ++i) {
        if(!ctxs[i] || !ctxs[i]->model_set) return COATI_GPU_E_ARG;
        // one model for the whole batch: same gap constants everywhere (the tables are the caller's word)
        if(std::memcmp(&ctxs[i]->gap, &ctxs[0]->gap, sizeof(GapConsts)) != 0) return COATI_GPU_E_ARG;
        for(int j = 0; j < i; ++j)
            if(ctxs[j] == ctxs[i]) return COATI_GPU_E_ARG;
    }
    if(n_ctx == 1 || npairs < 4096)
        return coati_gpu_alignpair_batch(ctxs[0], npairs, anc_all, anc_off, des_all, des_off, out_a, out_b, out_len,
                                         score, status);
    const BatchArgs A{npairs, nullptr, anc_off, nullptr, des_off, anc_all, des_all, out_a, out_b, out_len, score,
                      status, true, nullptr, true};
    RangeQueue q;
    try {
        std::vector<std::pair<size_t, size_t>> chunks;
        std::vector<double> cost;
        plan_chunks(npairs, anc_off, des_off, (size_t)n_ctx, chunks, cost);
        for(const auto& c : chunks) q.add(c.first, c.second, anc_off, des_off);
    } catch(const std::bad_alloc&) {
        return COATI_GPU_E_NOMEM;
    }
    std::vector<int> rcs((size_t)n_ctx, COATI_GPU_OK);
    std::vector<std::thread> workers;
    for(int i = 1; i < n_ctx; ++i)
        workers.emplace_back([&, i] { rcs[(size_t)i] = viterbi_batch_pipeline(ctxs[i], A, q, true); });
    rcs[0] = viterbi_batch_pipeline(ctxs[0], A, q, true);  // the calling thread drives the first device
    for(std::thread& t : workers) t.join();
    for(int rc : rcs)
        if(rc != COATI_GPU_OK) return rc;
    return COATI_GPU_OK;
}

// The same plan for callers that run one PROCESS per device (bench.py under torchrun): contiguous chunks,
// heaviest first, each given to the least-loaded shard (greedy longest-processing-time); every process
// computes the identical plan from the offsets alone.  Returns the number of ranges written (<= max_ranges;
// 0 if the arrays are too small).
extern "C" size_t coati_gpu_plan_shards(size_t npairs, const uint64_t* a_off, const uint64_t* b_off,
                                        uint32_t n_shards, size_t max_ranges, uint64_t* range_first,
                                        uint64_t* range_last, uint32_t* range_shard) {
    if(!a_off || !b_off || n_shards == 0 || !range_first || !range_last || !range_shard) return 0;
    std::vector<std::pair<size_t, size_t>> chunks;
    std::vector<double> cost;
    plan_chunks(npairs, a_off, b_off, n_shards, chunks, cost);
    if(chunks.size() > max_ranges) return 0;
    std::vector<double> load(n_shards, 0.0);
    for(size_t j = 0; j < chunks.size(); ++j) {
        const uint32_t sh = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
        load[sh] += cost[j];
        range_first[j] = chunks[j].first, range_last[j] = chunks[j].second, range_shard[j] = sh;
    }
    return chunks.size();
}

// alignpair for the ranges [first[j], last[j]) of one CSR batch (a shard of coati_gpu_plan_shards): inputs and
// outputs are addressed exactly as in coati_gpu_alignpair_batch, pairs outside the ranges are not touched.
extern "C" int coati_gpu_alignpair_batch_ranges(coati_gpu_ctx* ctx, size_t npairs, const char* anc_all,
                                                const uint64_t* anc_off, const char* des_all,
                                                const uint64_t* des_off, char* out_a, char* out_b,
                                                uint64_t* out_len, float* score, int32_t* status, size_t n_ranges,
                                                const uint64_t* first, const uint64_t* last) {
    if(!ctx || !ctx->model_set || (npairs && (!anc_off || !des_off || !anc_all || !des_all))) return COATI_GPU_E_ARG;
    if(n_ranges && (!first || !last)) return COATI_GPU_E_ARG;
    BatchArgs A{npairs, nullptr, anc_off, nullptr, des_off, anc_all, des_all, out_a, out_b, out_len, score,
                status, true, nullptr};
    RangeQueue q;
    size_t covered = 0;
    for(size_t j = 0; j < n_ranges; ++j) {
        if(first[j] > last[j] || last[j] > npairs) return COATI_GPU_E_ARG;
        if(first[j] < last[j]) q.add(first[j], last[j], anc_off, des_off), covered += last[j] - first[j];
    }
    if(q.ranges.empty()) return COATI_GPU_OK;
    A.shared_link = covered < npairs;  // a shard: the rest of the batch is some other device's, at the same time
    return viterbi_batch_pipeline(ctx, A, q, true);
}

// ---- pinned host memory for callers --------------------------------------------------------------------
// The pipelined calls overlap H2D, kernels and D2H only when the caller's arenas are page-locked (a copy
// from or to pageable memory is staged by the driver and blocks the issuing thread).  A C++ caller holding
// std::string / std::vector either allocates its arenas here or registers them once.
extern "C" void* coati_gpu_host_alloc(size_t bytes) {
    void* p = nullptr;
    if(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void coati_gpu_host_free(void* p) {
    if(p) cudaFreeHost(p);
}
extern "C" int coati_gpu_host_register(void* p, size_t bytes) {
    if(!p || !bytes) return COATI_GPU_E_ARG;
    if(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        cudaGetLastError();
        return COATI_GPU_E_CUDA;
    }
    return COATI_GPU_OK;
}
extern "C" int coati_gpu_host_unregister(void* p) {
    if(!p) return COATI_GPU_E_ARG;
    if(cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return COATI_GPU_E_CUDA;
    }
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_viterbi(coati_gpu_ctx* ctx, const uint8_t* a, size_t La, const uint8_t* b,
                                 size_t Lb, const char* anc, const char* des, char* out_a,
                                 char* out_b, size_t* out_len, float* score) {
    if(!ctx || !out_a || !out_b) return COATI_GPU_E_ARG;
    const uint64_t a_off[2] = {0, La}, b_off[2] = {0, Lb};
    uint64_t len = 0;
    float sc = 0.f;
    int32_t st = 0;
    int rc = coati_gpu_viterbi_batch(ctx, 1, a, a_off, b, b_off, anc, des, out_a, out_b, &len, &sc,
                                     &st);
    if(rc != COATI_GPU_OK) return rc;
    if(st != COATI_GPU_OK) return st;
    if(out_len) *out_len = len;
    if(score) *score = sc;
    return COATI_GPU_OK;
}

// ---------------------------------------------------------------------------------------------
// canonical decision byte per body cell, row-major, from either stream layout
template <class Layout>
__global__ void unpack_dirs_kernel(const uint8_t* __restrict__ dirs, PairDesc pd,
                                   uint8_t* __restrict__ out) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(idx >= (uint64_t)pd.la * pd.lb) return;
    const uint32_t r = idx / pd.lb + 1, c = idx % pd.lb + 1;
    const uint8_t* d = dirs + pd.dir_off;
    const int x = Layout::next(d, pd, ST_M, r, c), y = Layout::next(d, pd, ST_D, r, c),
              z = Layout::next(d, pd, ST_I, r, c);
    out[idx] = (uint8_t)(x | (y << 2) | ((z == ST_I ? 1 : 0) << 4));
}

extern "C" int coati_gpu_viterbi_directions(coati_gpu_ctx* ctx, const uint8_t* a, size_t La,
                                            const uint8_t* b, size_t Lb, uint8_t* dirs,
                                            float* score) {
    if(!ctx || !dirs) return COATI_GPU_E_ARG;
    const uint64_t a_off[2] = {0, La}, b_off[2] = {0, Lb};
    coati_gpu_batch* bt = nullptr;
    int rc = coati_gpu_batch_create(ctx, 1, a_off, b_off, &bt);
    if(rc != COATI_GPU_OK) return rc;
    std::vector<char> dummy_a(La + 1, 'A'), dummy_b(Lb + 1, 'A');
    rc = coati_gpu_batch_upload(bt, a, b, dummy_a.data(), dummy_b.data());
    if(rc == COATI_GPU_OK) rc = coati_gpu_batch_run(bt);
    int32_t st = 0;
    float sc = 0.f;
    if(rc == COATI_GPU_OK) rc = coati_gpu_batch_download(bt, nullptr, nullptr, nullptr, &sc, &st);
    if(rc == COATI_GPU_OK && st != 0) rc = st;
    if(rc == COATI_GPU_OK && La * Lb > 0) {
        DevBuf<uint8_t> rowmajor;
        if(rowmajor.alloc(La * Lb) != cudaSuccess) {
            rc = COATI_GPU_E_NOMEM;
        } else {
            const uint64_t n = La * Lb;
            const PairDesc pd = bt->descs[0];
            const unsigned grid = (unsigned)((n + 255) / 256);
            if(pd.cfg & 0x1ffu)
                unpack_dirs_kernel<PipeLayout><<<grid, 256, 0, ctx->stream>>>(bt->d_dirs.p, pd, rowmajor.p);
            else
                unpack_dirs_kernel<DiagLayout><<<grid, 256, 0, ctx->stream>>>(bt->d_dirs.p, pd, rowmajor.p);
            ++ctx->launches;
            if(cudaMemcpyAsync(dirs, rowmajor.p, n, cudaMemcpyDeviceToHost, ctx->stream) !=
                   cudaSuccess ||
               cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
                ctx->last_error = cudaGetErrorString(cudaGetLastError());
                rc = COATI_GPU_E_CUDA;
            }
        }
    }
    if(rc == COATI_GPU_OK && score) *score = sc;
    coati_gpu_batch_destroy(bt);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Forward fill + sampleback: a handle owns the state matrices of one pair or of a batch of pairs
struct coati_gpu_forward_t {
    coati_gpu_ctx* ctx = nullptr;
    size_t npairs = 1;
    uint32_t la = 0, lb = 0;       // of pair 0 (the single-pair entry points)
    std::vector<FwdDesc> descs;    // a_off / b_off relative to the first pair's; mat_off in floats
    std::vector<uint64_t> out_off; // batch sampling: first row byte of every pair (for n samples)
    uint64_t a_total = 0, b_total = 0;
    DevBuf<uint8_t> d_a, d_b;
    DevBuf<char> d_anc, d_des, d_out_a, d_out_b;
    DevBuf<FwdDesc> d_desc;
    DevBuf<float> d_mats, d_term, d_scores;
    DevBuf<uint64_t> d_rng, d_out_off;
    DevBuf<uint32_t> d_len, d_start;
    DevBuf<int32_t> d_status;
    DevBuf<unsigned int> d_counter;
    DevBuf<SampleRec> d_rec;   // per (cell, state) sampling records (sample_spec.cuh), built lazily
    DevBuf<U128> d_pw;         // MULT^(2^i)
    DevBuf<uint32_t> d_draws;
    DevBuf<uint64_t> d_starts, d_cursor;
    bool rec_ready = false;
    uint64_t model_gen = 0;  // the context's model when the matrices were filled
    std::vector<float> terms;  // adjusted terminal M, D, I per pair
    float term[3] = {0, 0, 0};
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    float fill_ms = 0, sample_ms = 0;
    ~coati_gpu_forward_t() {
        for(cudaEvent_t e : ev)
            if(e) cudaEventDestroy(e);
    }
};

// forward (align_pair.cc:149-152) for npairs pairs (CSR); offsets may start anywhere
static int forward_fill(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all, const uint64_t* a_off,
                        const uint8_t* b_all, const uint64_t* b_off, coati_gpu_forward_t** out) {
    if(!ctx || !out || !ctx->model_set || npairs == 0 || !a_off || !b_off || npairs > 0x7fffffffull)
        return COATI_GPU_E_ARG;
    *out = nullptr;
    const uint32_t k = ctx->gap.k;
    std::unique_ptr<coati_gpu_forward_t> h(new(std::nothrow) coati_gpu_forward_t);
    if(!h) return COATI_GPU_E_NOMEM;
    h->ctx = ctx;
    h->model_gen = ctx->model_gen;
    h->npairs = npairs;
    const uint64_t a0 = a_off[0], b0 = b_off[0];
    h->a_total = a_off[npairs] - a0;
    h->b_total = b_off[npairs] - b0;
    if((h->a_total && !a_all) || (h->b_total && !b_all)) return COATI_GPU_E_ARG;
    uint64_t mat_total = 0;
    try {
        h->descs.resize(npairs);
        h->terms.assign(3 * npairs, 0.0f);
    } catch(const std::bad_alloc&) {
        return COATI_GPU_E_NOMEM;
    }
    for(size_t p = 0; p < npairs; ++p) {
        const uint64_t La = a_off[p + 1] - a_off[p], Lb = b_off[p + 1] - b_off[p];
        if(La > 0x7fffffffull || Lb > 0x7fffffffull) return COATI_GPU_E_ARG;
        if(La % k != 0 || Lb % k != 0) return COATI_GPU_E_LENGTH;
        for(uint64_t x = a_off[p]; x < a_off[p + 1]; ++x)
            if(a_all[x] >= TABLE_ROWS) return COATI_GPU_E_SYMBOL;
        for(uint64_t x = b_off[p]; x < b_off[p + 1]; ++x)
            if(b_all[x] >= TABLE_COLS) return COATI_GPU_E_SYMBOL;
        h->descs[p] = FwdDesc{a_off[p] - a0, b_off[p] - b0, mat_total, (uint32_t)La, (uint32_t)Lb};
        mat_total += 3 * (La + 1) * (Lb + 1);
    }
    h->la = h->descs[0].la;
    h->lb = h->descs[0].lb;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if(e == cudaSuccess) e = r;
    };
    ok(h->d_a.alloc(h->a_total + 1, &ctx->pool));
    ok(h->d_b.alloc(h->b_total + 1, &ctx->pool));
    ok(h->d_desc.alloc(npairs, &ctx->pool));
    ok(h->d_mats.alloc(mat_total, &ctx->pool));
    ok(h->d_term.alloc(3 * npairs + 1, &ctx->pool));
    ok(h->d_counter.alloc(npairs + 1, &ctx->pool));
    if(e != cudaSuccess) {
        ctx->last_error = std::string("forward allocation: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return COATI_GPU_E_NOMEM;
    }
    for(cudaEvent_t& ev : h->ev) CU_TRY(ctx, cudaEventCreate(&ev));
    cudaStream_t s = ctx->stream;
    CU_TRY(ctx, cudaMemcpyAsync(h->d_desc.p, h->descs.data(), npairs * sizeof(FwdDesc), cudaMemcpyHostToDevice, s));
    if(h->a_total) CU_TRY(ctx, cudaMemcpyAsync(h->d_a.p, a_all + a0, h->a_total, cudaMemcpyHostToDevice, s));
    if(h->b_total) CU_TRY(ctx, cudaMemcpyAsync(h->d_b.p, b_all + b0, h->b_total, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemsetAsync(h->d_counter.p, 0, (npairs + 1) * sizeof(unsigned int), s));
    CU_TRY(ctx, cudaEventRecord(h->ev[0], s));
    const uint32_t sms = (uint32_t)ctx->prop.multiProcessorCount;
    if(k != 1 || ctx->forward_generic) {
        // any gap unit length: CTA per pair, anti-diagonal sweep over the stored planes (forward.cuh)
        forward_fill_kernel<<<(unsigned)std::min<size_t>(npairs, 4 * sms), 1023, 0, s>>>(
            h->d_desc.p, (uint32_t)npairs, h->d_a.p, h->d_b.p, ctx->d_table, ctx->gap, h->d_mats.p, h->d_term.p);
        ++ctx->launches;
    } else if(npairs <= 16) {
        // few pairs: each as a wavefront of 10-row bands over the whole GPU; stored values are their own
        // ready flags, so the matrices start as NaN
        CU_TRY(ctx, cudaMemsetAsync(h->d_mats.p, 0xff, mat_total * sizeof(float), s));
        for(size_t p = 0; p < npairs; ++p) {
            const uint32_t nbands = std::max(1u, (h->descs[p].la + FB_ROWS - 1) / FB_ROWS);
            const uint32_t grid = std::min((nbands + FB_WARPS - 1) / FB_WARPS, sms * (uint32_t)ctx->fwd_ctas_per_sm[1]);
            forward_band_kernel<true><<<grid, FB_WARPS * 32, 0, s>>>(
                h->d_desc.p, (uint32_t)p, (uint32_t)p + 1, h->d_counter.p + p, h->d_a.p, h->d_b.p, ctx->d_table,
                ctx->gap, h->d_mats.p, h->d_term.p);
            ++ctx->launches;
        }
    } else {
        const uint32_t grid = (uint32_t)std::min<size_t>((npairs + FB_WARPS - 1) / FB_WARPS,
                                                         (size_t)sms * ctx->fwd_ctas_per_sm[0]);
        forward_band_kernel<false><<<grid, FB_WARPS * 32, 0, s>>>(
            h->d_desc.p, 0, (uint32_t)npairs, h->d_counter.p, h->d_a.p, h->d_b.p, ctx->d_table, ctx->gap,
            h->d_mats.p, h->d_term.p);
        ++ctx->launches;
    }
    CU_TRY(ctx, cudaEventRecord(h->ev[1], s));
    CU_TRY(ctx, cudaMemcpyAsync(h->terms.data(), h->d_term.p, 3 * npairs * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaEventElapsedTime(&h->fill_ms, h->ev[0], h->ev[1]));
    for(int x = 0; x < 3; ++x) h->term[x] = h->terms[x];
    *out = h.release();
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_forward(coati_gpu_ctx* ctx, const uint8_t* a, size_t La, const uint8_t* b,
                                 size_t Lb, coati_gpu_forward_t** out) {
    if((La && !a) || (Lb && !b)) return COATI_GPU_E_ARG;
    const uint64_t a_off[2] = {0, La}, b_off[2] = {0, Lb};
    return forward_fill(ctx, 1, a, a_off, b, b_off, out);
}

extern "C" int coati_gpu_forward_batch(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all,
                                       const uint64_t* a_off, const uint8_t* b_all, const uint64_t* b_off,
                                       coati_gpu_forward_t** out) {
    return forward_fill(ctx, npairs, a_all, a_off, b_all, b_off, out);
}

extern "C" int coati_gpu_forward_batch_terminal(coati_gpu_forward_t* h, float* term, float* loglik,
                                                float* fill_ms) {
    if(!h) return COATI_GPU_E_ARG;
    for(size_t p = 0; p < h->npairs; ++p) {
        const float* t = &h->terms[3 * p];
        if(term)
            for(int x = 0; x < 3; ++x) term[3 * p + x] = t[x];
        if(loglik) {
            // log_sum_exp(log_sum_exp(M, D), I) with the host libm (utils.hpp:134-156)
            auto l1pe = [](float x) -> float {
                if(x <= -16.0f) return ::expf(x);
                if(x <= 8.0f) return ::log1pf(::expf(x));
                if(x <= 14.5f) return x + ::expf(-x);
                return x;
            };
            auto lse = [&](float a, float b) -> float { return std::max(a, b) + l1pe(-std::fabs(a - b)); };
            loglik[p] = lse(lse(t[0], t[1]), t[2]);
        }
    }
    if(fill_ms) *fill_ms = h->fill_ms;
    return COATI_GPU_OK;
}

// n serial samplebacks per pair, every pair on its own RNG stream (one thread per pair): the batch form of
// marg_sample's loop (align_marginal.cc:589-593).  rng_states: 2 x uint64 per pair, in-out.  Rows of
// sample s of pair p start at byte n * (a_off[p] + b_off[p] + p) + s * (La_p + Lb_p + 1) (offsets relative
// to the first pair's); out_len / scores are [p * n + s].
extern "C" int coati_gpu_sampleback_batch(coati_gpu_forward_t* h, const char* anc_all, const char* des_all,
                                          uint64_t* rng_states, size_t n, char* out_a, char* out_b,
                                          uint64_t* out_len, float* scores, float* sample_ms) {
    if(!h || !rng_states || (n && (!out_a || !out_b)) || n > 0x7fffffffull) return COATI_GPU_E_ARG;
    if((h->a_total && !anc_all) || (h->b_total && !des_all)) return COATI_GPU_E_ARG;
    coati_gpu_ctx* ctx = h->ctx;
    if(h->model_gen != ctx->model_gen) return COATI_GPU_E_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t np = h->npairs;
    const uint64_t out_total = n * (h->a_total + h->b_total + np);
    std::vector<uint64_t> st(2 * np);
    try {
        h->out_off.resize(np);
    } catch(const std::bad_alloc&) {
        return COATI_GPU_E_NOMEM;
    }
    for(size_t p = 0; p < np; ++p) {
        h->out_off[p] = n * (h->descs[p].a_off + h->descs[p].b_off + p);
        st[2 * p] = rng_states[2 * p] | 1ull;  // Lehmer64Fast::SetState (random.hpp:131-134)
        st[2 * p + 1] = rng_states[2 * p + 1];
    }
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if(e == cudaSuccess) e = r;
    };
    ok(h->d_anc.alloc(h->a_total + 1, &ctx->pool));
    ok(h->d_des.alloc(h->b_total + 1, &ctx->pool));
    ok(h->d_out_a.alloc(out_total + 1, &ctx->pool));
    ok(h->d_out_b.alloc(out_total + 1, &ctx->pool));
    ok(h->d_rng.alloc(2 * np, &ctx->pool));
    ok(h->d_out_off.alloc(np, &ctx->pool));
    ok(h->d_len.alloc(np * n + 1, &ctx->pool));
    ok(h->d_start.alloc(np * n + 1, &ctx->pool));
    ok(h->d_scores.alloc(np * n + 1, &ctx->pool));
    ok(h->d_status.alloc(np, &ctx->pool));
    if(e != cudaSuccess) {
        ctx->last_error = std::string("sampleback allocation: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return COATI_GPU_E_NOMEM;
    }
    if(h->a_total) CU_TRY(ctx, cudaMemcpyAsync(h->d_anc.p, anc_all, h->a_total, cudaMemcpyHostToDevice, s));
    if(h->b_total) CU_TRY(ctx, cudaMemcpyAsync(h->d_des.p, des_all, h->b_total, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(h->d_rng.p, st.data(), st.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(h->d_out_off.p, h->out_off.data(), np * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaEventRecord(h->ev[1], s));
    sampleback_kernel<<<(unsigned)((np + 31) / 32), 32, 0, s>>>(
        h->d_desc.p, (uint32_t)np, h->d_mats.p, h->d_term.p, ctx->d_table, h->d_a.p, h->d_b.p, h->d_anc.p,
        h->d_des.p, ctx->gap, h->d_rng.p, (uint32_t)n, h->d_out_off.p, h->d_out_a.p, h->d_out_b.p, h->d_len.p,
        h->d_start.p, h->d_scores.p, h->d_status.p);
    if(n)
        compact_samples_kernel<<<(unsigned)((np * n + 7) / 8), 256, 0, s>>>(
            h->d_desc.p, (uint32_t)np, (uint32_t)n, h->d_out_off.p, h->d_out_a.p, h->d_out_b.p, h->d_len.p,
            h->d_start.p);
    ctx->launches += 2;
    CU_TRY(ctx, cudaEventRecord(h->ev[2], s));
    std::vector<uint32_t> lens(np * n);
    std::vector<int32_t> status(np);
    if(n) {
        CU_TRY(ctx, cudaMemcpyAsync(out_a, h->d_out_a.p, out_total, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaMemcpyAsync(out_b, h->d_out_b.p, out_total, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaMemcpyAsync(lens.data(), h->d_len.p, np * n * 4, cudaMemcpyDeviceToHost, s));
        if(scores) CU_TRY(ctx, cudaMemcpyAsync(scores, h->d_scores.p, np * n * 4, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(ctx, cudaMemcpyAsync(st.data(), h->d_rng.p, st.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaMemcpyAsync(status.data(), h->d_status.p, np * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaEventElapsedTime(&h->sample_ms, h->ev[1], h->ev[2]));
    if(sample_ms) *sample_ms = h->sample_ms;
    if(out_len)
        for(size_t x = 0; x < np * n; ++x) out_len[x] = lens[x];
    for(size_t x = 0; x < 2 * np; ++x) rng_states[x] = st[x];
    for(size_t p = 0; p < np; ++p)
        if(status[p] != 0) return status[p];
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_forward_terminal(coati_gpu_forward_t* h, float term[3], float* fill_ms) {
    if(!h) return COATI_GPU_E_ARG;
    if(term)
        for(int x = 0; x < 3; ++x) term[x] = h->term[x];
    if(fill_ms) *fill_ms = h->fill_ms;
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_forward_matrices(coati_gpu_forward_t* h, float* mch, float* del, float* ins) {
    if(!h || !mch || !del || !ins || h->npairs != 1) return COATI_GPU_E_ARG;
    coati_gpu_ctx* ctx = h->ctx;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t plane = (uint64_t)(h->la + 1) * (h->lb + 1);
    cudaStream_t s = ctx->stream;
    CU_TRY(ctx, cudaMemcpyAsync(mch, h->d_mats.p, plane * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaMemcpyAsync(del, h->d_mats.p + plane, plane * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaMemcpyAsync(ins, h->d_mats.p + 2 * plane, plane * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    return COATI_GPU_OK;
}

extern "C" int coati_gpu_sampleback(coati_gpu_forward_t* h, const char* anc, const char* des,
                                    uint64_t rng_state[2], size_t n, char* out_a, char* out_b,
                                    size_t* out_len, float* scores, float* sample_ms) {
    if(!h || !rng_state || (n && (!out_a || !out_b)) || h->npairs != 1) return COATI_GPU_E_ARG;
    if((h->la && !anc) || (h->lb && !des) || n > 0x7fffffffull) return COATI_GPU_E_ARG;
    coati_gpu_ctx* ctx = h->ctx;
    // the matrices belong to the model they were filled under: sampling them with another table or other
    // gap constants would be silently wrong
    if(h->model_gen != ctx->model_gen) return COATI_GPU_E_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t stride = (size_t)h->la + h->lb + 1;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if(e == cudaSuccess) e = r;
    };
    ok(h->d_anc.alloc(h->la + 1, &ctx->pool));
    ok(h->d_des.alloc(h->lb + 1, &ctx->pool));
    ok(h->d_out_a.alloc(n * stride + 1, &ctx->pool));
    ok(h->d_out_b.alloc(n * stride + 1, &ctx->pool));
    ok(h->d_rng.alloc(2, &ctx->pool));
    ok(h->d_out_off.alloc(1, &ctx->pool));
    ok(h->d_len.alloc(n + 1, &ctx->pool));
    ok(h->d_start.alloc(n + 1, &ctx->pool));
    ok(h->d_scores.alloc(n + 1, &ctx->pool));
    ok(h->d_status.alloc(1, &ctx->pool));
    if(e != cudaSuccess) {
        ctx->last_error = std::string("sampleback allocation: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return COATI_GPU_E_NOMEM;
    }
    const uint64_t zero = 0;
    const uint64_t st[2] = {rng_state[0] | 1ull, rng_state[1]};  // Lehmer64Fast::SetState (random.hpp:131-134)
    if(n >= 4 && !ctx->sample_serial) {
        // ---- parallel path: records -> speculative draw counts -> chase -> parallel re-walk ----------
        const uint64_t ncells = (uint64_t)(h->la + 1) * (h->lb + 1);
        const uint64_t max_draws = (uint64_t)h->la + h->lb + 1;
        const uint32_t window = (uint32_t)std::min<uint64_t>(n * max_draws, 1ull << 19);
        U128 pw[64];
        pw[0] = U128{0xda942042e4dd58b5ull, 0};
        for(int i = 1; i < 64; ++i) pw[i] = mul128(pw[i - 1], pw[i - 1]);
        cudaError_t e2 = cudaSuccess;
        auto ok2 = [&](cudaError_t r) {
            if(e2 == cudaSuccess) e2 = r;
        };
        if(!h->rec_ready) ok2(h->d_rec.alloc(3 * ncells + 1, &ctx->pool));
        ok2(h->d_pw.alloc(64, &ctx->pool));
        ok2(h->d_draws.alloc(window, &ctx->pool));
        ok2(h->d_starts.alloc(n, &ctx->pool));
        ok2(h->d_cursor.alloc(4, &ctx->pool));
        if(e2 != cudaSuccess) {
            ctx->last_error = std::string("sampleback allocation: ") + cudaGetErrorString(e2);
            cudaGetLastError();
            return COATI_GPU_E_NOMEM;
        }
        if(h->la) CU_TRY(ctx, cudaMemcpyAsync(h->d_anc.p, anc, h->la, cudaMemcpyHostToDevice, s));
        if(h->lb) CU_TRY(ctx, cudaMemcpyAsync(h->d_des.p, des, h->lb, cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemcpyAsync(h->d_pw.p, pw, sizeof(pw), cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemsetAsync(h->d_cursor.p, 0, 4 * sizeof(uint64_t), s));
        CU_TRY(ctx, cudaEventRecord(h->ev[1], s));
        const FwdDesc fd{0, 0, 0, h->la, h->lb};
        if(!h->rec_ready) {
            sample_records_kernel<<<(unsigned)((ncells + 127) / 128), 128, 0, s>>>(
                fd, h->d_mats.p, h->d_term.p, ctx->d_table, h->d_a.p, h->d_b.p, ctx->gap, h->d_rec.p);
            ++ctx->launches;
            h->rec_ready = true;
        }
        const U128 state0{st[0], st[1]};
        uint64_t cursor[3] = {0, 0, 0};
        while(cursor[1] < n) {
            const uint64_t base = cursor[0];
            spec_steps_kernel<<<(window + 127) / 128, 128, 0, s>>>(h->d_rec.p, h->la, h->lb, ctx->gap.k, state0,
                                                                     h->d_pw.p, base, window, h->d_draws.p);
            chase_kernel<<<1, 1, 0, s>>>(h->d_draws.p, base, window, n, h->d_cursor.p, h->d_starts.p);
            ctx->launches += 2;
            CU_TRY(ctx, cudaMemcpyAsync(cursor, h->d_cursor.p, sizeof(cursor), cudaMemcpyDeviceToHost, s));
            CU_TRY(ctx, cudaStreamSynchronize(s));
            if(cursor[2] != 0) return COATI_GPU_E_INTERNAL;
            if(cursor[0] == base && cursor[1] < n) return COATI_GPU_E_INTERNAL;  // no progress
        }
        sample_paths_kernel<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(h->d_rec.p, h->la, h->lb, ctx->gap.k, state0,
                                                                      h->d_pw.p, h->d_starts.p, (uint32_t)n,
                                                                      h->d_out_b.p, h->d_len.p, h->d_start.p,
                                                                      h->d_scores.p);
        expand_samples_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(h->la, h->lb, (uint32_t)n, h->d_anc.p,
                                                                      h->d_des.p, h->d_out_a.p, h->d_out_b.p,
                                                                      h->d_len.p, h->d_start.p);
        ctx->launches += 2;
        CU_TRY(ctx, cudaEventRecord(h->ev[2], s));
        std::vector<uint32_t> lens2(n);
        CU_TRY(ctx, cudaMemcpyAsync(out_a, h->d_out_a.p, n * stride, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaMemcpyAsync(out_b, h->d_out_b.p, n * stride, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaMemcpyAsync(lens2.data(), h->d_len.p, n * 4, cudaMemcpyDeviceToHost, s));
        if(scores) CU_TRY(ctx, cudaMemcpyAsync(scores, h->d_scores.p, n * 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaStreamSynchronize(s));
        CU_TRY(ctx, cudaGetLastError());
        CU_TRY(ctx, cudaEventElapsedTime(&h->sample_ms, h->ev[1], h->ev[2]));
        if(sample_ms) *sample_ms = h->sample_ms;
        for(size_t x = 0; x < n; ++x) {
            if(lens2[x] == 0 && (h->la || h->lb)) return COATI_GPU_E_INTERNAL;
            if(out_len) out_len[x] = lens2[x];
        }
        const U128 fin = jump(state0, cursor[0], pw);  // the stream after all draws of the n samples
        rng_state[0] = fin.lo;
        rng_state[1] = fin.hi;
        return COATI_GPU_OK;
    }
    if(h->la) CU_TRY(ctx, cudaMemcpyAsync(h->d_anc.p, anc, h->la, cudaMemcpyHostToDevice, s));
    if(h->lb) CU_TRY(ctx, cudaMemcpyAsync(h->d_des.p, des, h->lb, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(h->d_rng.p, st, sizeof(st), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(h->d_out_off.p, &zero, sizeof(zero), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaEventRecord(h->ev[1], s));
    sampleback_kernel<<<1, 32, 0, s>>>(h->d_desc.p, 1, h->d_mats.p, h->d_term.p, ctx->d_table, h->d_a.p,
                                       h->d_b.p, h->d_anc.p, h->d_des.p, ctx->gap, h->d_rng.p,
                                       (uint32_t)n, h->d_out_off.p, h->d_out_a.p, h->d_out_b.p,
                                       h->d_len.p, h->d_start.p, h->d_scores.p, h->d_status.p);
    if(n) {
        compact_samples_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(
            h->d_desc.p, 1, (uint32_t)n, h->d_out_off.p, h->d_out_a.p, h->d_out_b.p, h->d_len.p,
            h->d_start.p);
    }
    ctx->launches += 2;
    CU_TRY(ctx, cudaEventRecord(h->ev[2], s));
    std::vector<uint32_t> lens(n);
    uint64_t st_out[2] = {0, 0};
    int32_t status = 0;
    if(n) {
        CU_TRY(ctx, cudaMemcpyAsync(out_a, h->d_out_a.p, n * stride, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaMemcpyAsync(out_b, h->d_out_b.p, n * stride, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaMemcpyAsync(lens.data(), h->d_len.p, n * 4, cudaMemcpyDeviceToHost, s));
        if(scores) CU_TRY(ctx, cudaMemcpyAsync(scores, h->d_scores.p, n * 4, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(ctx, cudaMemcpyAsync(st_out, h->d_rng.p, sizeof(st_out), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaMemcpyAsync(&status, h->d_status.p, sizeof(status), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaEventElapsedTime(&h->sample_ms, h->ev[1], h->ev[2]));
    if(sample_ms) *sample_ms = h->sample_ms;
    if(out_len)
        for(size_t x = 0; x < n; ++x) out_len[x] = lens[x];
    rng_state[0] = st_out[0];
    rng_state[1] = st_out[1];
    return status;
}

extern "C" void coati_gpu_forward_free(coati_gpu_forward_t* h) {
    if(!h) return;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    delete h;
}

extern "C" int coati_gpu_libm_eval(coati_gpu_ctx* ctx, int op, const float* in, float* out, size_t n) {
    if(!ctx || !in || !out || op < 0 || op > 4) return COATI_GPU_E_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf<float> d_in, d_out;
    if(d_in.alloc(n + 1, &ctx->pool) != cudaSuccess || d_out.alloc(n + 1, &ctx->pool) != cudaSuccess)
        return COATI_GPU_E_NOMEM;
    cudaStream_t s = ctx->stream;
    CU_TRY(ctx, cudaMemcpyAsync(d_in.p, in, n * 4, cudaMemcpyHostToDevice, s));
    libm_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(op, d_in.p, d_out.p, n);
    ++ctx->launches;
    CU_TRY(ctx, cudaMemcpyAsync(out, d_out.p, n * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    CU_TRY(ctx, cudaGetLastError());
    return COATI_GPU_OK;
}
