// Traceback over the packed direction stream, then expansion of the path into the two rows.
//
// Follows traceback<S>, src/lib/align_pair.cc:249-303: start at the terminal cell with
// max_mdi(M, D, I) of the adjusted terminal scores, then walk MATCH (-1,-1) / DELETION (-k, 0) /
// INSERTION (0, -k) until (0, 0); the next state is the decision byte of the cell just landed on
// (common.cuh: direction_byte), or implied on the margins where only one state is finite.
#pragma once

#include "common.cuh"
#include "viterbi_pipe.cuh"

namespace coati_gpu {

// Generic-k kernels: one decision byte per body cell, anti-diagonal-major (common.cuh).
struct DiagLayout {
    __device__ __forceinline__ static uint32_t touch(const uint8_t* dir, const PairDesc& pd, uint32_t r,
                                                     uint32_t c) {
        return dir[dir_index_diag(r, c, pd.la, pd.lb)];
    }
    __device__ __forceinline__ static int initial(const uint8_t*, const PairDesc&, PairResult& res) {
        const float tM = res.term[0], tD = res.term[1], tI = res.term[2];
        res.score = fmaxf(fmaxf(tM, tD), tI);  // align_pair.cc:265
        return max_mdi(tM, tD, tI);            // :266
    }
    __device__ __forceinline__ static int next(const uint8_t* dir, const PairDesc& pd, int st,
                                               uint32_t r, uint32_t c) {
        const uint32_t byte = dir[dir_index_diag(r, c, pd.la, pd.lb)];
        return st == ST_M ? (byte & 3) : st == ST_D ? ((byte >> 2) & 3) : ((byte >> 4) & 1) * 2;
    }
};

// Pipelined kernels: five bit-planes per row (viterbi_pipe.cuh).  R (= pd.cfg & 0xff) is a template
// constant so the index arithmetic of the serial walk is shifts and masks, not integer divisions.
template <int R>
struct PipeLayoutR {
    static constexpr uint32_t H = 32 * R, WPL = (5 * R + 3) & ~3u, BS = pipe_block_steps(R);
    __device__ __forceinline__ static uint64_t word_index(const PairDesc& pd, uint32_t r, uint32_t c,
                                                          uint32_t& shift) {
        const uint32_t band = (r - 1) / H, rr = (r - 1) % H, lane = rr / R, q = rr % R;
        const uint32_t t = (c - 1) + lane, nblocks = (pd.lb + 31 + BS - 1) / BS;
        shift = 31 - (t % BS);
        // one widening multiply; the in-block part stays 32-bit
        return (uint64_t)(band * nblocks + t / BS) * (32u * WPL) + (lane * WPL + q * 5);
    }
    __device__ __forceinline__ static int next(const uint8_t* dir, const PairDesc& pd, int st,
                                               uint32_t r, uint32_t c) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(dir);
        uint32_t sh;
        const uint64_t idx = word_index(pd, r, c, sh);
        // planes 0,1 (arrived by MATCH), 2,3 (DELETION) or 4 (INSERTION); both words are fetched
        // together so the second never waits for the first
        const uint32_t* q = w + idx + (st == ST_M ? 0 : st == ST_D ? 2 : 3);
        uint32_t w0, w1;
        asm volatile("ld.global.nc.u32 %0, [%2];\n\tld.global.nc.u32 %1, [%2+4];" : "=r"(w0), "=r"(w1) : "l"(q));
        const uint32_t b0 = (w0 >> sh) & 1u, b1 = (w1 >> sh) & 1u;
        if(st == ST_I) return b1 ? ST_M : ST_I;       // q -> planes 3,4: plane 4 is w1
        return b0 ? ST_M : (b1 ? ST_D : ST_I);
    }
    // touch the cache lines holding the decisions of cell (r, c) (warp-cooperative read-ahead)
    __device__ __forceinline__ static uint32_t touch(const uint8_t* dir, const PairDesc& pd, uint32_t r,
                                                     uint32_t c) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(dir);
        uint32_t sh;
        const uint64_t idx = word_index(pd, r, c, sh) - ((r - 1) % R) * 5;  // start of the lane's block
        uint32_t v = __ldg(w + idx + WPL - 1);
#pragma unroll
        for(uint32_t x = 0; x < WPL; x += 8) v ^= __ldg(w + idx + x);  // L1 fills 32-byte sectors
        return v;
    }
    // score = X(La, Lb) was written by the fill; max_mdi of the adjusted terminal scores is the
    // MATCH-lands decision of the terminal cell (align_pair.cc:130-138, 265-266).
    __device__ __forceinline__ static int initial(const uint8_t* dir, const PairDesc& pd,
                                                  PairResult&) {
        return next(dir, pd, ST_M, pd.la, pd.lb);
    }
};

// run-time R (debug unpack path only)
struct PipeLayout {
    __device__ __forceinline__ static int next(const uint8_t* dir, const PairDesc& pd, int st,
                                               uint32_t r, uint32_t c) {
        switch(pd.cfg & 0xffu) {
        case 2: return PipeLayoutR<2>::next(dir, pd, st, r, c);
        case 3: return PipeLayoutR<3>::next(dir, pd, st, r, c);
        case 4: return PipeLayoutR<4>::next(dir, pd, st, r, c);
        case 6: return PipeLayoutR<6>::next(dir, pd, st, r, c);
        case 10: return PipeLayoutR<10>::next(dir, pd, st, r, c);
        default: return PipeLayoutR<8>::next(dir, pd, st, r, c);
        }
    }
};

// WARP = false: one thread per pair (batches: thousands of independent walks hide the latency).
// WARP = true : one warp per pair (long pairs): every lane runs the same walk (uniform loads), lane 0
//               writes the rows, and every 8 steps lane j reads ahead the decision words the path
//               reaches in 64 + 8j steps if it keeps to its diagonal, so the serial walk finds its
//               words in L1/L2 instead of paying a DRAM round trip per step.
// The walk is a serial dependent chain (index -> load -> decode -> move), so it does nothing else:
// it records one op byte per alignment column (0 = M, 1 = D, 2 = I), right-aligned in the pair's
// out_b slot; expand_rows_kernel then builds both rows in parallel with coalesced accesses.
template <class Layout, bool WARP>
__device__ __forceinline__ void traceback_walk(const PairDesc& pd, const uint8_t* __restrict__ dirs,
                                               const GapConsts& gap, char* __restrict__ out_b,
                                               PairResult& res) {
    const uint32_t lane = WARP ? (threadIdx.x & 31) : 0;
    const bool lead = lane == 0;
    const uint32_t la = pd.la, lb = pd.lb, k = gap.k;
    const uint8_t* dir = dirs + pd.dir_off;
    char* ops = out_b + pd.out_off;

    uint32_t r = la, c = lb, pos = la + lb;
    int st;
    if(la == 0 || lb == 0) {
        // no body cell: the path is the margin itself (align_pair.cc:82-90, 130-138)
        const GapConsts g = gap;
        if(la == 0 && lb == 0) {
            res.score = (0.0f + g.ng) + g.ng;
            st = ST_M;
        } else if(la == 0) {
            res.score = ((g.go + g.ge * (float)(lb + k - 2)) + g.gs) + g.ng;
            st = ST_I;
        } else {
            res.score = ((g.ng + g.go) + g.ge * (float)(la + k - 2)) + g.gs;
            st = ST_D;
        }
    } else {
        st = Layout::initial(dir, pd, res);
    }
    int err = 0;
    uint32_t nstep = 0, sink = 0;
    if(k == 1 && la > 0 && lb > 0) {
        // Lean walk for gap unit 1 (a single warp runs dependent code at ~6 cycles per instruction, so
        // the step is kept to: emit, move, index, two loads, decode).  Inside the lattice body every
        // move is legal; on a margin only one state is finite (align_pair.cc:84-90), so the rest of
        // the path is emitted in bulk.
        for(;;) {
            --pos;
            if(lead) ops[pos] = (char)st;
            r -= (st != ST_I);
            c -= (st != ST_D);
            if(r == 0 || c == 0) break;
            if(WARP && (nstep++ & 7) == 0) {
                const uint32_t ahead = 64 + 8 * lane;
                if(r > ahead && c > ahead) sink ^= Layout::touch(dir, pd, r - ahead, c - ahead);
            }
            st = Layout::next(dir, pd, st, r, c);
        }
        if(lead) {
            for(; c > 0; --c) ops[--pos] = ST_I;  // top margin: insertions only
            for(; r > 0; --r) ops[--pos] = ST_D;  // left margin: deletions only
        } else {
            pos -= r + c;
        }
        r = c = 0;
    }
    while(r > 0 || c > 0) {  // :268  (j > k-1 || i > k-1)
        if(WARP && (nstep++ & 7) == 0) {
            const uint32_t ahead = 64 + 8 * lane;
            if(r > ahead && c > ahead) sink ^= Layout::touch(dir, pd, r - ahead, c - ahead);
        }
        if(st == ST_M) {
            if(r == 0 || c == 0) { err = 1; break; }
            --pos;
            if(lead) ops[pos] = ST_M;
            --r, --c;
        } else if(st == ST_D) {
            if(r < k) { err = 1; break; }
            for(uint32_t q = 0; q < k; ++q) {
                --pos;
                if(lead) ops[pos] = ST_D;
            }
            r -= k;
        } else {
            if(c < k) { err = 1; break; }
            for(uint32_t q = 0; q < k; ++q) {
                --pos;
                if(lead) ops[pos] = ST_I;
            }
            c -= k;
        }
        if(r == 0 && c == 0) break;
        int nst;
        if(r == 0) nst = ST_I;        // only ins(start, j) is finite on the top margin (:88-90)
        else if(c == 0) nst = ST_D;   // only del(i, start) is finite on the left margin (:84-87)
        else nst = Layout::next(dir, pd, st, r, c);
        st = nst;
    }
    if(!lead) {
        if(sink == 0x9e3779b9u) res.pad = sink;  // keeps the read-ahead loads alive; never true in practice
        return;
    }
    if(err) {
        res.status = -8;  // COATI_GPU_E_INTERNAL
        res.len = 0;
        res.start = la + lb;
        return;
    }
    res.len = la + lb - pos;
    res.start = pos;
}

template <class Layout, bool WARP>
__global__ void __launch_bounds__(64)
traceback_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                 const uint8_t* __restrict__ dirs, GapConsts gap, char* __restrict__ out_b,
                 PairResult* __restrict__ results) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t p = first + (WARP ? gtid >> 5 : gtid);
    if(p >= last) return;
    const PairDesc pd = pairs[p];
    PairResult& res = results[pd.orig];
    if(res.status != 0) return;
    traceback_walk<Layout, WARP>(pd, dirs, gap, out_b, res);
}

constexpr uint32_t LONG_SEG = 2048;  // columns per independently expanded segment of a long alignment
__host__ __device__ inline uint32_t long_ck_capacity(uint32_t la, uint32_t lb) {
    return (la + lb) / LONG_SEG + 8;  // header + one entry per segment boundary (+ margins, ends)
}

// Long pairs (k = 1): one warp per pair, RUNS of equal moves taken in one round.  In state st at (r, c)
// the next 32 cells along the state's own direction -- diagonal for MATCH, up for DELETION, left for
// INSERTION -- are decoded by the 32 lanes at once (lane j: the state the walk would be in after j + 1
// such moves); a ballot gives the length of the run, its op bytes are written by the lanes in parallel
// and the warp jumps to the end of the run.  The serial chain of the reference's loop
// (align_pair.cc:268-299: index -> load -> decode -> move, once per column) becomes one round per run,
// and the 32 loads of a round are independent.  Every round the lanes also touch the decision words
// 32 + 8j steps further down the current diagonal, so the rounds find their words in L1 / L2.
// CK: record checkpoints for expand_long_kernel (single long pairs); without them the op bytes go to
// the pair's out_b slot and expand_rows_kernel builds the rows (long pairs inside a batch).
template <class Layout, bool CK>
__device__ __forceinline__ void traceback_burst_walk(const PairDesc& pd, const uint8_t* __restrict__ dirs,
                                                     const GapConsts& gap, char* __restrict__ ops,
                                                     uint4* __restrict__ ck, PairResult& res) {
    const uint32_t lane = threadIdx.x & 31;
    if(CK && lane == 0) ck[0] = make_uint4(0, 0, 0, 0);
    if(res.status != 0) return;
    const uint32_t la = pd.la, lb = pd.lb;
    const uint8_t* dir = dirs + pd.dir_off;
    const uint32_t FULL = 0xffffffffu;
    uint32_t r = la, c = lb, pos = la + lb, round = 0, sink = 0;
    // checkpoints (pos, r, c) every >= LONG_SEG columns: ops[pos] is the column that consumes anc[r] /
    // des[c], so expand_long_kernel can expand the segments between checkpoints independently
    uint32_t n_ck = 0, last_ck = pos;
    auto mark = [&](uint32_t ps, uint32_t rr, uint32_t cc) {
        if(CK && lane == 0) ck[1 + n_ck] = make_uint4(ps, rr, cc, 0);
        ++n_ck;
        last_ck = ps;
    };
    mark(pos, r, c);
    if(la == 0 || lb == 0) {  // no body cell: the path is the margin itself (align_pair.cc:82-90, 130-138)
        const GapConsts g = gap;
        if(lane == 0)
            res.score = la == 0 && lb == 0 ? (0.0f + g.ng) + g.ng
                        : la == 0          ? ((g.go + g.ge * (float)(lb - 1)) + g.gs) + g.ng
                                           : ((g.ng + g.go) + g.ge * (float)(la - 1)) + g.gs;
    }
    int st = (la == 0 || lb == 0) ? ST_M : Layout::initial(dir, pd, res);
    if(la > 0 && lb > 0)
    for(;;) {
        const uint32_t dr = st != ST_I, dc = st != ST_D, j1 = lane + 1;
        const bool valid = dr * j1 < r + (1u - dr) && dc * j1 < c + (1u - dc);  // lands inside the body
        int sj = -1;
        if(valid) sj = Layout::next(dir, pd, st, r - dr * j1, c - dc * j1);
        const uint32_t same = __ballot_sync(FULL, sj == st);
        const uint32_t run = same == FULL ? 32u : (uint32_t)__ffs(~same);  // moves of this round (>= 1)
        if(lane < run) ops[pos - 1 - lane] = (char)st;
        pos -= run, r -= dr * run, c -= dc * run;
        if(r == 0 || c == 0) break;
        st = __shfl_sync(FULL, sj, run - 1);
        if(CK && pos + LONG_SEG <= last_ck) mark(pos, r, c);
        if((round++ & 1) == 0) {
            const uint32_t ahead = 32 + 8 * lane;
            if(r > ahead && c > ahead) sink ^= Layout::touch(dir, pd, r - ahead, c - ahead);
        }
    }
    // margins: only one state is finite there (align_pair.cc:84-90)
    while(c > 0) {
        const uint32_t m = min(c, LONG_SEG);
        if(pos != last_ck) mark(pos, r, c);
        for(uint32_t x = lane; x < m; x += 32) ops[pos - 1 - x] = ST_I;
        pos -= m, c -= m;
    }
    while(r > 0) {
        const uint32_t m = min(r, LONG_SEG);
        if(pos != last_ck) mark(pos, r, c);
        for(uint32_t x = lane; x < m; x += 32) ops[pos - 1 - x] = ST_D;
        pos -= m, r -= m;
    }
    if(pos != last_ck) mark(pos, 0, 0);
    if(lane == 0) {
        if(CK) ck[0] = make_uint4(n_ck, 0, 0, 0);
        res.len = la + lb - pos;
        res.start = pos;
    } else if(sink == 0x9e3779b9u) {
        res.pad = sink;  // keeps the read-ahead loads alive; never true in practice
    }
}

template <class Layout>
__global__ void __launch_bounds__(32)
traceback_burst_kernel(const PairDesc* __restrict__ pairs, uint32_t p, const uint8_t* __restrict__ dirs,
                       GapConsts gap, char* __restrict__ ops, uint4* __restrict__ ck,
                       PairResult* __restrict__ results) {
    const PairDesc pd = pairs[p];
    traceback_burst_walk<Layout, true>(pd, dirs, gap, ops, ck, results[pd.orig]);
}

// The longest pairs of a batch (listed by the host): run-at-a-time walks, one warp per pair, so the
// launch that walks the rest one thread per pair is not left waiting for a few thousand-column chains.
constexpr uint32_t BURST_MIN_COLUMNS = 1500;  // la + lb from which a batch pair is walked by a warp
__host__ __device__ inline bool burst_in_batch(const PairDesc& pd, uint32_t k) {
    return k == 1 && (pd.cfg & 0xffu) != 0 && !(pd.cfg & CFG_WAVE) && pd.la > 0 && pd.lb > 0 &&
           pd.la + pd.lb >= BURST_MIN_COLUMNS;
}
__global__ void __launch_bounds__(64)
traceback_burst_list_kernel(const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ list, uint32_t n,
                            const uint8_t* __restrict__ dirs, GapConsts gap, char* __restrict__ out_b,
                            PairResult* __restrict__ results) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if(w >= n) return;
    const PairDesc pd = pairs[list[w]];
    PairResult& res = results[pd.orig];
    char* ops = out_b + pd.out_off;
    switch(pd.cfg & 0xffu) {
    case 2: traceback_burst_walk<PipeLayoutR<2>, false>(pd, dirs, gap, ops, nullptr, res); break;
    case 4: traceback_burst_walk<PipeLayoutR<4>, false>(pd, dirs, gap, ops, nullptr, res); break;
    case 8: traceback_burst_walk<PipeLayoutR<8>, false>(pd, dirs, gap, ops, nullptr, res); break;
    case 10: traceback_burst_walk<PipeLayoutR<10>, false>(pd, dirs, gap, ops, nullptr, res); break;
    default: break;  // not listed by the host
    }
}

// One thread per pair over a whole chunk of inter-pair runs: the layout follows the pair's own kernel
// configuration (rows per lane R, or 0 for the anti-diagonal layout of the generic kernel), so all the
// walks of a chunk -- each a serial, latency-bound chain -- are in flight together instead of one
// launch per configuration.  Pairs are sorted by configuration, so a warp rarely mixes two layouts.
// 64 threads x 32 registers per CTA: fits into the registers the resident fill CTAs of another pipeline
// lane leave free.
__global__ void __launch_bounds__(64, 32)
traceback_chunk_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                       const uint8_t* __restrict__ dirs, GapConsts gap, char* __restrict__ out_b,
                       PairResult* __restrict__ results, uint32_t burst) {
    const uint32_t p = first + blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= last) return;
    const PairDesc pd = pairs[p];
    if(pd.cfg & CFG_WAVE) return;  // long pairs: warp-per-pair launch with read-ahead
    if(burst && burst_in_batch(pd, gap.k)) return;  // walked by traceback_burst_list_kernel
    PairResult& res = results[pd.orig];
    if(res.status != 0) return;
    switch(pd.cfg & 0xffu) {
    case 0: traceback_walk<DiagLayout, false>(pd, dirs, gap, out_b, res); break;
    case 2: traceback_walk<PipeLayoutR<2>, false>(pd, dirs, gap, out_b, res); break;
    case 3: traceback_walk<PipeLayoutR<3>, false>(pd, dirs, gap, out_b, res); break;
    case 4: traceback_walk<PipeLayoutR<4>, false>(pd, dirs, gap, out_b, res); break;
    case 6: traceback_walk<PipeLayoutR<6>, false>(pd, dirs, gap, out_b, res); break;
    case 8: traceback_walk<PipeLayoutR<8>, false>(pd, dirs, gap, out_b, res); break;
    case 10: traceback_walk<PipeLayoutR<10>, false>(pd, dirs, gap, out_b, res); break;
    default: res.status = -8; res.len = 0; res.start = pd.la + pd.lb; break;
    }
}

// One warp per pair: expand the op bytes (right-aligned in the out_b slot, first op at res.start)
// into the two gapped rows, left-aligned and NUL-terminated (align_pair.cc:270-302: MATCH emits
// (anc, des), DELETION (anc, '-'), INSERTION ('-', des); the reference reverses at the end, here the
// ops are simply read front to back).  Source indices are running counts of the ops seen so far
// (warp ballot + popcount).  In-place on out_b is safe: chunk i is read before chunk i is written and
// reads never trail writes (read index = write index + start).
// (64 threads x 32 registers per CTA, see traceback_chunk_kernel)
__global__ void __launch_bounds__(64, 32)
expand_rows_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                                   const char* __restrict__ anc_all, const char* __restrict__ des_all,
                                   char* __restrict__ out_a, char* __restrict__ out_b,
                                   PairResult* __restrict__ results, float stop_gap, uint32_t skip_long) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t p = first + warp;
    if(p >= last) return;
    const PairDesc pd = pairs[p];
    if(skip_long && (pd.cfg & CFG_WAVE)) return;  // expanded segment-wise by expand_long_kernel
    const PairResult res = results[pd.orig];
    const char* anc = anc_all + pd.a_off;
    const char* des = des_all + pd.b_off;
    char* oa = out_a + pd.out_off;
    char* ob = out_b + pd.out_off;
    uint32_t n = res.status == 0 ? res.len : 0;
    const uint32_t shift = res.start;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t ia = 0, ib = 0;
    for(uint32_t base = 0; base < n; base += 32) {
        const uint32_t x = base + lane;
        const int op = x < n ? ob[shift + x] : -1;
        const bool useA = op == ST_M || op == ST_D, useB = op == ST_M || op == ST_I;
        const uint32_t ma = __ballot_sync(0xffffffffu, useA), mb = __ballot_sync(0xffffffffu, useB);
        char va = '-', vb = '-';
        if(useA) va = anc[ia + __popc(ma & lt)];
        if(useB) vb = des[ib + __popc(mb & lt)];
        __syncwarp();
        if(x < n) {
            oa[x] = va;
            ob[x] = vb;
        }
        ia += __popc(ma);
        ib += __popc(mb);
        __syncwarp();
    }
    // restore_end_stops (utils.cc:1044-1063): both / neither -> append as they were; only one -> the
    // codon opposite "---" and the gap penalty log(g * e * e) added to the score
    const bool sa = pd.cfg & CFG_STOP_A, sb = pd.cfg & CFG_STOP_B;
    if(res.status == 0 && (sa || sb)) {
        if(lane < 3) {
            oa[n + lane] = sa ? anc[pd.la + lane] : '-';
            ob[n + lane] = sb ? des[pd.lb + lane] : '-';
        }
        if(lane == 0) {
            results[pd.orig].len = n + 3;
            if(sa != sb) results[pd.orig].score = res.score + stop_gap;
        }
        n += 3;
    }
    if(lane == 0) {
        oa[n] = 0;
        ob[n] = 0;
    }
}

// Rows -> the caller's page-locked arenas (device-visible host memory), written by the GPU itself instead of a
// D2H copy of the output buffers.  A pair's rows own La + Lb + 1 bytes per row and use about half of it, and a
// contiguous copy cannot skip the padding: sending only the bytes that exist halves what a step puts on PCIe
// (1.69 -> 0.87 GB per 1 M pairs of C5), which is what bounds the end-to-end figure from four GPUs up (DESIGN.md
// section 5).  Runs after the expansion kernels (rows and final lengths are in device memory), as a SMALL
// persistent grid: a CTA that waits on the link must not take the place of a fill CTA on every SM (the
// one-stage form of this -- every expansion warp writing to the host, 16 k CTAs -- cost the concurrent fills
// 11 ms per 1 M pairs).  One warp per pair; the link wants large writes, so a lane gathers sixteen bytes of
// the row (byte loads, L1 hits: device and host slots need not be aligned alike) and stores them as one aligned
// 16-byte word, 512 bytes per warp and instruction; a 16-byte unit shared with the neighbouring slot (the first and
// the last of a row) leaves as ONE instruction of sixteen lanes, a byte each -- contiguous bytes of one sector.
constexpr uint32_t R2H_THREADS = 256;
__global__ void __launch_bounds__(R2H_THREADS)
rows_to_host_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                    const char* __restrict__ d_out_a, const char* __restrict__ d_out_b,
                    char* __restrict__ h_out_a, char* __restrict__ h_out_b,
                    const PairResult* __restrict__ results) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for(uint32_t p = first + warp; p < last; p += nwarp) {
        const PairDesc pd = pairs[p];
        const PairResult res = results[pd.orig];
        const uint32_t n_out = (res.status == 0 ? res.len : 0u) + 1u;  // columns (end stops included), NUL
#pragma unroll
        for(int i = 0; i < 2; ++i) {
            const char* src = (i ? d_out_b : d_out_a) + pd.out_off;
            char* dst = (i ? h_out_b : h_out_a) + pd.out_off;
            // the row lives at v = x + mis of the 16-byte aligned address al; its bytes are v in [mis, hi)
            const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u), hi = mis + n_out;
            char* al = dst - mis;
            for(uint32_t base = 0; base < hi; base += 512u) {  // (warp-uniform trip count)
                const uint32_t v0 = base + 16u * lane;
                const bool full = v0 >= mis && v0 + 16u <= hi;
                if(full) {
                    const unsigned char* s8 = reinterpret_cast<const unsigned char*>(src) + (v0 - mis);
                    uint32_t w[4];
#pragma unroll
                    for(int q = 0; q < 4; ++q)
                        w[q] = (uint32_t)s8[4 * q] | ((uint32_t)s8[4 * q + 1] << 8) |
                               ((uint32_t)s8[4 * q + 2] << 16) | ((uint32_t)s8[4 * q + 3] << 24);
                    *reinterpret_cast<uint4*>(al + v0) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                uint32_t part = __ballot_sync(0xffffffffu, !full && v0 < hi && v0 + 16u > mis);
                for(; part; part &= part - 1u) {
                    const uint32_t v = base + 16u * (uint32_t)(__ffs(part) - 1) + lane;
                    if(lane < 16u && v >= mis && v < hi) al[v] = src[v - mis];
                }
            }
        }
    }
}

// Long pairs: one warp per segment between two checkpoints of traceback_burst_kernel (ops in `ops`, rows
// written left-aligned into the pair's output slots); the warp of the last segment also restores the end
// stops and terminates the rows, as expand_rows_kernel does.
__global__ void __launch_bounds__(64, 32)
expand_long_kernel(const PairDesc* __restrict__ pairs, uint32_t p, const char* __restrict__ ops,
                   const uint4* __restrict__ ck, const char* __restrict__ anc_all,
                   const char* __restrict__ des_all, char* __restrict__ out_a, char* __restrict__ out_b,
                   PairResult* __restrict__ results, float stop_gap) {
    const uint32_t seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const PairDesc pd = pairs[p];
    const PairResult res = results[pd.orig];
    const char* anc = anc_all + pd.a_off;
    const char* des = des_all + pd.b_off;
    char* oa = out_a + pd.out_off;
    char* ob = out_b + pd.out_off;
    const uint32_t n_ck = ck[0].x;
    if(res.status != 0) {
        if(seg == 0 && lane == 0) oa[0] = 0, ob[0] = 0;
        return;
    }
    if(seg + 1 >= n_ck && seg != 0) return;
    const bool body = seg + 1 < n_ck;
    const uint4 hi = body ? ck[1 + seg] : make_uint4(0, 0, 0, 0);  // columns [lo.x, hi.x),
    const uint4 lo = body ? ck[2 + seg] : make_uint4(0, 0, 0, 0);  // sources from (lo.y, lo.z)
    const uint32_t start = res.start, lt = (1u << lane) - 1u;
    uint32_t ia = lo.y, ib = lo.z;
    for(uint32_t base = lo.x; base < hi.x; base += 32) {
        const uint32_t x = base + lane;
        const int op = x < hi.x ? ops[x] : -1;
        const bool useA = op == ST_M || op == ST_D, useB = op == ST_M || op == ST_I;
        const uint32_t ma = __ballot_sync(0xffffffffu, useA), mb = __ballot_sync(0xffffffffu, useB);
        if(x < hi.x) {
            oa[x - start] = useA ? anc[ia + __popc(ma & lt)] : '-';
            ob[x - start] = useB ? des[ib + __popc(mb & lt)] : '-';
        }
        ia += __popc(ma);
        ib += __popc(mb);
    }
    if(seg != 0) return;
    uint32_t n = res.len;
    const bool sa = pd.cfg & CFG_STOP_A, sb = pd.cfg & CFG_STOP_B;
    if(sa || sb) {  // restore_end_stops (utils.cc:1044-1063), see expand_rows_kernel
        if(lane < 3) {
            oa[n + lane] = sa ? anc[pd.la + lane] : '-';
            ob[n + lane] = sb ? des[pd.lb + lane] : '-';
        }
        if(lane == 0) {
            results[pd.orig].len = n + 3;
            if(sa != sb) results[pd.orig].score = res.score + stop_gap;
        }
        n += 3;
    }
    if(lane == 0) oa[n] = 0, ob[n] = 0;
}

// Raw-sequence entry point: marginal_seq_encoding (utils.cc:496-528) on the device, one warp per pair.
// anc -> codon61 * 3 + phase (any symbol outside ACGTUacgtu: E_AMBIGUOUS; an in-frame stop codon:
// E_STOP), des -> IUPAC code (codes 15 / 16, i.e. '-' or anything else: E_SYMBOL, where the reference
// would index past the table).  pd.la / pd.lb are the lengths after end-stop trimming.
__device__ __forceinline__ uint32_t nt16_code(unsigned char ch) {
    switch(ch | 0x20) {  // case-insensitive for letters
    case 'a': return 0;
    case 'c': return 1;
    case 'g': return 2;
    case 't': case 'u': return 3;
    case 'r': return 4;
    case 'y': return 5;
    case 'm': return 6;
    case 'k': return 7;
    case 's': return 8;
    case 'w': return 9;
    case 'b': return 10;
    case 'd': return 11;
    case 'h': return 12;
    case 'v': return 13;
    case 'n': return 14;
    default: return ch == '-' ? 15 : 16;
    }
}

__global__ void encode_pairs_kernel(const PairDesc* __restrict__ pairs, uint32_t npairs,
                                    const char* __restrict__ anc_all, const char* __restrict__ des_all,
                                    uint8_t* __restrict__ a_all, uint8_t* __restrict__ b_all,
                                    PairResult* __restrict__ results,
                                    uint32_t* __restrict__ any_ambiguity_code) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if(warp >= npairs) return;
    const PairDesc pd = pairs[warp];
    if(results[pd.orig].status != 0) return;
    const char* anc = anc_all + pd.a_off;
    const char* des = des_all + pd.b_off;
    uint8_t* a = a_all + pd.a_off;
    uint8_t* b = b_all + pd.b_off;
    // first offending ancestor codon decides the error, as the reference's loop does (utils.cc:504-515)
    uint32_t first_bad = 0xffffffffu;  // (codon index << 2) | 1: ambiguous, | 2: stop
    bool bad_des = false, beyond_acgt = false;
    for(uint32_t cod = lane; cod * 3 < pd.la; cod += 32) {
        const uint32_t n0 = nt16_code(anc[3 * cod]), n1 = nt16_code(anc[3 * cod + 1]),
                       n2 = nt16_code(anc[3 * cod + 2]);
        uint32_t c61 = 0;
        if((n0 | n1 | n2) > 3) {
            first_bad = min(first_bad, (cod << 2) | 1u);
        } else {
            const uint32_t c64 = (n0 << 4) | (n1 << 2) | n2;
            if(c64 == 48 || c64 == 50 || c64 == 56) first_bad = min(first_bad, (cod << 2) | 2u);
            c61 = c64 < 48 ? c64 : c64 == 49 ? 48 : c64 < 57 ? c64 - 2 : c64 - 3;  // utils.cc:1144-1165
        }
        a[3 * cod] = (uint8_t)(3 * c61);
        a[3 * cod + 1] = (uint8_t)(3 * c61 + 1);
        a[3 * cod + 2] = (uint8_t)(3 * c61 + 2);
    }
    for(uint32_t x = lane; x < pd.lb; x += 32) {
        const uint32_t code = nt16_code(des[x]);
        if(code > 14) bad_des = true;
        if(code > 3) beyond_acgt = true;
        b[x] = (uint8_t)(code > 14 ? 0 : code);
    }
    if(__any_sync(0xffffffffu, beyond_acgt) && lane == 0) *any_ambiguity_code = 1u;  // benign race
    for(int o = 16; o > 0; o >>= 1) first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, o));
    bad_des = __any_sync(0xffffffffu, bad_des);
    if(lane == 0) {
        if(first_bad != 0xffffffffu) results[pd.orig].status = (first_bad & 1u) ? -6 : -7;
        else if(bad_des) results[pd.orig].status = -4;
    }
}

}  // namespace coati_gpu
