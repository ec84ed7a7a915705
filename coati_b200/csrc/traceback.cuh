// Traceback over the packed direction stream + left-alignment of the emitted rows.
//
// Follows traceback<S>, src/lib/align_pair.cc:249-303: start at the terminal cell with
// max_mdi(M, D, I) of the adjusted terminal scores, then walk MATCH (-1,-1) / DELETION (-k, 0) /
// INSERTION (0, -k) until (0, 0); the next state is the decision byte of the cell just landed on
// (common.cuh: direction_byte), or implied on the margins where only one state is finite.
#pragma once

#include "common.cuh"

namespace coati_gpu {

// Generic-k kernels: one decision byte per body cell, anti-diagonal-major (common.cuh).
struct DiagLayout {
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

// Pipelined kernels: five bit-planes per row (viterbi_pipe.cuh).  pd.cfg = R.
struct PipeLayout {
    __device__ __forceinline__ static uint32_t plane_bit(const uint32_t* w, const PairDesc& pd,
                                                         uint32_t r, uint32_t c, uint32_t plane) {
        const uint32_t R = pd.cfg, H = 32 * R, wpl = (5 * R + 3) & ~3u;
        const uint32_t band = (r - 1) / H, rr = (r - 1) % H, lane = rr / R, q = rr % R;
        const uint32_t t = (c - 1) + lane, nblocks = (pd.lb + 62) / 32;
        const uint64_t idx = ((uint64_t)(band * nblocks + (t >> 5)) * 32 + lane) * wpl + q * 5 + plane;
        return (w[idx] >> (31 - (t & 31))) & 1u;
    }
    __device__ __forceinline__ static int next(const uint8_t* dir, const PairDesc& pd, int st,
                                               uint32_t r, uint32_t c) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(dir);
        if(st == ST_I) return plane_bit(w, pd, r, c, 4) ? ST_M : ST_I;
        const uint32_t base = st == ST_M ? 0 : 2;
        if(plane_bit(w, pd, r, c, base)) return ST_M;
        return plane_bit(w, pd, r, c, base + 1) ? ST_D : ST_I;
    }
    // score = X(La, Lb) was written by the fill; max_mdi of the adjusted terminal scores is the
    // MATCH-lands decision of the terminal cell (align_pair.cc:130-138, 265-266).
    __device__ __forceinline__ static int initial(const uint8_t* dir, const PairDesc& pd,
                                                  PairResult&) {
        return next(dir, pd, ST_M, pd.la, pd.lb);
    }
};

// One thread per pair.  Rows are written right-aligned into the pair's output slot
// [out_off, out_off + la + lb]; compact_rows_kernel moves them to the front afterwards.
template <class Layout>
__global__ void traceback_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                                 const uint8_t* __restrict__ dirs, const char* __restrict__ anc_all,
                                 const char* __restrict__ des_all, GapConsts gap,
                                 char* __restrict__ out_a, char* __restrict__ out_b,
                                 PairResult* __restrict__ results) {
    const uint32_t p = first + blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= last) return;
    const PairDesc pd = pairs[p];
    PairResult& res = results[pd.orig];
    if(res.status != 0) return;
    const uint32_t la = pd.la, lb = pd.lb, k = gap.k;
    const char* anc = anc_all + pd.a_off;
    const char* des = des_all + pd.b_off;
    const uint8_t* dir = dirs + pd.dir_off;
    char* oa = out_a + pd.out_off;
    char* ob = out_b + pd.out_off;

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
    while(r > 0 || c > 0) {  // :268  (j > k-1 || i > k-1)
        if(st == ST_M) {
            if(r == 0 || c == 0) { err = 1; break; }
            --pos;
            oa[pos] = anc[r - 1];
            ob[pos] = des[c - 1];
            --r, --c;
        } else if(st == ST_D) {
            if(r < k) { err = 1; break; }
            for(uint32_t q = 0; q < k; ++q) {
                --pos;
                oa[pos] = anc[r - 1 - q];
                ob[pos] = '-';
            }
            r -= k;
        } else {
            if(c < k) { err = 1; break; }
            for(uint32_t q = 0; q < k; ++q) {
                --pos;
                oa[pos] = '-';
                ob[pos] = des[c - 1 - q];
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
    if(err) {
        res.status = -8;  // COATI_GPU_E_INTERNAL
        res.len = 0;
        res.start = la + lb;
        return;
    }
    res.len = la + lb - pos;
    res.start = pos;
}

// One warp per pair: move the right-aligned rows to the start of the slot and NUL-terminate.
// Forward chunked copy is safe for overlapping ranges because src >= dst.
__global__ void compact_rows_kernel(const PairDesc* __restrict__ pairs, uint32_t first,
                                    uint32_t last, char* __restrict__ out_a,
                                    char* __restrict__ out_b,
                                    const PairResult* __restrict__ results) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t p = first + warp;
    if(p >= last) return;
    const PairDesc pd = pairs[p];
    const PairResult res = results[pd.orig];
    char* oa = out_a + pd.out_off;
    char* ob = out_b + pd.out_off;
    const uint32_t n = res.status == 0 ? res.len : 0, shift = res.start;
    if(shift != 0) {
        for(uint32_t base = 0; base < n; base += 32) {
            const uint32_t x = base + lane;
            char va = 0, vb = 0;
            if(x < n) {
                va = oa[shift + x];
                vb = ob[shift + x];
            }
            __syncwarp();
            if(x < n) {
                oa[x] = va;
                ob[x] = vb;
            }
            __syncwarp();
        }
    }
    if(lane == 0) {
        oa[n] = 0;
        ob[n] = 0;
    }
}

}  // namespace coati_gpu
