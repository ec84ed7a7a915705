// Traceback over the packed direction stream + left-alignment of the emitted rows.
//
// Follows traceback<S>, src/lib/align_pair.cc:249-303: start at the terminal cell with
// max_mdi(M, D, I) of the adjusted terminal scores, then walk MATCH (-1,-1) / DELETION (-k, 0) /
// INSERTION (0, -k) until (0, 0); the next state is the decision byte of the cell just landed on
// (common.cuh: direction_byte), or implied on the margins where only one state is finite.
#pragma once

#include "common.cuh"

namespace coati_gpu {

struct DiagLayout {
    __device__ __forceinline__ static uint64_t index(uint32_t r, uint32_t c, uint32_t la,
                                                     uint32_t lb) {
        return dir_index_diag(r, c, la, lb);
    }
};

// One thread per pair.  Rows are written right-aligned into the pair's output slot
// [out_off, out_off + la + lb]; compact_rows_kernel moves them to the front afterwards.
template <class Layout>
__global__ void traceback_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                                 const uint8_t* __restrict__ dirs, const char* __restrict__ anc_all,
                                 const char* __restrict__ des_all, uint32_t k,
                                 char* __restrict__ out_a, char* __restrict__ out_b,
                                 PairResult* __restrict__ results) {
    const uint32_t p = first + blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= last) return;
    const PairDesc pd = pairs[p];
    PairResult& res = results[pd.orig];
    if(res.status != 0) return;
    const uint32_t la = pd.la, lb = pd.lb;
    const char* anc = anc_all + pd.a_off;
    const char* des = des_all + pd.b_off;
    const uint8_t* dir = dirs + pd.dir_off;
    char* oa = out_a + pd.out_off;
    char* ob = out_b + pd.out_off;

    const float tM = res.term[0], tD = res.term[1], tI = res.term[2];
    res.score = fmaxf(fmaxf(tM, tD), tI);  // align_pair.cc:265
    int st = max_mdi(tM, tD, tI);          // :266
    uint32_t r = la, c = lb, pos = la + lb;
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
        else {
            const uint32_t byte = dir[Layout::index(r, c, la, lb)];
            nst = st == ST_M ? (byte & 3) : st == ST_D ? ((byte >> 2) & 3) : ((byte >> 4) & 1) * 2;
        }
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
