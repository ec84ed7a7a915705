// Generic-k Viterbi fill: one CTA per pair, anti-diagonal wavefront, scores in a rolling ring of
// anti-diagonals.  Handles ANY gap unit length k (the reference accepts any `-k`, utils.cc:135);
// the register-pipelined kernels in viterbi_pipe.cuh cover the documented k = 1 and k = 3 cases.
//
// Recurrence: forward_impl<tropical, align_pair_work_mem_t>, src/lib/align_pair.cc:81-138.
#pragma once

#include "common.cuh"

namespace coati_gpu {

// Ring depth: a cell on diagonal d = r + c reads diagonals d-2 (match) and d-k (gaps).
__host__ __device__ inline uint32_t ring_depth(uint32_t k) { return (k > 2 ? k : 2) + 1; }

// margin values, align_pair.cc:82-90 with i = r + k - 1 (raw matrix index quirk kept: SURVEY fact 6)
__device__ __forceinline__ void margin_cell(uint32_t r, uint32_t c, const GapConsts& g, float& M,
                                            float& D, float& I) {
    M = D = I = LOWEST;
    if(r == 0 && c == 0) {
        M = 0.0f;  // S::one()
    } else if(c == 0) {
        if(r % g.k == 0) D = (g.ng + g.go) + g.ge * (float)(size_t)(r + g.k - 2);
    } else {
        if(c % g.k == 0) I = g.go + g.ge * (float)(size_t)(c + g.k - 2);
    }
}

// Work distribution: CTAs pull sorted pair indices [first, last) from an atomic counter.
// ring: per-CTA scratch of 3 * ring_depth(k) * (max_la + 1) floats in global memory.
__global__ void __launch_bounds__(256)
viterbi_generic_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                       unsigned int* __restrict__ counter, const uint8_t* __restrict__ a_all,
                       const uint8_t* __restrict__ b_all, const float* __restrict__ table,
                       GapConsts g, float* __restrict__ ring_all, uint32_t ring_stride,
                       uint8_t* __restrict__ dirs, PairResult* __restrict__ results) {
    __shared__ float s_table[TABLE_ROWS * TABLE_LD];
    __shared__ unsigned int s_pair;
    const uint32_t depth = ring_depth(g.k);
    float* ring = ring_all + (size_t)blockIdx.x * 3 * depth * ring_stride;
    uint32_t cached_model = 0xffffffffu;

    for(;;) {
        if(threadIdx.x == 0) s_pair = first + atomicAdd(counter, 1u);
        __syncthreads();
        const uint32_t p = s_pair;
        __syncthreads();
        if(p >= last) break;
        const PairDesc pd = pairs[p];
        if(results[pd.orig].status != 0) continue;
        if((pd.cfg >> CFG_MODEL_SHIFT) != cached_model) {  // substitution table of this pair's model
            cached_model = pd.cfg >> CFG_MODEL_SHIFT;
            const float* tab = table + (size_t)cached_model * (TABLE_ROWS * TABLE_LD);
            for(int x = threadIdx.x; x < TABLE_ROWS * TABLE_LD; x += blockDim.x) s_table[x] = tab[x];
            __syncthreads();
        }
        const uint32_t la = pd.la, lb = pd.lb, k = g.k;
        const uint8_t* a = a_all + pd.a_off;
        const uint8_t* b = b_all + pd.b_off;
        uint8_t* dir = dirs + pd.dir_off;
        auto slotM = [&](uint32_t d) { return ring + (size_t)(d % depth) * 3 * ring_stride; };

        for(uint32_t d = 0; d <= la + lb; ++d) {
            const uint32_t rlo = d > lb ? d - lb : 0, rhi = d < la ? d : la;
            float* curM = slotM(d);
            float* curD = curM + ring_stride;
            float* curI = curD + ring_stride;
            const float* m2 = d >= 2 ? slotM(d - 2) : nullptr;
            const float* mk = d >= k ? slotM(d - k) : nullptr;
            for(uint32_t r = rlo + threadIdx.x; r <= rhi; r += blockDim.x) {
                const uint32_t c = d - r;
                float M, D, I;
                if(r == 0 || c == 0) {
                    margin_cell(r, c, g, M, D, I);
                } else {
                    const float s = s_table[a[r - 1] * TABLE_LD + b[c - 1]];
                    // (r-1, c-1)
                    const float pM = m2[r - 1], pD = m2[ring_stride + r - 1],
                                pI = m2[2 * ring_stride + r - 1];
                    const float m2m = ((pM + g.ng) + g.ng) + s;
                    const float d2m = (pD + g.gs) + s;
                    const float i2m = ((pI + g.gs) + g.ng) + s;
                    // (r-k, c)
                    float uM = LOWEST, uD = LOWEST, uI = LOWEST;
                    if(r >= k) {
                        uM = mk[r - k];
                        uD = mk[ring_stride + r - k];
                        uI = mk[2 * ring_stride + r - k];
                    }
                    const float m2d = ((uM + g.ng) + g.go) + g.gk1;
                    const float i2d = ((uI + g.gs) + g.go) + g.gk1;
                    const float d2d = uD + g.gk;
                    // (r, c-k)
                    float lM = LOWEST, lI = LOWEST;
                    if(c >= k) {
                        lM = mk[r];
                        lI = mk[2 * ring_stride + r];
                    }
                    const float m2i = (lM + g.go) + g.gk1;
                    const float i2i = lI + g.gk;
                    M = fmaxf(fmaxf(m2m, d2m), i2m);
                    D = fmaxf(fmaxf(m2d, d2d), i2d);
                    I = fmaxf(m2i, i2i);
                    if(r == la && c == lb) {  // terminal: never consulted by traceback
                        dir[dir_index_diag(r, c, la, lb)] = 0;
                    } else {
                        dir[dir_index_diag(r, c, la, lb)] = direction_byte(M, D, I, g);
                    }
                }
                curM[r] = M;
                curD[r] = D;
                curI[r] = I;
                if(r == la && c == lb) {  // align_pair.cc:130-138 terminal adjustment
                    PairResult& res = results[pd.orig];
                    res.term[0] = (M + g.ng) + g.ng;
                    res.term[1] = D + g.gs;
                    res.term[2] = (I + g.gs) + g.ng;
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace coati_gpu
