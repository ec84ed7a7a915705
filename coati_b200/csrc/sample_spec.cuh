// Parallel, bit-identical sampleback for many samples of one pair.
//
// The reference draws its N samples serially from ONE RNG stream whose consumption is path dependent
// (1 + #steps draws per sample, align_pair.cc:401-458, align_marginal.cc:589-593), so sample i+1
// cannot start before sample i has finished.  The stream, however, is a multiplicative congruential
// generator (Lehmer64Fast, random.hpp:80-136): draw number s is state0 * MULT^s (mod 2^128), reachable
// directly.  So:
//   1. sample_records_kernel : per lattice cell and state, everything sample_mdi/sample_mi need
//                              (exp'd terms, scale, the three log-probabilities), with the reference's
//                              float operations and the libm twins -- embarrassingly parallel;
//   2. spec_steps_kernel     : for EVERY candidate stream offset s in a window, the number of draws a
//                              sample starting at s consumes (one thread per candidate);
//   3. chase_kernel          : follow next(s) = s + draws(s) from the current offset -> the start
//                              offsets of the real samples;
//   4. sample_paths_kernel   : re-run the real samples in parallel from their offsets, recording ops
//                              and the score (sequential float adds in path order, as the reference);
//   5. expand_samples_kernel : build the gapped rows.
// Every random number, comparison and addition is the one the reference performs, so the output is
// identical sample for sample; only the order of evaluation in time differs.
#pragma once

#include "common.cuh"
#include "devmath.cuh"
#include "forward.cuh"

namespace coati_gpu {

// {m, dm = d + m, scale, lpM} {lpD, lpI, -, -}: next = p < m ? M : p < dm ? D : I with p = u * scale
struct SampleRec {
    float4 a, b;
};

__device__ __forceinline__ SampleRec make_rec3(float lm, float ld, float li) {  // align_pair.cc:336-357
    const float m = libm_expf(lm), d = libm_expf(ld), n = libm_expf(li);
    const float scale = __fadd_rn(__fadd_rn(m, d), n);
    const float ls = libm_logf(scale);
    SampleRec r;
    r.a = make_float4(m, __fadd_rn(d, m), scale, __fsub_rn(lm, ls));
    r.b = make_float4(__fsub_rn(ld, ls), __fsub_rn(li, ls), 0.f, 0.f);
    return r;
}
__device__ __forceinline__ SampleRec make_rec2(float lm, float li) {  // align_pair.cc:369-385
    const float m = libm_expf(lm), n = libm_expf(li);
    const float scale = __fadd_rn(m, n);
    const float ls = libm_logf(scale);
    SampleRec r;
    r.a = make_float4(m, m, scale, __fsub_rn(lm, ls));  // dm = m: the DELETION branch is never taken
    r.b = make_float4(0.f, __fsub_rn(li, ls), 0.f, 0.f);
    return r;
}

// rec[(cell * 3 + state)], cell = r * (lb + 1) + c; rec[3 * ncells] = the initial pick at the terminal
__global__ void sample_records_kernel(FwdDesc pd, const float* __restrict__ mats,
                                      const float* __restrict__ term, const float* __restrict__ table,
                                      const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                      GapConsts g, SampleRec* __restrict__ rec) {
    const uint32_t la = pd.la, lb = pd.lb, k = g.k, ld = lb + 1;
    const uint64_t ncells = (uint64_t)(la + 1) * ld;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float* M = mats;
    const float* D = M + ncells;
    const float* I = D + ncells;
    const float tM = term[0], tD = term[1], tI = term[2];
    if(idx == 0) {
        const float w = fmaxf(fmaxf(tM, tD), tI);  // align_pair.cc:414-416
        rec[3 * ncells] = make_rec3(tM - w, tD - w, tI - w);
    }
    if(idx >= ncells) return;
    const uint32_t r = idx / ld, c = idx % ld;
    const bool terminal = r == la && c == lb, margin = r == 0 || c == 0;
    // MATCH (:422-428)
    {
        float mm = LOWEST, dm = LOWEST, im = LOWEST;
        if(!margin) {
            const uint64_t dg = (uint64_t)(r - 1) * ld + (c - 1);
            const float sb = table[a[r - 1] * TABLE_LD + b[c - 1]];
            mm = ((M[dg] + g.ng) + g.ng) + sb;
            dm = (D[dg] + g.gs) + sb;
            im = ((I[dg] + g.gs) + g.ng) + sb;
        }
        const float w = terminal ? tM : M[idx];
        rec[3 * idx + ST_M] = make_rec3(mm - w, dm - w, im - w);
    }
    // DELETION (:432-441)
    {
        float md = LOWEST, dd = LOWEST, id = LOWEST;
        if(margin) {
            dd = D[idx];
        } else {
            float uM = LOWEST, uD = LOWEST, uI = LOWEST;
            if(r >= k) {
                const uint64_t up = (uint64_t)(r - k) * ld + c;
                uM = M[up], uD = D[up], uI = I[up];
            }
            md = ((uM + g.ng) + g.go) + g.gk1;
            id = ((uI + g.gs) + g.go) + g.gk1;
            dd = uD + g.gk;
        }
        const float w = terminal ? tD : D[idx];
        rec[3 * idx + ST_D] = make_rec3(md - w, dd - w, id - w);
    }
    // INSERTION (:444-452)
    {
        float mi = LOWEST, ii = LOWEST;
        if(margin) {
            ii = I[idx];
        } else {
            float lM = LOWEST, lI = LOWEST;
            if(c >= k) {
                const uint64_t lf = (uint64_t)r * ld + (c - k);
                lM = M[lf], lI = I[lf];
            }
            mi = (lM + g.go) + g.gk1;
            ii = lI + g.gk;
        }
        const float w = terminal ? tI : I[idx];
        rec[3 * idx + ST_I] = make_rec2(mi - w, ii - w);
    }
}

// ---- 128-bit MCG jump-ahead ---------------------------------------------------------------------------
struct U128 {
    uint64_t lo, hi;
};
__host__ __device__ __forceinline__ U128 mul128(U128 x, U128 y) {
#ifdef __CUDA_ARCH__
    const uint64_t hi = __umul64hi(x.lo, y.lo) + x.lo * y.hi + x.hi * y.lo;
#else
    const unsigned __int128 p = (unsigned __int128)x.lo * y.lo;
    const uint64_t hi = (uint64_t)(p >> 64) + x.lo * y.hi + x.hi * y.lo;
#endif
    return U128{x.lo * y.lo, hi};
}
// pw[i] = MULT^(2^i) mod 2^128, i < 64 (filled by the host)
__host__ __device__ __forceinline__ U128 jump(U128 state, uint64_t n, const U128* pw) {
    for(int i = 0; n != 0; ++i, n >>= 1)
        if(n & 1) state = mul128(state, pw[i]);
    return state;
}

// One sample walked from stream offset (already jumped-to) `rs`.  Returns draws consumed, or 0 on a
// walk that leaves the lattice.  WRITE: record ops (right-aligned) and the score.
template <bool WRITE>
__device__ __forceinline__ uint32_t walk_sample(const SampleRec* __restrict__ rec, uint32_t la,
                                                uint32_t lb, uint32_t k, Lehmer rs, char* ops,
                                                uint32_t* out_pos, float* out_score) {
    const uint32_t ld = lb + 1;
    const uint64_t ncells = (uint64_t)(la + 1) * ld;
    uint32_t r = la, c = lb, pos = la + lb, draws = 0;
    float score = 0.0f;
    const SampleRec* q = rec + 3 * ncells;
    int pick;
    for(;;) {
        const float4 ra = __ldg(&q->a);
        const float p = __fmul_rn(rs.f24(), ra.z);
        ++draws;
        const int next = p < ra.x ? ST_M : (p < ra.y ? ST_D : ST_I);
        if(WRITE) {
            const float4 rb = __ldg(&q->b);
            score = __fadd_rn(score, next == ST_M ? ra.w : next == ST_D ? rb.x : rb.y);
        }
        if(draws > 1) {  // complete the move of the state that was current at this cell
            if(pick == ST_M) {
                if(WRITE) ops[--pos] = ST_M;
                --r, --c;
            } else if(pick == ST_D) {
                if(WRITE)
                    for(uint32_t x = 0; x < k; ++x) ops[--pos] = ST_D;
                r -= k;
            } else {
                if(WRITE)
                    for(uint32_t x = 0; x < k; ++x) ops[--pos] = ST_I;
                c -= k;
            }
        }
        pick = next;
        if(r == 0 && c == 0) break;  // align_pair.cc:419 loop condition
        if((pick == ST_M && (r == 0 || c == 0)) || (pick == ST_D && r < k) || (pick == ST_I && c < k)) {
            draws = 0;
            break;
        }
        q = rec + 3 * ((uint64_t)r * ld + c) + pick;
    }
    if(WRITE) {
        *out_pos = pos;
        *out_score = score;
    }
    return draws;
}

__global__ void spec_steps_kernel(const SampleRec* __restrict__ rec, uint32_t la, uint32_t lb, uint32_t k,
                                  U128 state0, const U128* __restrict__ pw, uint64_t base, uint32_t window,
                                  uint32_t* __restrict__ draws_out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if(x >= window) return;
    const U128 st = jump(state0, base + x, pw);
    draws_out[x] = walk_sample<false>(rec, la, lb, k, Lehmer{st.lo, st.hi}, nullptr, nullptr, nullptr);
}

// cursor[0] = current stream offset, cursor[1] = samples placed so far, cursor[2] = error flag
__global__ void chase_kernel(const uint32_t* __restrict__ draws, uint64_t base, uint32_t window, uint64_t n,
                             uint64_t* __restrict__ cursor, uint64_t* __restrict__ starts) {
    uint64_t s = cursor[0], cnt = cursor[1];
    while(cnt < n && s >= base && s < base + window) {
        const uint32_t d = draws[s - base];
        if(d == 0) {
            cursor[2] = 1;
            break;
        }
        starts[cnt++] = s;
        s += d;
    }
    cursor[0] = s;
    cursor[1] = cnt;
}

__global__ void sample_paths_kernel(const SampleRec* __restrict__ rec, uint32_t la, uint32_t lb, uint32_t k,
                                    U128 state0, const U128* __restrict__ pw,
                                    const uint64_t* __restrict__ starts, uint32_t n, char* __restrict__ out_b,
                                    uint32_t* __restrict__ out_len, uint32_t* __restrict__ out_start,
                                    float* __restrict__ scores) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if(s >= n) return;
    const U128 st = jump(state0, starts[s], pw);
    const uint32_t stride = la + lb + 1;
    uint32_t pos = 0;
    float score = 0.f;
    const uint32_t d = walk_sample<true>(rec, la, lb, k, Lehmer{st.lo, st.hi}, out_b + (uint64_t)s * stride, &pos,
                                         &score);
    out_len[s] = d ? la + lb - pos : 0;
    out_start[s] = d ? pos : la + lb;
    scores[s] = score;
}

// warp per sample: ops (right-aligned in the out_b slot) -> gapped rows (see expand_rows_kernel)
__global__ void expand_samples_kernel(uint32_t la, uint32_t lb, uint32_t n, const char* __restrict__ anc,
                                      const char* __restrict__ des, char* __restrict__ out_a,
                                      char* __restrict__ out_b, const uint32_t* __restrict__ out_len,
                                      const uint32_t* __restrict__ out_start) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if(warp >= n) return;
    const uint32_t stride = la + lb + 1;
    char* oa = out_a + (uint64_t)warp * stride;
    char* ob = out_b + (uint64_t)warp * stride;
    const uint32_t len = out_len[warp], shift = out_start[warp], lt = (1u << lane) - 1u;
    uint32_t ia = 0, ib = 0;
    for(uint32_t base = 0; base < len; base += 32) {
        const uint32_t x = base + lane;
        const int op = x < len ? ob[shift + x] : -1;
        const bool useA = op == ST_M || op == ST_D, useB = op == ST_M || op == ST_I;
        const uint32_t ma = __ballot_sync(0xffffffffu, useA), mb = __ballot_sync(0xffffffffu, useB);
        char va = '-', vb = '-';
        if(useA) va = anc[ia + __popc(ma & lt)];
        if(useB) vb = des[ib + __popc(mb & lt)];
        __syncwarp();
        if(x < len) oa[x] = va, ob[x] = vb;
        ia += __popc(ma);
        ib += __popc(mb);
        __syncwarp();
    }
    if(lane == 0) oa[len] = 0, ob[len] = 0;
}

}  // namespace coati_gpu
