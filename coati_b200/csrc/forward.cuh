// Forward (log-semiring) fill and seeded stochastic sampleback.
//
//   forward_fill_kernel : forward_impl<semiring::log, align_pair_work_t>, src/lib/align_pair.cc:62-139,
//                         one CTA per pair, anti-diagonal wavefront, stores the three state matrices
//                         M, D, I (12 B/cell) in lattice coordinates (La+1) x (Lb+1), row-major.
//   sampleback_kernel   : sampleback, src/lib/align_pair.cc:401-458.  The reference keeps eight more
//                         matrices with the transition values (align_pair.hpp:94-103); they are pure
//                         functions of the neighbouring state values, so they are recomputed here with
//                         the same float operations instead of being stored (44 -> 12 B/cell).
//                         Margin cells keep the reference's quirk: only del_del / ins_ins are copied
//                         from the margins (align_pair.hpp:108-111), every other transition there is
//                         `lowest`.
// Samples of one pair share one RNG stream whose consumption is path dependent (1 + #steps draws per
// sample), so they are drawn serially by one thread, exactly like the reference; pairs are
// independent (one thread each).
#pragma once

#include "common.cuh"
#include "devmath.cuh"

namespace coati_gpu {

struct FwdDesc {
    uint64_t a_off, b_off;   // into a_all / b_all (and anc_all / des_all)
    uint64_t mat_off;        // into the matrix arena, in floats; M, D, I are consecutive planes
    uint32_t la, lb;
};

__device__ __forceinline__ float lse3(float x, float y, float z) {  // semiring.hpp:68-71
    return log_sum_exp(log_sum_exp(x, y), z);
}

// term: adjusted terminal M, D, I per pair (align_pair.cc:130-138); the matrices keep the
// un-adjusted values at (La, Lb).
__global__ void __launch_bounds__(1023)
forward_fill_kernel(const FwdDesc* __restrict__ pairs, uint32_t npairs,
                    const uint8_t* __restrict__ a_all, const uint8_t* __restrict__ b_all,
                    const float* __restrict__ table, GapConsts g, float* __restrict__ mats,
                    float* __restrict__ term) {
    __shared__ float s_table[TABLE_ROWS * TABLE_LD];
    for(int x = threadIdx.x; x < TABLE_ROWS * TABLE_LD; x += blockDim.x) s_table[x] = table[x];
    __syncthreads();
    for(uint32_t p = blockIdx.x; p < npairs; p += gridDim.x) {
        const FwdDesc pd = pairs[p];
        const uint32_t la = pd.la, lb = pd.lb, k = g.k, ld = lb + 1;
        const uint64_t plane = (uint64_t)(la + 1) * ld;
        float* M = mats + pd.mat_off;
        float* D = M + plane;
        float* I = D + plane;
        const uint8_t* a = a_all + pd.a_off;
        const uint8_t* b = b_all + pd.b_off;
        // three threads per cell: the M, D and I updates are independent log_sum_exp chains
        const uint32_t sub = threadIdx.x % 3, slot = threadIdx.x / 3, nslots = blockDim.x / 3;
        for(uint32_t d = 0; d <= la + lb; ++d) {
            const uint32_t rlo = d > lb ? d - lb : 0, rhi = d < la ? d : la;
            for(uint32_t r = rlo + slot; r <= rhi && slot < nslots; r += nslots) {
                const uint32_t c = d - r;
                const uint64_t at = (uint64_t)r * ld + c;
                float v = LOWEST;
                if(r == 0 || c == 0) {  // align_pair.cc:82-90
                    if(sub == ST_M) { if(r == 0 && c == 0) v = 0.0f; }
                    else if(sub == ST_D) { if(c == 0 && r > 0 && r % k == 0) v = (g.ng + g.go) + g.ge * (float)(r + k - 2); }
                    else { if(r == 0 && c > 0 && c % k == 0) v = g.go + g.ge * (float)(c + k - 2); }
                } else if(sub == ST_M) {
                    const float s = s_table[a[r - 1] * TABLE_LD + b[c - 1]];
                    const uint64_t dg = (uint64_t)(r - 1) * ld + (c - 1);
                    const float m2m = ((M[dg] + g.ng) + g.ng) + s;
                    const float d2m = (D[dg] + g.gs) + s;
                    const float i2m = ((I[dg] + g.gs) + g.ng) + s;
                    v = lse3(m2m, d2m, i2m);  // :119 plus(mch2mch, del2mch, ins2mch)
                } else if(sub == ST_D) {
                    float uM = LOWEST, uD = LOWEST, uI = LOWEST;
                    if(r >= k) {
                        const uint64_t up = (uint64_t)(r - k) * ld + c;
                        uM = M[up], uD = D[up], uI = I[up];
                    }
                    const float m2d = ((uM + g.ng) + g.go) + g.gk1;
                    const float i2d = ((uI + g.gs) + g.go) + g.gk1;
                    const float d2d = uD + g.gk;
                    v = lse3(m2d, d2d, i2d);  // :120 plus(mch2del, del2del, ins2del)
                } else {
                    float lM = LOWEST, lI = LOWEST;
                    if(c >= k) {
                        const uint64_t lf = (uint64_t)r * ld + (c - k);
                        lM = M[lf], lI = I[lf];
                    }
                    const float m2i = (lM + g.go) + g.gk1;
                    const float i2i = lI + g.gk;
                    v = log_sum_exp(m2i, i2i);  // :121
                }
                (sub == ST_M ? M : sub == ST_D ? D : I)[at] = v;
                if(r == la && c == lb)  // :130-138
                    term[3 * p + sub] = sub == ST_M ? (v + g.ng) + g.ng : sub == ST_D ? v + g.gs : (v + g.gs) + g.ng;
            }
            __threadfence_block();
            __syncthreads();
        }
    }
}

// ---- RNG: fragmites Lehmer64Fast, contrib/random/random.hpp:80-136, 213-216 -----------------------
struct Lehmer {
    uint64_t lo, hi;
    __device__ __forceinline__ uint64_t bits() {  // state *= MULT (mod 2^128); return high 64 bits
        const uint64_t MULT = 0xda942042e4dd58b5ull;
        const uint64_t nlo = lo * MULT;
        hi = hi * MULT + __umul64hi(lo, MULT);
        lo = nlo;
        return hi;
    }
    __device__ __forceinline__ float f24() {
        const long long n = (long long)(bits() >> 40);
        return __fdiv_rn((float)n, 16777216.0f);
    }
};

struct Pick {
    int st;
    float logp;
};
// align_pair.cc:336-357
__device__ __forceinline__ Pick sample_mdi(float lm, float ld, float li, float p) {
    const float m = libm_expf(lm), d = libm_expf(ld), n = libm_expf(li);
    const float scale = __fadd_rn(__fadd_rn(m, d), n);
    p = __fmul_rn(p, scale);
    Pick r;
    if(p < m) r.st = ST_M, r.logp = lm;
    else if(p < __fadd_rn(d, m)) r.st = ST_D, r.logp = ld;
    else r.st = ST_I, r.logp = li;
    r.logp = __fsub_rn(r.logp, libm_logf(scale));
    return r;
}
// align_pair.cc:369-385
__device__ __forceinline__ Pick sample_mi(float lm, float li, float p) {
    const float m = libm_expf(lm), n = libm_expf(li);
    const float scale = __fadd_rn(m, n);
    p = __fmul_rn(p, scale);
    Pick r;
    if(p < m) r.st = ST_M, r.logp = lm;
    else r.st = ST_I, r.logp = li;
    r.logp = __fsub_rn(r.logp, libm_logf(scale));
    return r;
}

// One thread per pair; nsamples sequential samplebacks on a shared RNG stream.
// rng: 2 x uint64 per pair {lo, hi}, in-out.  Rows are written right-aligned into the sample's slot
// (stride la + lb + 1) and left-aligned afterwards by compact_samples_kernel.
__global__ void sampleback_kernel(const FwdDesc* __restrict__ pairs, uint32_t npairs,
                                  const float* __restrict__ mats, const float* __restrict__ term,
                                  const float* __restrict__ table, const uint8_t* __restrict__ a_all,
                                  const uint8_t* __restrict__ b_all, const char* __restrict__ anc_all,
                                  const char* __restrict__ des_all, GapConsts g,
                                  uint64_t* __restrict__ rng, uint32_t nsamples,
                                  const uint64_t* __restrict__ out_off, char* __restrict__ out_a,
                                  char* __restrict__ out_b, uint32_t* __restrict__ out_len,
                                  uint32_t* __restrict__ out_start, float* __restrict__ scores,
                                  int32_t* __restrict__ status) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= npairs) return;
    const FwdDesc pd = pairs[p];
    const uint32_t la = pd.la, lb = pd.lb, k = g.k, ld = lb + 1;
    const uint64_t plane = (uint64_t)(la + 1) * ld;
    const float* M = mats + pd.mat_off;
    const float* D = M + plane;
    const float* I = D + plane;
    const uint8_t* a = a_all + pd.a_off;
    const uint8_t* b = b_all + pd.b_off;
    const char* anc = anc_all + pd.a_off;
    const char* des = des_all + pd.b_off;
    Lehmer rs{rng[2 * p], rng[2 * p + 1]};
    const float tM = term[3 * p], tD = term[3 * p + 1], tI = term[3 * p + 2];
    const uint32_t stride = la + lb + 1;
    int err = 0;
    for(uint32_t s = 0; s < nsamples && !err; ++s) {
        char* oa = out_a + out_off[p] + (uint64_t)s * stride;
        char* ob = out_b + out_off[p] + (uint64_t)s * stride;
        uint32_t r = la, c = lb, pos = la + lb;
        float score = 0.0f;
        float w = fmaxf(fmaxf(tM, tD), tI);  // :414
        Pick pick = sample_mdi(tM - w, tD - w, tI - w, rs.f24());
        score = __fadd_rn(score, pick.logp);
        while(r > 0 || c > 0) {  // :419
            const bool terminal = r == la && c == lb;
            const bool margin = r == 0 || c == 0;
            const uint64_t at = (uint64_t)r * ld + c;
            if(pick.st == ST_M) {
                if(r == 0 || c == 0) { err = 1; break; }
                --pos;
                oa[pos] = anc[r - 1];
                ob[pos] = des[c - 1];
                w = terminal ? tM : M[at];
                const uint64_t dg = (uint64_t)(r - 1) * ld + (c - 1);
                const float sb = table[a[r - 1] * TABLE_LD + b[c - 1]];
                const float mm = ((M[dg] + g.ng) + g.ng) + sb;
                const float dm = (D[dg] + g.gs) + sb;
                const float im = ((I[dg] + g.gs) + g.ng) + sb;
                pick = sample_mdi(mm - w, dm - w, im - w, rs.f24());
                score = __fadd_rn(score, pick.logp);
                --r, --c;
            } else if(pick.st == ST_D) {
                if(r < k) { err = 1; break; }
                for(uint32_t q = 0; q < k; ++q) {
                    --pos;
                    oa[pos] = anc[r - 1 - q];
                    ob[pos] = '-';
                }
                w = terminal ? tD : D[at];
                float md = LOWEST, dd2 = LOWEST, id = LOWEST;
                if(margin) {
                    dd2 = D[at];  // init_margins(): del_del = del on the margins
                } else {
                    float uM = LOWEST, uD = LOWEST, uI = LOWEST;
                    if(r >= k) {
                        const uint64_t up = (uint64_t)(r - k) * ld + c;
                        uM = M[up], uD = D[up], uI = I[up];
                    }
                    md = ((uM + g.ng) + g.go) + g.gk1;
                    id = ((uI + g.gs) + g.go) + g.gk1;
                    dd2 = uD + g.gk;
                }
                pick = sample_mdi(md - w, dd2 - w, id - w, rs.f24());
                score = __fadd_rn(score, pick.logp);
                r -= k;
            } else {
                if(c < k) { err = 1; break; }
                for(uint32_t q = 0; q < k; ++q) {
                    --pos;
                    oa[pos] = '-';
                    ob[pos] = des[c - 1 - q];
                }
                w = terminal ? tI : I[at];
                float mi = LOWEST, ii2 = LOWEST;
                if(margin) {
                    ii2 = I[at];  // init_margins(): ins_ins = ins on the margins
                } else {
                    float lM = LOWEST, lI = LOWEST;
                    if(c >= k) {
                        const uint64_t lf = (uint64_t)r * ld + (c - k);
                        lM = M[lf], lI = I[lf];
                    }
                    mi = (lM + g.go) + g.gk1;
                    ii2 = lI + g.gk;
                }
                pick = sample_mi(mi - w, ii2 - w, rs.f24());
                score = __fadd_rn(score, pick.logp);
                c -= k;
            }
        }
        const uint64_t so = (uint64_t)p * nsamples + s;
        out_len[so] = err ? 0 : la + lb - pos;
        out_start[so] = err ? la + lb : pos;
        scores[so] = score;
    }
    status[p] = err ? -8 : 0;
    rng[2 * p] = rs.lo;
    rng[2 * p + 1] = rs.hi;
}

// One warp per (pair, sample): left-align the rows and NUL-terminate (src >= dst, chunked copy).
__global__ void compact_samples_kernel(const FwdDesc* __restrict__ pairs, uint32_t npairs,
                                       uint32_t nsamples, const uint64_t* __restrict__ out_off,
                                       char* __restrict__ out_a, char* __restrict__ out_b,
                                       const uint32_t* __restrict__ out_len,
                                       const uint32_t* __restrict__ out_start) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if(warp >= (uint64_t)npairs * nsamples) return;
    const uint32_t p = warp / nsamples, s = warp % nsamples;
    const FwdDesc pd = pairs[p];
    const uint32_t stride = pd.la + pd.lb + 1;
    char* oa = out_a + out_off[p] + (uint64_t)s * stride;
    char* ob = out_b + out_off[p] + (uint64_t)s * stride;
    const uint32_t n = out_len[warp], shift = out_start[warp];
    if(shift != 0) {
        for(uint32_t base = 0; base < n; base += 32) {
            const uint32_t x = base + lane;
            char va = 0, vb = 0;
            if(x < n) va = oa[shift + x], vb = ob[shift + x];
            __syncwarp();
            if(x < n) oa[x] = va, ob[x] = vb;
            __syncwarp();
        }
    }
    if(lane == 0) oa[n] = 0, ob[n] = 0;
}

// test hook: evaluate one of the libm twins over an array
__global__ void libm_eval_kernel(int op, const float* __restrict__ in, float* __restrict__ out,
                                 uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float x = in[i];
    out[i] = op == 0 ? libm_expf(x) : op == 1 ? libm_logf(x) : op == 2 ? libm_log1pf(x) : op == 3 ? log1p_exp(x)
                                                                                         : log1p_exp_neg(x);
}

}  // namespace coati_gpu
