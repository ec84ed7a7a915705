// Forward (log-semiring) fill for gap unit length 1, as a register pipeline.
//
// forward_impl<semiring::log, align_pair_work_t>, src/lib/align_pair.cc:62-139, with every float
// operation of the reference in its own order (log_sum_exp / log1p_exp of utils.hpp:134-160 through the
// libm twins of devmath.cuh); the three state matrices M, D, I are stored in lattice coordinates
// (La+1) x (Lb+1), row-major, exactly as forward_fill_kernel (forward.cuh, any k) stores them.
//
// A cell costs five log_sum_exp (~100 dependent instructions each), and its M, D and I updates are
// independent chains, so the unit of work is one (row, state): lane 3g + s of a warp owns state s of row
// g of a band of FB_ROWS = 10 rows and sweeps the columns, skewed one step per row, like the Viterbi
// pipeline (viterbi_pipe.cuh).  A warp-step is 10 cells = two log_sum_exp deep (M and D fold three terms,
// I two).  What a lane needs comes from the lanes of the row above (M, D, I of column c for the D update;
// of column c - 1, i.e. what it fetched one step earlier, for the M update) or of its own row (I update)
// by shuffle; the row above the band comes from the stored matrices one column ahead of use.
//
// WAVE = false: one warp per pair (batches): the bands of a pair are swept one after another.
// WAVE = true : one pair, every warp of the grid pulls BANDS from `counter`: all bands run concurrently
//               as a systolic wavefront over the SMs.  The matrices are pre-filled with a NaN sentinel and
//               every value is one relaxed 32-bit store, so a stored value is its own ready flag
//               (viterbi_pipe1.cuh uses the same hand-off): band b + 1 polls row 10(b+1) of the matrices
//               one column ahead and trails band b by about a dozen steps.
#pragma once

#include "common.cuh"
#include "devmath.cuh"
#include "forward.cuh"

namespace coati_gpu {

constexpr int FB_ROWS = 10;   // rows per band: 30 of the 32 lanes carry a (row, state)
constexpr int FB_WARPS = 4;   // warps per CTA

__device__ __forceinline__ float ld_relaxed_f(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
template <bool WAVE>
__device__ __forceinline__ void st_cell(float* p, float v) {
    if(WAVE) asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else *p = v;
}

template <bool WAVE>
__global__ void __launch_bounds__(FB_WARPS * 32)
forward_band_kernel(const FwdDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                    unsigned int* __restrict__ counter, const uint8_t* __restrict__ a_all,
                    const uint8_t* __restrict__ b_all, const float* __restrict__ table, GapConsts g,
                    float* __restrict__ mats, float* __restrict__ term) {
    __shared__ float s_sub[FB_WARPS][FB_ROWS][TABLE_LD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t FULL = 0xffffffffu;
    const int grp = lane / 3, st = lane % 3;  // lanes 30, 31: grp == 10, idle
    const bool has_row = grp < FB_ROWS;
    // whom a lane listens to: the row above (M and D updates) or its own row (I update)
    const int src = 3 * (st == ST_I || grp == 0 ? grp : grp - 1);
    const int s0 = has_row ? src : 0;

    for(;;) {
        uint32_t p = first, band0 = 0;
        if(!WAVE) {
            if(lane == 0) p = first + atomicAdd(counter, 1u);
            p = __shfl_sync(FULL, p, 0);
            if(p >= last) break;
        }
        const FwdDesc pd = pairs[p];
        const uint32_t la = pd.la, lb = pd.lb, ld = lb + 1;
        const uint64_t plane = (uint64_t)(la + 1) * ld;
        float* M = mats + pd.mat_off;
        float* D = M + plane;
        float* I = D + plane;
        float* mine = st == ST_M ? M : st == ST_D ? D : I;
        const uint8_t* a = a_all + pd.a_off;
        const uint8_t* b = b_all + pd.b_off;
        const uint32_t nbands = la == 0 ? 1 : (la + FB_ROWS - 1) / FB_ROWS;
        if(WAVE) {
            if(lane == 0) band0 = atomicAdd(counter, 1u);
            band0 = __shfl_sync(FULL, band0, 0);
            if(band0 >= nbands) break;
        }
        // top margin row (align_pair.cc:82, 88-90)
        if(band0 == 0) {
            for(uint32_t c = lane; c <= lb; c += 32) {
                st_cell<WAVE>(M + c, c == 0 ? 0.0f : LOWEST);
                st_cell<WAVE>(D + c, LOWEST);
                st_cell<WAVE>(I + c, c == 0 ? LOWEST : g.go + g.ge * (float)(c - 1));
            }
            if(la == 0 && lane < 3)  // terminal cell on the margin (:130-138)
                term[3 * p + lane] = lane == ST_M   ? ((lb == 0 ? 0.0f : LOWEST) + g.ng) + g.ng
                                     : lane == ST_D ? LOWEST + g.gs
                                                    : ((lb == 0 ? LOWEST : g.go + g.ge * (float)(lb - 1)) + g.gs) + g.ng;
        }
        __syncwarp();
        if(la == 0) {
            if(WAVE) break;
            continue;
        }

        for(uint32_t band = band0; band < (WAVE ? band0 + 1 : nbands); ++band) {
            const uint32_t r0 = band * FB_ROWS + 1;  // first row of the band
            const uint32_t row = r0 + grp;
            const bool row_ok = has_row && row <= la;
            // substitution scores of the lane's row (15 columns + padding), read by its M lane
            if(has_row) {
                const uint32_t code = row_ok ? a[row - 1] : 0;
                for(int n = st; n < TABLE_LD; n += 3) s_sub[warp][grp][n] = table[code * TABLE_LD + n];
            }
            // left margin (:84-87): D(r, 0) = (ng + go) + ge * (r - 1), M and I lowest
            float v = LOWEST;
            if(row_ok) {
                if(st == ST_D) v = (g.ng + g.go) + g.ge * (float)(row - 1);
                st_cell<WAVE>(mine + (uint64_t)row * ld, v);
            }
            if(lb == 0) {  // only the left margin: the terminal cell is (la, 0)
                if(row_ok && row == la)
                    term[3 * p + st] = st == ST_M ? (v + g.ng) + g.ng : st == ST_D ? v + g.gs : (v + g.gs) + g.ng;
                continue;
            }
            // what the M lane saw one step ago: column 0 of the row it listens to
            float h0, h1, h2;
            {
                const uint32_t ra = row - 1;  // row above
                h0 = ra == 0 ? 0.0f : LOWEST;
                h1 = ra == 0 ? LOWEST : (g.ng + g.go) + g.ge * (float)(ra - 1);
                h2 = LOWEST;
            }
            // the row above the band, one column ahead of use; polled until the producer's value is there
            const float* upM = M + (uint64_t)(r0 - 1) * ld;
            const float* upD = D + (uint64_t)(r0 - 1) * ld;
            const float* upI = I + (uint64_t)(r0 - 1) * ld;
            auto fetch = [&](uint32_t c, float& m, float& d, float& i) {
                c = min(c, lb);
                if(WAVE) {
                    m = ld_relaxed_f(upM + c), d = ld_relaxed_f(upD + c), i = ld_relaxed_f(upI + c);
                } else {
                    m = upM[c], d = upD[c], i = upI[c];
                }
            };
            // Per-lane addends of the three transition terms, so that the update is straight-line code:
            //   A = ((x0 + a1) + a2) + a3   B = (xb + b1) + b2   C = ((x2 + gs) + c2) + c3
            //   M (:98-102):  ((M+ng)+ng)+s   (D+gs)+s       ((I+gs)+ng)+s      from (r-1, c-1)
            //   D (:106-112): ((M+ng)+go)+gk1  D+gk          ((I+gs)+go)+gk1    from (r-1, c)
            //   I (:115-118): (M+go)+gk1       I+gk           -                  from (r, c-1)
            // An absent addend is -0.0f: x + (-0.0f) is x for every x, signed zeros included.
            const float NZ = -0.0f;
            const float a1 = st == ST_I ? g.go : g.ng, a2 = st == ST_M ? g.ng : st == ST_D ? g.go : g.gk1;
            const float a3d = st == ST_D ? g.gk1 : NZ;                 // M: the substitution score
            const float b1 = st == ST_M ? g.gs : g.gk;                 // M: then + s
            const float c2 = st == ST_M ? g.ng : g.go, c3d = g.gk1;    // M: + s
            const bool isM = st == ST_M, isI = st == ST_I;
            float nM[2], nD[2], nI[2];  // row above at the next two columns (two register sets, no copies)
            fetch(1, nM[0], nD[0], nI[0]);
            __syncwarp();
            // the M lane's substitution score, looked up one step ahead of use
            float sub_next = isM && has_row ? s_sub[warp][grp][b[0]] : 0.0f;
            const uint32_t nsteps = lb + FB_ROWS - 1;
            auto step = [&](uint32_t t, float& bM, float& bD, float& bI, float& pM, float& pD, float& pI) {
                if(WAVE) {
                    while(bM != bM || bD != bD || bI != bI) {  // NaN sentinel: not written yet
                        __nanosleep(20);
                        fetch(t + 1, bM, bD, bI);
                    }
                }
                fetch(t + 2, pM, pD, pI);  // for the next step
                float f0 = __shfl_sync(FULL, v, s0);
                float f1 = __shfl_sync(FULL, v, s0 + 1);
                float f2 = __shfl_sync(FULL, v, s0 + 2);
                if(grp == 0 && !isI) f0 = bM, f1 = bD, f2 = bI;
                const uint32_t c = t + 1 - (uint32_t)grp;  // unsigned wrap => inactive
                const float s = sub_next;
                if(isM && has_row) sub_next = s_sub[warp][grp][b[min(c, lb - 1)]];  // column c + 1
                if(row_ok && c >= 1 && c <= lb) {
                    const float x0 = isM ? h0 : f0, x1 = isM ? h1 : f1, x2 = isM ? h2 : f2;
                    const float a3 = isM ? s : a3d, b2 = isM ? s : NZ, c3 = isM ? s : c3d;
                    const float A = ((x0 + a1) + a2) + a3;
                    const float B = ((isI ? x2 : x1) + b1) + b2;
                    const float C = ((x2 + g.gs) + c2) + c3;
                    v = log_sum_exp_fast(A, B);              // :119-121, plus(plus(x, y), z)
                    if(!isI) v = log_sum_exp_fast(v, C);
                    st_cell<WAVE>(mine + (uint64_t)row * ld + c, v);
                    if(row == la && c == lb)  // adjusted terminal values (:130-138); the matrices keep the raw ones
                        term[3 * p + st] = st == ST_M ? (v + g.ng) + g.ng : st == ST_D ? v + g.gs : (v + g.gs) + g.ng;
                }
                h0 = f0, h1 = f1, h2 = f2;
            };
            uint32_t t = 0;
            for(; t + 1 < nsteps; t += 2) {
                step(t, nM[0], nD[0], nI[0], nM[1], nD[1], nI[1]);
                step(t + 1, nM[1], nD[1], nI[1], nM[0], nD[0], nI[0]);
            }
            if(t < nsteps) step(t, nM[0], nD[0], nI[0], nM[1], nD[1], nI[1]);
            __syncwarp();  // the band's bottom row is the next band's row above (same warp when !WAVE)
        }
        if(WAVE) continue;  // next band ticket
    }
}

}  // namespace coati_gpu
