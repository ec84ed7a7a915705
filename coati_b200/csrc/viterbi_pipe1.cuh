// K = 1 specialisation of the register-pipelined Viterbi fill (see viterbi_pipe.cuh for the scheme
// and the exactness argument).  Same lattice decomposition, same decision-plane stream (PipeLayout),
// tuned for issue slots on sm_100a:
//   * the match/insert halves of two adjacent rows are evaluated with packed add.rn.f32x2 (FADD2); each
//     lane of a packed add is an IEEE round-to-nearest FADD, so results are bit-identical to the scalar
//     form;
//   * decisions are the sign bits of five packed subtractions per row pair, pushed into the plane
//     accumulators by funnel shifts: 1.5 instructions per decision bit (an FSETP + predicated IMAD
//     form, 2 per bit, was measured in round 1 and removed: profiles/r01_pipe1_fadd2_imad_ncu.txt);
//   * the symbol, the row above the band and the row below it move on uniform addresses
//     (lane 31 carries lane 0's inputs in its outgoing shuffle registers), so the per-step
//     overhead is three shuffles, two broadcast loads and one predicated store;
//   * the step loop runs in blocks of 32 steps (one word of every plane) with the flush between blocks.
#pragma once

#include "common.cuh"
#include "viterbi_pipe.cuh"

namespace coati_gpu {

struct f2 {
    unsigned long long v;
};
__device__ __forceinline__ f2 mk2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(f2 a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return lo;
}
__device__ __forceinline__ float hi2(f2 a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return hi;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {  // two independent round-to-nearest FADDs
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
// acc = (acc << 1) | sign(d): one funnel shift.  For finite a <= b, sign(a - b) is set exactly when
// a != b (a - a is +0 in round-to-nearest), and sign(a - b) is set exactly when b > a.
__device__ __forceinline__ void push_sign(uint32_t& acc, float d) {
    acc = __funnelshift_l(__float_as_uint(d), acc, 1);
}

// Sign-shift form of a row PAIR: the five decisions of both rows are the sign bits of
// five packed subtractions, pushed into the plane accumulators by funnel shifts -- 1.5 instructions per
// decision bit instead of FSETP + predicated IMAD.  Planes 0-3 are accumulated inverted (bit = "differs
// from the maximum") and complemented at the flush.  Every score is finite here (|x| <= FLT_MAX and the
// penalties cannot round LOWEST away from -FLT_MAX), so the differences never produce NaN or -0.
#define COATI_ROWPAIR_SGN(q)                                                                  \
    {                                                                                         \
        const f2 M2 = mk2(Mv[q], Mv[q + 1]);                                                  \
        const f2 I2 = mk2(Zp[q], Zp[q + 1]);                                                  \
        const f2 t1 = add2(M2, ng2), xm = add2(t1, ng2), ym = add2(t1, go2), zm = add2(M2, go2); \
        const f2 t2 = add2(I2, gs2), xi = add2(t2, ng2), yi = add2(t2, go2), zi = add2(I2, ge2); \
        const float xd0 = D + g.gs, yd0 = D + g.ge;                                           \
        const float X0 = fmaxf(fmaxf(lo2(xm), xd0), lo2(xi));                                 \
        const float Y0 = fmaxf(fmaxf(lo2(ym), yd0), lo2(yi));                                 \
        const float xd1 = Y0 + g.gs, yd1 = Y0 + g.ge;                                         \
        const float X1 = fmaxf(fmaxf(hi2(xm), xd1), hi2(xi));                                 \
        const float Y1 = fmaxf(fmaxf(hi2(ym), yd1), hi2(yi));                                 \
        D = Y1;                                                                               \
        const f2 X2 = mk2(X0, X1), Y2 = mk2(Y0, Y1);                                          \
        const f2 d0 = sub2(xm, X2), d1 = sub2(mk2(xd0, xd1), X2);                             \
        const f2 d2 = sub2(ym, Y2), d3 = sub2(mk2(yd0, yd1), Y2), d4 = sub2(zi, zm);          \
        push_sign(acc[q][0], lo2(d0)), push_sign(acc[q + 1][0], hi2(d0));                     \
        push_sign(acc[q][1], lo2(d1)), push_sign(acc[q + 1][1], hi2(d1));                     \
        push_sign(acc[q][2], lo2(d2)), push_sign(acc[q + 1][2], hi2(d2));                     \
        push_sign(acc[q][3], lo2(d3)), push_sign(acc[q + 1][3], hi2(d3));                     \
        push_sign(acc[q][4], lo2(d4)), push_sign(acc[q + 1][4], hi2(d4));                     \
        Xp[q] = X0, Xp[q + 1] = X1;                                                           \
        Zp[q] = fmaxf(lo2(zm), lo2(zi)), Zp[q + 1] = fmaxf(hi2(zm), hi2(zi));                 \
    }

// WAVE = false: inter-pair scheme, one warp per pair, bands of a pair processed one after another by
//                the same warp (pairs [first, last) pulled from `counter`).
// WAVE = true : intra-pair scheme for long pairs: the kernel works on the single pair `first`; every
//                warp of the grid pulls BANDS from `counter`, so the bands of one lattice run
//                concurrently as a systolic wavefront across the whole GPU.  Band b reads the row
//                above it from wave_bnd[b] and writes its bottom row to wave_bnd[b + 1].  The rows
//                are pre-filled with a NaN sentinel and every entry is one aligned 64-bit store, so
//                the data is its own ready flag: the consumer re-reads (L2, relaxed) until the
//                sentinel is gone -- no flags, no fences on the producer's critical path.  Tickets
//                are issued in band order and the grid is fully resident, so a waiting band's
//                producer is always running.
__device__ __forceinline__ float2 ld_relaxed_f2(const float2* p) {
    float2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// The producer's side of that hand-off: one 64-bit relaxed store at gpu scope (a plain weak store racing
// with the relaxed polls would be a data race under the PTX memory model).  The inter-pair scheme reads
// the row back from the same warp after __syncwarp(): a plain store.
template <bool WAVE>
__device__ __forceinline__ void st_boundary(float2* p, float x, float y) {
    if(WAVE) asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
    else *p = make_float2(x, y);
}
// The row above arrives from another SM through L2 (~700 cycles), so it is fetched WAVE_G columns at a time
// (lane j % WAVE_G holds column base + j), one granule ahead of use and WITHOUT waiting: the load is issued
// when the previous granule becomes current and only examined when its own turn comes (re-polled then if the
// producer had not got there).  Round 1 fetched whole 32-column blocks and waited for them a block ahead.
// Measured on B200 (tools/gpu/run16.sh: granule 8 / 16 / 32, sleep 0 / 20 / 40 ns between polls; fill ms at
// 10k / 40k / 160k): R = 4: 2.36-2.49 / 9.8-10.3 / 55-58 against 2.81 / 11.5 / 68.7 before; R = 10: 3.0-3.1 /
// 12.3-12.5 / 49.6-50.6 against 2.92 / 11.75 / 47.2.  The granule hardly matters; 16 with a 20 ns sleep is kept.
#ifndef COATI_WAVE_G
#define COATI_WAVE_G 16
#endif
#ifndef COATI_WAVE_SLEEP
#define COATI_WAVE_SLEEP 20
#endif
constexpr uint32_t WAVE_G = COATI_WAVE_G;
// A symbol of the descendant, loaded NOW into a register that is then kept: with a plain `b[i]` the compiler
// re-loads the byte at the point of use instead (the pointer is const __restrict__), which put an L2 round trip
// on every granule swap of the wavefront (14 % of its time in the first profile of this scheme).
__device__ __forceinline__ uint32_t ld_symbol_now(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// NC = substitution-table columns kept per lane: 16 (all IUPAC codes) or 4 when no descendant of the
// batch carries an ambiguity code (the common case) -- a quarter of the shared memory, so more
// resident warps to fill issue slots.
// nc_flag (raw-sequence batches): device word set by encode_pairs_kernel when any descendant carries an
// ambiguity code; both NC variants are launched and the one that does not apply returns at once, so the
// host never waits for the flag.
template <int R, bool WAVE, int NC>
__global__ void __launch_bounds__(PIPE_WARPS * 32)
viterbi_pipe1_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                     unsigned int* __restrict__ counter, const uint8_t* __restrict__ a_all,
                     const uint8_t* __restrict__ b_all, const float* __restrict__ table, GapConsts g,
                     float4* __restrict__ bnd_all, uint32_t bnd_stride, uint8_t* __restrict__ dirs,
                     PairResult* __restrict__ results, const unsigned int* __restrict__ nc_flag) {
    static_assert(R % 2 == 0, "rows are processed in pairs");
    constexpr int R4 = (R + 3) / 4;
    constexpr int H = 32 * R;
    constexpr uint32_t WPL = (5 * R + 3) & ~3u;
    extern __shared__ float4 s_dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* s_tab = s_dyn + (size_t)warp * R4 * NC * 32;
    float2* bnd = reinterpret_cast<float2*>(bnd_all + ((size_t)blockIdx.x * PIPE_WARPS + warp) * 2 * bnd_stride);
    if(nc_flag && ((*nc_flag != 0) != (NC == 16))) return;
    const uint32_t FULL = 0xffffffffu;
    const int rot = (lane + 31) & 31;
    const f2 ng2 = mk2(g.ng, g.ng), go2 = mk2(g.go, g.go), gs2 = mk2(g.gs, g.gs), ge2 = mk2(g.ge, g.ge);
    const char* tab_lane = reinterpret_cast<const char*>(s_tab) + lane * 16;

    for(;;) {
        uint32_t p = first, band0 = 0;
        if(!WAVE) {
            if(lane == 0) p = first + atomicAdd(counter, 1u);
            p = __shfl_sync(FULL, p, 0);
            if(p >= last) break;
        }
        const PairDesc pd = pairs[p];
        if(results[pd.orig].status != 0 || pd.la == 0 || pd.lb == 0) {
            if(WAVE) break;
            continue;
        }
        const uint32_t la = pd.la, lb = pd.lb;
        const float* tab = table + (size_t)(pd.cfg >> CFG_MODEL_SHIFT) * (TABLE_ROWS * TABLE_LD);
        const uint8_t* a = a_all + pd.a_off;
        const uint8_t* b = b_all + pd.b_off;
        uint4* dir = reinterpret_cast<uint4*>(dirs + pd.dir_off);
        const uint32_t nblocks = pipe_nblocks(lb, R);
        const uint32_t nbands = (la + H - 1) / H;
        const uint32_t nsteps = lb + 31;
        if(WAVE) {
            if(lane == 0) band0 = atomicAdd(counter, 1u);
            band0 = __shfl_sync(FULL, band0, 0);
            if(band0 >= nbands) break;
            bnd = reinterpret_cast<float2*>(bnd_all);  // wave_bnd[band] = bnd + band * 2 * bnd_stride
        }

        // row above band 0 = top margin row r = 0 (align_pair.cc:88-90)
        if(!WAVE || band0 == 0) {
            for(uint32_t c = 1 + lane; c <= lb; c += 32) {
                const CellOut o = cell_out<1>(LOWEST, LOWEST, margin_ins<1>(c, g), g);
                bnd[c] = make_float2(o.X, o.Y);
            }
        }
        __syncwarp();

        for(uint32_t band = band0; band < (WAVE ? band0 + 1 : nbands); ++band) {
            const float2* bin = bnd + (size_t)(WAVE ? band : (band & 1)) * 2 * bnd_stride;
            float2* bout = bnd + (size_t)(WAVE ? band + 1 : ((band + 1) & 1)) * 2 * bnd_stride;
            // WAVE: WAVE_G columns of the row above, one per lane (mod WAVE_G); peek issues the load, settle
            // re-polls until the producer's values are there (NaN sentinel: not written yet)
            const uint32_t gl = (uint32_t)lane & (WAVE_G - 1);
            auto peek = [&](uint32_t col0) { return ld_relaxed_f2(bin + min(col0 + gl, lb)); };
            auto settle = [&](float2 v, uint32_t col0) {
                while(__any_sync(FULL, v.x != v.x)) {
                    if(COATI_WAVE_SLEEP) __nanosleep(COATI_WAVE_SLEEP);
                    v = peek(col0);
                }
                return v;
            };
            const uint32_t r0 = band * H + lane * R + 1;  // first row of this lane
            // ---- private substitution rows: s_tab[h][nuc][lane] = rows 4h..4h+3 ---------------
#pragma unroll
            for(int h = 0; h < R4; ++h) {
                float rowv[4][NC];
#pragma unroll
                for(int x = 0; x < 4; ++x) {
                    const uint32_t r = r0 + 4 * h + x;
                    const bool ok = (4 * h + x < R) && r <= la;
                    const uint32_t code = ok ? a[r - 1] : 0;
#pragma unroll
                    for(int n = 0; n < NC; ++n) rowv[x][n] = ok ? tab[code * TABLE_LD + n] : 0.0f;
                }
#pragma unroll
                for(int n = 0; n < NC; ++n)
                    s_tab[(h * NC + n) * 32 + lane] = make_float4(rowv[0][n], rowv[1][n], rowv[2][n], rowv[3][n]);
            }
            // ---- state at column 0 (left margin, align_pair.cc:84-87) -------------------------
            float Xp[R], Zp[R], diagX;
            uint32_t acc[R][5];
#pragma unroll
            for(int q = 0; q < R; ++q) {
                Xp[q] = margin_del<1>(r0 + q, g) + g.gs;  // X(r, 0): only D is finite
                Zp[q] = LOWEST;                           // Z(r, 0)
#pragma unroll
                for(int j = 0; j < 5; ++j) acc[q][j] = 0;
            }
            diagX = r0 == 1 ? (0.0f + g.ng) + g.ng : margin_del<1>(r0 - 1, g) + g.gs;
            // lane 31's outgoing registers carry lane 0's inputs: row above the band + symbol
            float outX = 0.f, outY = 0.f;
            uint32_t boff = 0;
            // WAVE: the row above comes from another SM through L2, far too slow to fetch one column
            // per step on the critical path; keep two 32-column blocks of it (and of the symbols) in
            // registers, one column per lane, refilled a whole block ahead of use.
            float2 blkA = make_float2(0.f, 0.f), blkB = blkA;
            uint32_t symA = 0, symB = 0;
            if(WAVE) {
                const float2 first = settle(peek(1u - gl), 1u - gl);  // every lane: column 1
                blkA = settle(peek(2u), 2u);
                symA = ld_symbol_now(b + min(1u + gl, lb - 1));
                blkB = peek(2u + WAVE_G);
                symB = ld_symbol_now(b + min(1u + WAVE_G + gl, lb - 1));
                if(lane == 31) outX = first.x, outY = first.y, boff = (uint32_t)b[0] * 512u;
            } else if(lane == 31) {
                const float2 v = bin[1];
                outX = v.x, outY = v.y;
                boff = (uint32_t)b[0] * 512u;
            }
            uint32_t u = 0u - (uint32_t)lane;  // u = t - lane = c - 1
            __syncwarp();

            // blocks of 32 steps = one word of every decision plane; the flush sits between blocks
            for(uint32_t t0 = 0; t0 < nsteps; t0 += 32) {
              const uint32_t t1 = min(t0 + 32u, nsteps);
              const float2* pbin = bin + t0 + 2;  // lane 0's inputs for the NEXT step (column t + 2);
              const uint8_t* pb = b + t0 + 1;     // both arrays are padded past column lb
              for(uint32_t t = t0; t < t1; ++t, ++u) {
                if(WAVE && (t & (WAVE_G - 1)) == 0 && t != 0) {  // the next granule becomes current
                    blkA = settle(blkB, t + 2), symA = symB;
                    blkB = peek(t + WAVE_G + 2);
                    symB = ld_symbol_now(b + min(t + WAVE_G + 1 + gl, lb - 1));
                }
                const float recvX = __shfl_sync(FULL, outX, rot);
                const float recvY = __shfl_sync(FULL, outY, rot);
                const uint32_t bo = __shfl_sync(FULL, boff, rot);
                float2 bnv;
                uint32_t bl;
                if(WAVE) {
                    bnv.x = __shfl_sync(FULL, blkA.x, t & (WAVE_G - 1));
                    bnv.y = __shfl_sync(FULL, blkA.y, t & (WAVE_G - 1));
                    bl = __shfl_sync(FULL, symA, t & (WAVE_G - 1));
                } else {  // uniform addresses, written by this warp one band earlier
                    bnv = *pbin++;
                    bl = *pb++;
                }
                if(u < lb) {
                    float sv[R4 * 4];
#pragma unroll
                    for(int h = 0; h < R4; ++h) {
                        const float4 v = *reinterpret_cast<const float4*>(tab_lane + bo + h * (NC * 512));
                        sv[4 * h] = v.x, sv[4 * h + 1] = v.y, sv[4 * h + 2] = v.z, sv[4 * h + 3] = v.w;
                    }
                    float D = recvY;
                    float Mv[R];  // every match score first, from the previous column's X
                    Mv[0] = diagX + sv[0];
#pragma unroll
                    for(int q = 1; q < R; ++q) Mv[q] = Xp[q - 1] + sv[q];
#pragma unroll
                    for(int q = 0; q < R; q += 2) COATI_ROWPAIR_SGN(q)
                    outX = Xp[R - 1];
                    outY = D;
                    diagX = recvX;
                    if(lane == 31) st_boundary<WAVE>(bout + u + 1, outX, outY);
                }
                boff = bo;
                if(lane == 31) {
                    outX = bnv.x, outY = bnv.y;
                    boff = bl * 512u;
                }
              }
              // ---- flush the 32-step block of decision planes ---------------------------------
              {
                    const uint32_t t = t1 - 1;  // last step of the block
                    uint4* dst = dir + ((size_t)(band * nblocks + (t >> 5)) * 32 + lane) * (WPL / 4);
                    uint32_t w[WPL];
#pragma unroll
                    for(int x = 0; x < (int)WPL; ++x) w[x] = x < 5 * R ? acc[x / 5][x % 5] : 0u;
                    // bits were pushed in at the bottom: the lane's last column of this block goes
                    // to bit 31 - step % 32 (a lane that ends inside the block stopped pushing early)
                    const uint32_t t_end = min(t, lb - 1 + (uint32_t)lane);
                    const uint32_t sh = 31u - (t_end & 31u);
#pragma unroll
                    for(int x = 0; x < 5 * R; ++x) w[x] = (x % 5 < 4 ? ~w[x] : w[x]) << sh;
#pragma unroll
                    for(int x = 0; x < (int)WPL / 4; ++x)
                        dst[x] = make_uint4(w[4 * x], w[4 * x + 1], w[4 * x + 2], w[4 * x + 3]);
#pragma unroll
                    for(int q = 0; q < R; ++q)
#pragma unroll
                        for(int j = 0; j < 5; ++j) acc[q][j] = 0;
              }
            }
            // Viterbi score = X(La, Lb): max3 of the adjusted terminal scores (align_pair.cc:130-138,265)
            if(band == nbands - 1) {
                const uint32_t rr = (la - 1) % H;
                if((uint32_t)lane == rr / R) {
                    float score = 0.f;
#pragma unroll
                    for(int q = 0; q < R; ++q)
                        if((uint32_t)q == rr % R) score = Xp[q];
                    results[pd.orig].score = score;
                }
            }
            __syncwarp();  // bout of this band is bin of the next
        }
        if(WAVE) continue;  // next band ticket
    }
}

#undef COATI_ROWPAIR_SGN

}  // namespace coati_gpu
