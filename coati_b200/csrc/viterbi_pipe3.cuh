// K = 3 (codon-unit gaps, `-k 3`) specialisation of the register-pipelined inter-pair Viterbi fill:
// the scheme, the decision-plane stream (PipeLayout) and the exactness argument of viterbi_pipe.cuh,
// with the issue-slot tuning of viterbi_pipe1.cuh (packed add.rn.f32x2 for the match/insert halves of
// row pairs, sign-shift decisions, lane 31 carrying lane 0's inputs).
//
// What K = 3 changes (src/lib/align_pair.cc:97-124 with look_back = 3):
//   D(r, c) comes from row r-3: the three bottom rows of a lane go to the lane below (3 shuffles),
//   I(r, c) comes from column c-3: a 3-deep history of Z per row,
//   the fill adds ge*(k-1) / ge*k (S::power) where traceback compares with +ge only
//   (align_pair.cc:285-296), so the decision maxima and the fill maxima are separate.
#pragma once

#include <type_traits>

#include "common.cuh"
#include "viterbi_pipe.cuh"
#include "viterbi_pipe1.cuh"

namespace coati_gpu {

// Sign-shift form of a row pair (see COATI_ROWPAIR_SGN in viterbi_pipe1.cuh): decisions are the sign
// bits of packed subtractions against the decision maxima X and Yd.
#define COATI_ROWPAIR3_SGN(q, ZS)                                                               \
    {                                                                                         \
        const f2 M2 = mk2(Mv[q], Mv[q + 1]);                                                  \
        const f2 I2 = mk2(Zh[q][ZS], Zh[q + 1][ZS]); /* Z of three columns back */            \
        const f2 t1 = add2(M2, ng2), xm = add2(t1, ng2), ym = add2(t1, go2), zm = add2(M2, go2); \
        const f2 t2 = add2(I2, gs2), xi = add2(t2, ng2), yi = add2(t2, go2), zi = add2(I2, ge2); \
        const f2 zmk = add2(zm, gk12), ik = add2(I2, gk2);                                    \
        const float D0 = (q) < 3 ? recvY[(q) < 3 ? (q) : 0] : Ycur[(q) >= 3 ? (q)-3 : 0];     \
        const float D1 = (q) + 1 < 3 ? recvY[(q) + 1 < 3 ? (q) + 1 : 0] : Ycur[(q) + 1 >= 3 ? (q)-2 : 0]; \
        const f2 Dp = mk2(D0, D1);                                                            \
        const f2 xd = add2(Dp, gs2), yd = add2(Dp, ge2), dk = add2(Dp, gk2);                  \
        const float X0 = fmaxf(fmaxf(lo2(xm), lo2(xd)), lo2(xi));                             \
        const float X1 = fmaxf(fmaxf(hi2(xm), hi2(xd)), hi2(xi));                             \
        const float Y0 = fmaxf(fmaxf(lo2(ym), lo2(yd)), lo2(yi));                             \
        const float Y1 = fmaxf(fmaxf(hi2(ym), hi2(yd)), hi2(yi));                             \
        const f2 X2 = mk2(X0, X1), Y2 = mk2(Y0, Y1);                                          \
        const f2 d0 = sub2(xm, X2), d1 = sub2(xd, X2), d2 = sub2(ym, Y2), d3 = sub2(yd, Y2);  \
        const f2 d4 = sub2(zi, zm);                                                           \
        push_sign(acc[q][0], lo2(d0)), push_sign(acc[q + 1][0], hi2(d0));                     \
        push_sign(acc[q][1], lo2(d1)), push_sign(acc[q + 1][1], hi2(d1));                     \
        push_sign(acc[q][2], lo2(d2)), push_sign(acc[q + 1][2], hi2(d2));                     \
        push_sign(acc[q][3], lo2(d3)), push_sign(acc[q + 1][3], hi2(d3));                     \
        push_sign(acc[q][4], lo2(d4)), push_sign(acc[q + 1][4], hi2(d4));                     \
        Xp[q] = X0, Xp[q + 1] = X1;                                                           \
        /* fill maxima (align_pair.cc:106-118): max(max(ym, yi) + gk1, D + gk) */             \
        const f2 ymi = add2(mk2(fmaxf(lo2(ym), lo2(yi)), fmaxf(hi2(ym), hi2(yi))), gk12);     \
        Ycur[q] = fmaxf(lo2(ymi), lo2(dk)), Ycur[q + 1] = fmaxf(hi2(ymi), hi2(dk));           \
        Zh[q][ZS] = fmaxf(lo2(zmk), lo2(ik)), Zh[q + 1][ZS] = fmaxf(hi2(zmk), hi2(ik));       \
    }

template <int R, int NC>
__global__ void __maxnreg__(COATI_PIPE1_REGS)  // see viterbi_pipe1.cuh
viterbi_pipe3_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                     unsigned int* __restrict__ counter, const uint8_t* __restrict__ a_all,
                     const uint8_t* __restrict__ b_all, const float* __restrict__ table, GapConsts g,
                     float4* __restrict__ bnd_all, uint32_t bnd_stride, uint8_t* __restrict__ dirs,
                     PairResult* __restrict__ results, const unsigned int* __restrict__ nc_flag) {
    static_assert(R % 6 == 0, "rows are processed in pairs and handed over in threes");
    constexpr int K = 3;
    constexpr int R4 = (R + 3) / 4;
    constexpr int H = 32 * R;
    constexpr uint32_t WPL = (5 * R + 3) & ~3u;
    extern __shared__ float4 s_dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* s_tab = s_dyn + (size_t)warp * R4 * NC * 32;
    float4* bnd = bnd_all + ((size_t)blockIdx.x * PIPE_WARPS + warp) * 2 * bnd_stride;
    const uint32_t FULL = 0xffffffffu;
    const int rot = (lane + 31) & 31;
    const f2 ng2 = mk2(g.ng, g.ng), go2 = mk2(g.go, g.go), gs2 = mk2(g.gs, g.gs), ge2 = mk2(g.ge, g.ge);
    const f2 gk12 = mk2(g.gk1, g.gk1), gk2 = mk2(g.gk, g.gk);
    const char* tab_lane = reinterpret_cast<const char*>(s_tab) + lane * 16;
    if(nc_flag && ((*nc_flag != 0) != (NC == 16))) return;  // see viterbi_pipe1.cuh

    for(;;) {
        uint32_t p = 0;
        if(lane == 0) p = first + atomicAdd(counter, 1u);
        p = __shfl_sync(FULL, p, 0);
        if(p >= last) break;
        const PairDesc pd = pairs[p];
        if(results[pd.orig].status != 0 || pd.la == 0 || pd.lb == 0) continue;
        const uint32_t la = pd.la, lb = pd.lb;
        const float* tab = table + (size_t)(pd.cfg >> CFG_MODEL_SHIFT) * (TABLE_ROWS * TABLE_LD);
        const uint8_t* a = a_all + pd.a_off;
        const uint8_t* b = b_all + pd.b_off;
        uint4* dir = reinterpret_cast<uint4*>(dirs + pd.dir_off);
        const uint32_t nblocks = pipe_nblocks(lb, R);
        const uint32_t nbands = (la + H - 1) / H;
        const uint32_t nsteps = lb + 31;

        // rows above band 0: top margin row r = 0 (align_pair.cc:88-90) and two padding rows
        for(uint32_t c = 1 + lane; c <= lb; c += 32) {
            const CellOut o = cell_out<K>(LOWEST, LOWEST, margin_ins<K>(c, g), g);
            bnd[c] = make_float4(o.X, LOWEST, LOWEST, o.Y);
        }
        __syncwarp();

        for(uint32_t band = 0; band < nbands; ++band) {
            const float4* bin = bnd + (size_t)(band & 1) * bnd_stride;
            float4* bout = bnd + (size_t)((band + 1) & 1) * bnd_stride;
            const uint32_t r0 = band * H + lane * R + 1;
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
            float Xp[R], Zh[R][K], Ycur[R], diagX, recvY[K];
            uint32_t acc[R][5];
#pragma unroll
            for(int q = 0; q < R; ++q) {
                Xp[q] = margin_del<K>(r0 + q, g) + g.gs;
                Ycur[q] = 0.f;
#pragma unroll
                for(int z = 0; z < K; ++z) Zh[q][z] = LOWEST;
#pragma unroll
                for(int j = 0; j < 5; ++j) acc[q][j] = 0;
            }
            diagX = r0 == 1 ? (0.0f + g.ng) + g.ng : margin_del<K>(r0 - 1, g) + g.gs;
            float outX = 0.f, outY0 = 0.f, outY1 = 0.f, outY2 = 0.f;
            uint32_t boff = 0;
            if(lane == 31) {
                const float4 v = bin[1];
                outX = v.x, outY0 = v.y, outY1 = v.z, outY2 = v.w;
                boff = (uint32_t)b[0] * 512u;
            }
            uint32_t u = 0u - (uint32_t)lane;
            __syncwarp();

            // Z(r, c) is read three columns later: the history is a ring indexed by step % 3 (a lane's
            // steps are consecutive), fixed at compile time by unrolling the step loop three times
            // The decision words hold BS = 30 steps, so a flush block is ten whole triples.
            constexpr uint32_t BS = pipe_block_steps(R);
            static_assert(BS % 3 == 0, "a flush block is a whole number of ring turns");
            const float4* pbin = bin + 2;  // lane 0's inputs for the NEXT step (column t + 2); the row above
            const uint8_t* pb = b + 1;     // and the symbols are padded past column lb
            auto step = [&](auto zs_c) {
                constexpr int ZS = decltype(zs_c)::value;
                const float recvX = __shfl_sync(FULL, outX, rot);
                recvY[0] = __shfl_sync(FULL, outY0, rot);
                recvY[1] = __shfl_sync(FULL, outY1, rot);
                recvY[2] = __shfl_sync(FULL, outY2, rot);
                const uint32_t bo = __shfl_sync(FULL, boff, rot);
                const float4 bnv = *pbin++;
                const uint32_t bl = *pb++;
                if(u < lb) {
                    float sv[R4 * 4];
#pragma unroll
                    for(int h = 0; h < R4; ++h) {
                        const float4 v = *reinterpret_cast<const float4*>(tab_lane + bo + h * (NC * 512));
                        sv[4 * h] = v.x, sv[4 * h + 1] = v.y, sv[4 * h + 2] = v.z, sv[4 * h + 3] = v.w;
                    }
                    // every match score first, from the previous column's X; D of a row is this
                    // column's Y three rows up (rows are evaluated top down)
                    float Mv[R];
                    Mv[0] = diagX + sv[0];
#pragma unroll
                    for(int q = 1; q < R; ++q) Mv[q] = Xp[q - 1] + sv[q];
#pragma unroll
                    for(int q = 0; q < R; q += 2) COATI_ROWPAIR3_SGN(q, ZS)
                    outX = Xp[R - 1];
                    outY0 = Ycur[R - 3], outY1 = Ycur[R - 2], outY2 = Ycur[R - 1];
                    diagX = recvX;
                    if(lane == 31) bout[u + 1] = make_float4(outX, outY0, outY1, outY2);
                }
                boff = bo;
                if(lane == 31) {
                    outX = bnv.x, outY0 = bnv.y, outY1 = bnv.z, outY2 = bnv.w;
                    boff = bl * 512u;
                }
                ++u;
            };
            // The same step for blocks in which all 32 lanes are inside the lattice (viterbi_pipe1.cuh): no activity
            // test, every lane loads its own symbol one step ahead, pointers with immediate offsets, a whole ring
            // turn (three steps) to a basic block.
            const uint8_t* ps = b;
            const float4* pw = bin;
            float4* pst = bout;
            uint32_t off = 0;
            auto fstep = [&](auto zs_c, int i) {
                constexpr int ZS = decltype(zs_c)::value;
                const float recvX = __shfl_sync(FULL, outX, rot);
                recvY[0] = __shfl_sync(FULL, outY0, rot);
                recvY[1] = __shfl_sync(FULL, outY1, rot);
                recvY[2] = __shfl_sync(FULL, outY2, rot);
                const float4 bnv = pw[i];
                const uint32_t s1 = ld_symbol_now(ps + i + 1);
                float sv[R4 * 4];
#pragma unroll
                for(int h = 0; h < R4; ++h) {
                    const float4 v = *reinterpret_cast<const float4*>(tab_lane + off + h * (NC * 512));
                    sv[4 * h] = v.x, sv[4 * h + 1] = v.y, sv[4 * h + 2] = v.z, sv[4 * h + 3] = v.w;
                }
                float Mv[R];
                Mv[0] = diagX + sv[0];
#pragma unroll
                for(int q = 1; q < R; ++q) Mv[q] = Xp[q - 1] + sv[q];
#pragma unroll
                for(int q = 0; q < R; q += 2) COATI_ROWPAIR3_SGN(q, ZS)
                diagX = recvX;
                if(lane == 31) pst[i] = make_float4(Xp[R - 1], Ycur[R - 3], Ycur[R - 2], Ycur[R - 1]);
                outX = lane == 31 ? bnv.x : Xp[R - 1];
                outY0 = lane == 31 ? bnv.y : Ycur[R - 3];
                outY1 = lane == 31 ? bnv.z : Ycur[R - 2];
                outY2 = lane == 31 ? bnv.w : Ycur[R - 1];
                boff = off;  // what the lane below receives on the next step, should an edge block follow
                off = s1 * 512u;
            };
            for(uint32_t t0 = 0; t0 < nsteps; t0 += BS) {
                const uint32_t tn = min(BS, nsteps - t0);
                if(t0 >= 32u && t0 + BS + 1 <= lb) {
                    ps = b + t0 - lane;       // ps[i]: this lane's symbol on step t0 + i
                    pw = bin + t0 + 2;        // pw[i]: lane 0's inputs for step t0 + i + 1
                    pst = bout + t0 + 1 - lane;  // pst[i]: lane 31's cell of step t0 + i
                    off = ld_symbol_now(ps) * 512u;
#pragma unroll 1
                    for(uint32_t tt = 0; tt < BS; tt += 3, ps += 3, pw += 3, pst += 3) {
                        fstep(std::integral_constant<int, 0>{}, 0);
                        fstep(std::integral_constant<int, 1>{}, 1);
                        fstep(std::integral_constant<int, 2>{}, 2);
                    }
                    if(lane == 31) boff = (uint32_t)b[t0 + BS] * 512u;  // lane 31 hands lane 0 the symbol of ITS next step
                    u += BS, pbin += BS, pb += BS;
                } else
                for(uint32_t tt = 0; tt < tn; tt += 3) {
                    step(std::integral_constant<int, 0>{});
                    if(tt + 1 < tn) step(std::integral_constant<int, 1>{});
                    if(tt + 2 < tn) step(std::integral_constant<int, 2>{});
                }
                // ---- flush the block of decision planes (see viterbi_pipe1.cuh: align the pushed bits,
                // complement planes 0-3)
                const uint32_t t = t0 + tn - 1;  // last step of the block
                uint4* dst = dir + ((size_t)(band * nblocks + t0 / BS) * 32 + lane) * (WPL / 4);
                uint32_t w[WPL];
#pragma unroll
                for(int x = 0; x < (int)WPL; ++x) w[x] = x < 5 * R ? acc[x / 5][x % 5] : 0u;
                const uint32_t t_end = min(t, lb - 1 + (uint32_t)lane);
                const uint32_t sh = 31u - (t_end % BS);
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
            __syncwarp();
        }
    }
}

#undef COATI_ROWPAIR3_SGN

}  // namespace coati_gpu
