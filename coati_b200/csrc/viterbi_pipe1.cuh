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
#include "rowpair1.cuh"
#include "viterbi_pipe.cuh"

namespace coati_gpu {

#ifndef COATI_PIPE1_UNROLL
#define COATI_PIPE1_UNROLL 4
#endif
constexpr int PIPE1_UNROLL = COATI_PIPE1_UNROLL;  // steps per basic block in the interior blocks

// Inter-pair scheme: one warp per pair, the bands of a pair processed one after another by the same warp (pairs
// [first, last) pulled from `counter`).  The intra-pair wavefront for long single pairs is viterbi_wave1.cuh.
// NC = substitution-table columns kept per lane: 16 (all IUPAC codes) or 4 when no descendant of the
// batch carries an ambiguity code (the common case) -- a quarter of the shared memory, so more
// resident warps to fill issue slots.
// nc_flag (raw-sequence batches): device word set by encode_pairs_kernel when any descendant carries an
// ambiguity code; both NC variants are launched and the one that does not apply returns at once, so the
// host never waits for the flag.
// Register budget: 128 = four CTAs per SM (the four-step block would take 143 and drop to three).  Four CTAs of
// 128 x 128 registers fill the register file, and the fills are persistent, so the short kernels of the neighbouring
// sub-batches (traceback, expansion, encode) get onto an SM only as CTAs of a fill retire; with four pipeline lanes
// that costs nothing measurable end to end (1 M pairs: 319-320 ms at 128 registers, 314-330 ms at 120, fills alone
// 303 / 309 ms), so the faster fill is kept.
#ifndef COATI_PIPE1_REGS
#define COATI_PIPE1_REGS 128
#endif
template <int R, int NC>
__global__ void __maxnreg__(COATI_PIPE1_REGS)
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
        uint32_t p = first;
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

        // row above band 0 = top margin row r = 0 (align_pair.cc:88-90)
        for(uint32_t c = 1 + lane; c <= lb; c += 32) {
            const CellOut o = cell_out<1>(LOWEST, LOWEST, margin_ins<1>(c, g), g);
            bnd[c] = make_float2(o.X, o.Y);
        }
        __syncwarp();

        for(uint32_t band = 0; band < nbands; ++band) {
            const float2* bin = bnd + (size_t)(band & 1) * 2 * bnd_stride;
            float2* bout = bnd + (size_t)((band + 1) & 1) * 2 * bnd_stride;
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
            if(lane == 31) {
                const float2 v = bin[1];
                outX = v.x, outY = v.y;
                boff = (uint32_t)b[0] * 512u;
            }
            uint32_t u = 0u - (uint32_t)lane;  // u = t - lane = c - 1
            __syncwarp();

            // blocks of 32 steps = one word of every decision plane; the flush sits between blocks
            for(uint32_t t0 = 0; t0 < nsteps; t0 += 32) {
              const uint32_t t1 = min(t0 + 32u, nsteps);
              if(PIPE1_UNROLL > 0 && t0 >= 32u && t0 + 33u <= lb) {
                // ---- all 32 lanes are inside the lattice for the whole block: no activity test, every lane loads
                // its own symbol (one step ahead), pointers run with immediate offsets, four steps to a basic block
                const uint8_t* ps = b + t0 - lane;          // ps[i]: this lane's symbol on step t0 + i
                const float2* pw = bin + t0 + 2;            // pw[i]: lane 0's inputs for step t0 + i + 1
                float2* pst = bout + t0 + 1 - lane;         // pst[i]: lane 31's cell of step t0 + i
                uint32_t off = ld_symbol_now(ps) * 512u;
#pragma unroll 1
                for(int grp = 0; grp < 32 / (PIPE1_UNROLL > 0 ? PIPE1_UNROLL : 1); ++grp, ps += PIPE1_UNROLL, pw += PIPE1_UNROLL, pst += PIPE1_UNROLL) {
#pragma unroll
                    for(int i = 0; i < PIPE1_UNROLL; ++i) {
                        const float recvX = __shfl_sync(FULL, outX, rot);
                        const float recvY = __shfl_sync(FULL, outY, rot);
                        const float2 bnv = pw[i];
                        const uint32_t s1 = ld_symbol_now(ps + i + 1);
                        float sv[R4 * 4];
#pragma unroll
                        for(int h = 0; h < R4; ++h) {
                            const float4 v = *reinterpret_cast<const float4*>(tab_lane + off + h * (NC * 512));
                            sv[4 * h] = v.x, sv[4 * h + 1] = v.y, sv[4 * h + 2] = v.z, sv[4 * h + 3] = v.w;
                        }
                        float D = recvY;
                        float Mv[R];
                        Mv[0] = diagX + sv[0];
#pragma unroll
                        for(int q = 1; q < R; ++q) Mv[q] = Xp[q - 1] + sv[q];
#pragma unroll
                        for(int q = 0; q < R; q += 2) COATI_ROWPAIR_SGN(q)
                        diagX = recvX;
                        if(lane == 31) pst[i] = make_float2(Xp[R - 1], D);
                        outX = lane == 31 ? bnv.x : Xp[R - 1];
                        outY = lane == 31 ? bnv.y : D;
                        boff = off;  // what the lane below receives on the next step, should an edge block follow
                        off = s1 * 512u;
                    }
                }
                if(lane == 31) boff = (uint32_t)b[t0 + 32] * 512u;  // lane 31 hands lane 0 the symbol of ITS next step
                u += 32;
              } else {
              const float2* pbin = bin + t0 + 2;  // lane 0's inputs for the NEXT step (column t + 2);
              const uint8_t* pb = b + t0 + 1;     // both arrays are padded past column lb
              for(uint32_t t = t0; t < t1; ++t, ++u) {
                const float recvX = __shfl_sync(FULL, outX, rot);
                const float recvY = __shfl_sync(FULL, outY, rot);
                const uint32_t bo = __shfl_sync(FULL, boff, rot);
                const float2 bnv = *pbin++;  // uniform addresses, written by this warp one band earlier
                const uint32_t bl = *pb++;
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
                    if(lane == 31) bout[u + 1] = make_float2(outX, outY);
                }
                boff = bo;
                if(lane == 31) {
                    outX = bnv.x, outY = bnv.y;
                    boff = bl * 512u;
                }
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
    }
}

}  // namespace coati_gpu
