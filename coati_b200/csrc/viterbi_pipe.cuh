// Register-pipelined Viterbi fill for gap unit length K = 1 and K = 3 (inter-pair scheme).
//
// One warp per pair.  The lattice is cut into bands of H = 32*R rows; inside a band lane l owns
// R consecutive rows and sweeps the columns left to right, one column per step, skewed by one
// step per lane (lane l is at column t - l + 1 on step t), so a warp-step is one anti-diagonal
// of R-row register tiles.  Per row a lane keeps only
//      X(r,c) = max3((M+ng)+ng, D+gs, (I+gs)+ng)     -> M(r+1,c+1) = X + subst
//      Y(r,c) = D(r+K,c)   and   Z(r,c) = I(r,c+K)
// which is an exact refactoring of forward_impl (src/lib/align_pair.cc:94-129): rounding is
// monotone, so max(x+s, y+s, z+s) == max(x,y,z)+s bit for bit, and every addition below is one
// of the reference's own, in its left-to-right order (no FMA: built with -fmad=false).
// The bottom rows of a lane go to the lane below by shuffle, the bottom rows of a band go to the
// next band through a per-warp global scratch row.
//
// Output: the decisions traceback<S> (align_pair.cc:268-299) would take at each cell, as five
// bit-planes per row (see PipeLayout), 0.625-0.7 B/cell, flushed with 128-bit stores every 32
// steps.  The Viterbi score and the initial traceback state are X and its planes at (La, Lb),
// because max3 of the adjusted terminal scores (align_pair.cc:130-138, 265-266) IS X(La, Lb).
#pragma once

#include "common.cuh"

namespace coati_gpu {

// ---- direction-stream layout of the pipelined kernels -------------------------------------------
// planes: 0: xm == X (MATCH lands -> M)   1: xd == X (-> D, if plane 0 clear; else I)
//         2: ym == Y (DELETION lands)     3: yd == Y            4: zm > zi (INSERTION lands -> M)
// word index = ((band * nblocks + step/BS) * 32 + lane) * WPL + q * 5 + plane, bit 31 - step%BS,
// step = (c - 1) + lane, lane = ((r-1) % H) / R, q = (r-1) % R.  BS = steps per word: 32, or 30 for the
// K = 3 configurations (R = 3, 6), whose step loop is unrolled three times (viterbi_pipe3.cuh) and
// wants whole triples between two flushes.
__host__ __device__ __forceinline__ uint32_t pipe_wpl(uint32_t R) { return (5 * R + 3) & ~3u; }
__host__ __device__ constexpr uint32_t pipe_block_steps(uint32_t R) { return R % 3 == 0 ? 30u : 32u; }
__host__ __device__ __forceinline__ uint32_t pipe_nblocks(uint32_t lb, uint32_t R) {
    return (lb + 31 + pipe_block_steps(R) - 1) / pipe_block_steps(R);
}
__host__ __device__ __forceinline__ uint64_t pipe_dir_bytes(uint32_t la, uint32_t lb, uint32_t R) {
    const uint64_t nbands = (la + 32 * R - 1) / (32 * R);
    return nbands * pipe_nblocks(lb, R) * 32ull * pipe_wpl(R) * 4ull;
}

template <int K>
__device__ __forceinline__ float margin_del(uint32_t r, const GapConsts& g) {  // D(r, 0), r > 0
    return (r % K == 0) ? (g.ng + g.go) + g.ge * (float)(r + K - 2) : LOWEST;
}
template <int K>
__device__ __forceinline__ float margin_ins(uint32_t c, const GapConsts& g) {  // I(0, c), c > 0
    return (c % K == 0) ? g.go + g.ge * (float)(c + K - 2) : LOWEST;
}

// Everything later cells need from a cell's (M, D, I); also returns the five decision masks
// (all-ones / zero).
struct CellOut {
    float X, Y, Z;
    uint32_t e1, e2, f1, f2, zz;
};

template <int K>
__device__ __forceinline__ CellOut cell_out(float M, float D, float I, const GapConsts& g) {
    CellOut o;
    const float t1 = M + g.ng;
    const float xm = t1 + g.ng;
    const float ym = t1 + g.go;
    const float zm = M + g.go;
    const float xd = D + g.gs;
    const float yd = D + g.ge;
    const float t2 = I + g.gs;
    const float xi = t2 + g.ng;
    const float yi = t2 + g.go;
    const float zi = I + g.ge;
    o.X = fmaxf(fmaxf(xm, xd), xi);
    const float Yd = fmaxf(fmaxf(ym, yd), yi);  // traceback's comparison values (align_pair.cc:285-287)
    o.e1 = (xm == o.X) ? 0xffffffffu : 0u;
    o.e2 = (xd == o.X) ? 0xffffffffu : 0u;
    o.f1 = (ym == Yd) ? 0xffffffffu : 0u;
    o.f2 = (yd == Yd) ? 0xffffffffu : 0u;
    o.zz = (zm > zi) ? 0xffffffffu : 0u;
    if(K == 1) {
        o.Y = Yd;               // gk1 == -0.0f, gk == ge: the fill's terms are the same floats
        o.Z = fmaxf(zm, zi);
    } else {
        // fill terms (align_pair.cc:106-118): ((.)+go)+gk1, D+gk, (M+go)+gk1, I+gk
        o.Y = fmaxf(fmaxf(ym, yi) + g.gk1, D + g.gk);
        o.Z = fmaxf(zm + g.gk1, I + g.gk);
    }
    return o;
}

constexpr int PIPE_WARPS = 4;  // warps per CTA

// bnd: per-warp scratch, 2 buffers of (bnd_stride) float4: {X(rb,c), Y(rb-2,c), Y(rb-1,c), Y(rb,c)}
// for K = 3; K = 1 uses .x and .w only.
template <int K, int R>
__global__ void __launch_bounds__(PIPE_WARPS * 32)
viterbi_pipe_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                    unsigned int* __restrict__ counter, const uint8_t* __restrict__ a_all,
                    const uint8_t* __restrict__ b_all, const float* __restrict__ table, GapConsts g,
                    float4* __restrict__ bnd_all, uint32_t bnd_stride, uint8_t* __restrict__ dirs,
                    PairResult* __restrict__ results) {
    static_assert(K == 1 || (K == 3 && R % 3 == 0), "K = 3 needs R % 3 == 0");
    constexpr int R4 = (R + 3) / 4;
    constexpr int H = 32 * R;
    constexpr uint32_t WPL = (5 * R + 3) & ~3u;
    // private substitution rows: s_tab[warp][h][nuc][lane] (float4 = rows 4h..4h+3 of the lane)
    extern __shared__ float4 s_dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* s_tab = s_dyn + (size_t)warp * R4 * 16 * 32;
    float4* bnd = bnd_all + ((size_t)blockIdx.x * PIPE_WARPS + warp) * 2 * bnd_stride;
    const uint32_t FULL = 0xffffffffu;

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
        constexpr uint32_t BS = pipe_block_steps(R);
        const uint32_t nbands = (la + H - 1) / H;
        const uint32_t nsteps = lb + 31;

        // boundary row above band 0 = top margin row r = 0 (align_pair.cc:88-90)
        for(uint32_t c = 1 + lane; c <= lb; c += 32) {
            const CellOut o = cell_out<K>(LOWEST, LOWEST, margin_ins<K>(c, g), g);
            // rows -2, -1 are padding: their Y is lowest
            bnd[c] = make_float4(o.X, LOWEST, LOWEST, o.Y);
        }
        __syncwarp();

        float score = 0.f;
        for(uint32_t band = 0; band < nbands; ++band) {
            float4* bin = bnd + (band & 1) * bnd_stride;
            float4* bout = bnd + ((band + 1) & 1) * bnd_stride;
            const uint32_t r0 = band * H + lane * R + 1;  // first row of this lane
            // ---- private substitution rows --------------------------------------------------
#pragma unroll
            for(int h = 0; h < R4; ++h) {
                float rowv[4][16];
#pragma unroll
                for(int x = 0; x < 4; ++x) {
                    const uint32_t r = r0 + 4 * h + x;
                    const bool ok = (4 * h + x < R) && r <= la;
                    const uint32_t code = ok ? a[r - 1] : 0;
#pragma unroll
                    for(int n = 0; n < 16; ++n) rowv[x][n] = ok ? tab[code * TABLE_LD + n] : 0.0f;
                }
#pragma unroll
                for(int n = 0; n < 16; ++n)
                    s_tab[(h * 16 + n) * 32 + lane] = make_float4(rowv[0][n], rowv[1][n], rowv[2][n], rowv[3][n]);
            }
            // ---- state at column 0 (left margin, align_pair.cc:84-87) -------------------------
            float Xp[R], Zh[R][K], diagX, recvY[K];
            uint32_t acc[R][5];
#pragma unroll
            for(int q = 0; q < R; ++q) {
                const uint32_t r = r0 + q;
                Xp[q] = margin_del<K>(r, g) + g.gs;  // X(r, 0): only D is finite
#pragma unroll
                for(int z = 0; z < K; ++z) Zh[q][z] = LOWEST;  // Z(r, c <= 0)
#pragma unroll
                for(int j = 0; j < 5; ++j) acc[q][j] = 0;
            }
            {
                const uint32_t r = r0 - 1;  // row above the lane
                diagX = r == 0 ? (0.0f + g.ng) + g.ng : margin_del<K>(r, g) + g.gs;
            }
            float outX = 0.f, outY[K];
#pragma unroll
            for(int z = 0; z < K; ++z) outY[z] = 0.f;
            float4 bnext = lb >= 1 ? bin[1] : make_float4(0, 0, 0, 0);  // lane 31 feeds lane 0
            uint32_t bcode = lane == 0 ? b[0] : 0;
            __syncwarp();

            for(uint32_t t = 0; t < nsteps; ++t) {
                // ---- uniform part: neighbour exchange -----------------------------------------
                // lane 31 injects the boundary row above the band for lane 0's column (t + 1)
                float sx = outX, sy[K];
#pragma unroll
                for(int z = 0; z < K; ++z) sy[z] = outY[z];
                if(lane == 31) {
                    sx = bnext.x;
                    if(K == 3) {
                        sy[0] = bnext.y;
                        sy[1] = bnext.z;
                    }
                    sy[K - 1] = bnext.w;
                }
                const float recvX = __shfl_sync(FULL, sx, (lane + 31) & 31);
#pragma unroll
                for(int z = 0; z < K; ++z) recvY[z] = __shfl_sync(FULL, sy[z], (lane + 31) & 31);
                if(lane == 31 && t + 2 <= lb) bnext = bin[t + 2];
                const uint32_t c = t - lane + 1;  // unsigned wrap => inactive
                const bool active = c >= 1 && c <= lb;
                uint32_t bn = 0;
                if(c + 1 >= 1 && c + 1 <= lb) bn = b[c];  // next step's symbol (c + 1)
                if(active) {
                    // ---- R cells of column c ------------------------------------------------
                    const uint32_t bm = 1u << (31 - (t % BS));
                    float sv[R4 * 4];
#pragma unroll
                    for(int h = 0; h < R4; ++h) {
                        const float4 v = s_tab[(h * 16 + bcode) * 32 + lane];
                        sv[4 * h] = v.x, sv[4 * h + 1] = v.y, sv[4 * h + 2] = v.z, sv[4 * h + 3] = v.w;
                    }
                    float dX = diagX, Ycur[R];
#pragma unroll
                    for(int q = 0; q < R; ++q) {
                        const float M = dX + sv[q];
                        const float D = q < K ? recvY[q] : Ycur[q - K];
                        const float I = Zh[q][K - 1];
                        const CellOut o = cell_out<K>(M, D, I, g);
                        dX = Xp[q];
                        Xp[q] = o.X;
                        Ycur[q] = o.Y;
#pragma unroll
                        for(int z = K - 1; z > 0; --z) Zh[q][z] = Zh[q][z - 1];
                        Zh[q][0] = o.Z;
                        acc[q][0] |= o.e1 & bm;
                        acc[q][1] |= o.e2 & bm;
                        acc[q][2] |= o.f1 & bm;
                        acc[q][3] |= o.f2 & bm;
                        acc[q][4] |= o.zz & bm;
                    }
                    outX = Xp[R - 1];
#pragma unroll
                    for(int z = 0; z < K; ++z) outY[z] = Ycur[R - K + z];
                    diagX = recvX;
                    if(lane == 31) {
                        if(K == 1) bout[c] = make_float4(outX, 0.f, 0.f, outY[0]);
                        else bout[c] = make_float4(outX, outY[0], outY[1], outY[K - 1]);
                    }
                }
                bcode = bn;
                // ---- flush the 32-step block of decision planes ---------------------------------
                if(t % BS == BS - 1 || t == nsteps - 1) {
                    uint4* dst = dir + ((size_t)(band * nblocks + t / BS) * 32 + lane) * (WPL / 4);
                    uint32_t w[WPL];
#pragma unroll
                    for(int x = 0; x < (int)WPL; ++x) w[x] = x < 5 * R ? acc[x / 5][x % 5] : 0u;
#pragma unroll
                    for(int x = 0; x < (int)WPL / 4; ++x)
                        dst[x] = make_uint4(w[4 * x], w[4 * x + 1], w[4 * x + 2], w[4 * x + 3]);
#pragma unroll
                    for(int q = 0; q < R; ++q)
#pragma unroll
                        for(int j = 0; j < 5; ++j) acc[q][j] = 0;
                }
            }
            // Viterbi score = X(La, Lb), held by the lane/row that owns row La after its last step
            if(band == nbands - 1) {
                const uint32_t rr = (la - 1) % H;
                if((uint32_t)lane == rr / R) {
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
