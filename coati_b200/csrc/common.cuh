// Shared device/host definitions for the marginal Gotoh kernels (sm_100a only).
//
// Lattice coordinates used everywhere in csrc/: r in [0, La], c in [0, Lb] with r = i - (k-1),
// c = j - (k-1) for the reference's matrix indices (i, j) (align_pair.cc:72-79: matrices are
// (La+k) x (Lb+k), `start = k-1`).  Rows/columns of the reference matrices below `start` are
// padding that only ever holds `lowest`; they are represented here by "r - k < 0 -> LOWEST".
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace coati_gpu {

constexpr float LOWEST = -FLT_MAX;  // semiring.hpp:82-84 zero() = numeric_limits<float>::lowest()
constexpr int TABLE_ROWS = 183;     // 61 sense codons x 3 phases
constexpr int TABLE_COLS = 15;      // IUPAC descendant symbols
constexpr int TABLE_LD = 16;        // device row stride (padded)

// states, numbered as the reference's AlnState (align_pair.cc:200)
enum : int { ST_M = 0, ST_D = 1, ST_I = 2 };

// log-space gap constants, computed on the host with glibc exactly as align_pair.cc:66-69 and
// semiring.hpp:109-111 (power(x, n) = x * float(n)) do.
struct GapConsts {
    float ng;   // log1pf(-g)   "no_gap"
    float gs;   // log1pf(-e)   "gap_stop"
    float go;   // logf(g)      "gap_open"
    float ge;   // logf(e)      "gap_extend"
    float gk1;  // ge * float(k-1)
    float gk;   // ge * float(k)
    uint32_t k;
    float stop_gap;  // logf(g * e * e): restore_end_stops penalty (utils.cc:1049)
};

// One alignment task.  Offsets are in bytes/symbols from the start of the batch arenas.
struct PairDesc {
    uint64_t a_off;    // into a_all / anc_all
    uint64_t b_off;    // into b_all / des_all
    uint64_t dir_off;  // into the chunk's direction buffer
    uint64_t out_off;  // into out_a / out_b (capacity la + lb + 1)
    uint32_t la, lb;
    uint32_t orig;     // index in caller order
    uint32_t cfg;      // 0: generic-k kernel (DiagLayout); else rows per lane R of the pipelined
                       // kernel, | CFG_WAVE when the pair runs as an intra-pair wavefront

};

constexpr uint32_t CFG_WAVE = 0x100u;
// raw-sequence entry point: a terminal stop codon was trimmed from the ancestor / descendant
// (utils.cc:945-967) and is put back by expand_rows_kernel (utils.cc:1044-1063)
constexpr uint32_t CFG_STOP_A = 0x10000u, CFG_STOP_B = 0x20000u;
// substitution-model index of the pair (per-leaf branch lengths of the msa driver, align_msa.cc:285-318)
constexpr uint32_t CFG_MODEL_SHIFT = 18, CFG_MAX_MODELS = 1u << 14;

// Per-pair results (device side, caller order).
struct PairResult {
    float term[3];  // adjusted terminal M, D, I (align_pair.cc:130-138)
    float score;
    uint32_t len;   // alignment columns
    uint32_t start; // first byte of the (right-aligned) rows inside the pair's output slot
    int32_t status;
    uint32_t pad;
};

// argmax with the reference's tie rules (align_pair.cc:210-221): M wins ties over D over I.
__device__ __forceinline__ int max_mdi(float m, float d, float i) {
    int st = ST_M;
    float val = m;
    if(d > val) {
        val = d;
        st = ST_D;
    }
    if(i > val) return ST_I;
    return st;
}
// align_pair.cc:230-232: INSERTION wins ties.
__device__ __forceinline__ int max_mi(float m, float i) { return m > i ? ST_M : ST_I; }

// The byte the fill kernels emit per body cell: the decisions traceback<S> (align_pair.cc:275-296)
// would take when a step LANDS on this cell, one per state it can arrive from.  The additions are
// the reference's own, in its left-to-right association, with no FMA contraction (-fmad=false).
__device__ __forceinline__ uint8_t direction_byte(float M, float D, float I, const GapConsts& g) {
    float mn = M + g.ng, is = I + g.gs;
    int x = max_mdi(mn + g.ng, D + g.gs, is + g.ng);  // after MATCH
    int y = max_mdi(mn + g.go, D + g.ge, is + g.go);  // after DELETION
    int z = max_mi(M + g.go, I + g.ge);               // after INSERTION
    return (uint8_t)(x | (y << 2) | ((z == ST_I ? 1 : 0) << 4));
}

// ---- direction-stream layouts -------------------------------------------------------------------
// LAYOUT_DIAG: body cells (r in [1,La], c in [1,Lb]) stored anti-diagonal-major so that the cells a
// wavefront step produces are contiguous (coalesced byte stores).  Diagonal e = r + c - 2 holds
// rows max(1, e+2-Lb) .. min(La, e+1) in increasing r.
__host__ __device__ __forceinline__ uint64_t diag_offset(uint64_t e, uint64_t la, uint64_t lb) {
    uint64_t m = la < lb ? la : lb, mx = la < lb ? lb : la;
    if(e <= m) return e * (e + 1) / 2;
    if(e <= mx) return m * (m + 1) / 2 + (e - m) * m;
    uint64_t rem = la + lb - 1 - e;  // diagonals e .. la+lb-2 hold rem, rem-1, ..., 1 cells
    return la * lb - rem * (rem + 1) / 2;
}
__host__ __device__ __forceinline__ uint64_t dir_index_diag(uint32_t r, uint32_t c, uint32_t la,
                                                            uint32_t lb) {
    uint64_t e = (uint64_t)r + c - 2;
    uint32_t rlo = (e + 2 > (uint64_t)lb + 1) ? (uint32_t)(e + 2 - lb) : 1u;  // max(1, e+2-lb)
    return diag_offset(e, la, lb) + (r - rlo);
}

}  // namespace coati_gpu
