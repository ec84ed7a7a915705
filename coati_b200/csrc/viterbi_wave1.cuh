// Intra-pair wavefront of the K = 1 Viterbi fill (BASELINE config 3: single pairs of 10k - 160k nt).
//
// One kernel works on ONE pair; every warp of the grid pulls BANDS of 32 * R rows from a ticket counter, so the
// bands of the lattice run concurrently as a systolic pipeline across the whole GPU.  Band b reads the row above
// it from wave_bnd[b] and writes its bottom row to wave_bnd[b + 1].  The rows are pre-filled with a NaN sentinel
// and every entry is one aligned 64-bit relaxed store at gpu scope, so the data is its own ready flag: no flags,
// no fences on the producer's critical path.  Tickets are issued in band order and the grid is fully resident, so
// a waiting band's producer is always running.
//
// Same lattice decomposition, cell update (rowpair1.cuh) and decision-plane stream (PipeLayout, viterbi_pipe.cuh)
// as the inter-pair fill viterbi_pipe1_kernel, but the step loop is built for the opposite regime.  There, four
// warps share a scheduler and the loop is tuned for issue slots; here a band is ONE warp, about one warp per
// scheduler, and the fill time is
//        (lb + bands * lag) steps  x  the time one warp needs for a step,
// i.e. a latency chain (measured with tools/wave_lag.py, modelled from the SASS control words with
// tools/sass_sched.py).  Round 2's first wavefront shared viterbi_pipe1's loop: 298 cycles per step at R = 4
// (485 at R = 10) and a lag of 85 steps per band.  What this file does about both:
//   * nothing but the recurrence sits on the step's dependent chain.  The symbols do not depend on the fill, so
//     each lane loads its own descendant symbol a group (four steps) ahead (L1) and its substitution scores one
//     step ahead (shared memory); the old loop passed the symbol down the lanes by shuffle and then waited for
//     the LDS.
//   * EVERY lane computes on EVERY step, inside the lattice or not (see `step`): a step is one basic block without
//     an activity test, four of them are scheduled together, and the sign pushes of one step (a third of its
//     instructions, off the chain) fill the shuffle latency of the next.  The first and last blocks of a band --
//     lanes entering and leaving the lattice -- are on the critical path of the whole wavefront and cost three
//     selects per row more, nothing else.
//   * the row above reaches lane 0 through a rolling window: lane j holds column j (mod 32); every group (eight
//     steps; four for the widest tile) as many lanes fetch the columns needed WAVE_D steps later (relaxed L2 loads
//     into registers of their own, not waited for), an earlier fetch enters the window and the vote on the next
//     group's columns is issued a group before it is branched on (re-polled only if the producer has not got there).
//   * one band per scheduler: the host picks the narrowest lane tile whose band count fits 4 x SMs.
// Measured (B200, tools/wave_lag.py): 108 / 163 / 274 / 331 cycles per step at R = 2 / 4 / 8 / 10; fills of the
// sampledata pairs 10k / 40k / 160k: 1.13 / 5.2 / 33.2 ms (2.36 / 9.8 / 47 before).  DESIGN.md 4.2 has the evidence
// and the list of what was measured and dropped.
#pragma once

#include <type_traits>

#include "rowpair1.cuh"
#include "viterbi_pipe.cuh"

namespace coati_gpu {

// Steps per group (one window update, one basic block) and how far ahead of its use a column of the row above is
// fetched.  Eight-step groups halve the window overhead and give ptxas more to interleave: modelled 141 instead of
// 164 cycles per step at R = 4 (97 / 129 at R = 2, 244 / 277 at R = 8), measured 164 / 177 (109 / 139, 277 / 299);
// the widest tile gains nothing in the step (332) and loses in the lag (24.6 -> 30.9 k cycles), so it keeps four.
__host__ __device__ constexpr uint32_t wave_group(int R) { return R >= 10 ? 4u : 8u; }
__host__ __device__ constexpr uint32_t wave_ahead(int R) { return R >= 10 ? 12u : 16u; }  // three groups of four / two groups of eight
#ifndef COATI_WAVE_SLEEP
#define COATI_WAVE_SLEEP 20  // ns between polls of a column the producer has not written yet (0 / 20 / 40 measured: no difference)
#endif

// The consumer's side of the hand-off: a relaxed load at gpu scope (L2), re-issued until the sentinel is gone.
__device__ __forceinline__ float2 ld_relaxed_f2(const float2* p) {
    float2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// The producer's side of the hand-off: one 64-bit relaxed store at gpu scope (a plain weak store racing with the
// relaxed polls would be a data race under the PTX memory model; weak / volatile / write-through stores and an
// L2 exchange were measured: the first three change nothing, the exchange doubles the step).  Predicated inside
// the asm: an `if` around it becomes a divergent branch and cuts the step's basic block in two.
__device__ __forceinline__ void st_relaxed_f2_if(bool on, float2* p, float x, float y) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q st.relaxed.gpu.global.v2.f32 [%0], {%1, %2}; }" ::"l"(p),
                 "f"(x), "f"(y), "r"((uint32_t)on));
}
__device__ __forceinline__ void st_f32_if(bool on, float* p, float x) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.f32 [%0], %1; }" ::"l"(p), "f"(x), "r"((uint32_t)on));
}

#ifdef COATI_WAVE_TRACE  // diagnostics build (tools/wave_trace.py): per band, the time it reaches a few marks
__device__ unsigned long long g_wave_trace[8 * 8192];
__device__ __forceinline__ void wave_mark(uint32_t band, int slot, int lane) {
    if(lane == 0 && band < 8192) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_wave_trace[band * 8 + slot] = t;
    }
}
#define COATI_WAVE_MARK(slot) wave_mark(band, slot, lane)
#else
#define COATI_WAVE_MARK(slot)
#endif

template <int R, int NC>
__global__ void __launch_bounds__(PIPE_WARPS * 32)
viterbi_wave1_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t last,
                     unsigned int* __restrict__ counter, const uint8_t* __restrict__ a_all,
                     const uint8_t* __restrict__ b_all, const float* __restrict__ table, GapConsts g,
                     float4* __restrict__ bnd_all, uint32_t bnd_stride, uint8_t* __restrict__ dirs,
                     PairResult* __restrict__ results, const unsigned int* __restrict__ nc_flag) {
    static_assert(R % 2 == 0, "rows are processed in pairs");
    constexpr uint32_t WAVE_U = wave_group(R), WAVE_D = wave_ahead(R);
    static_assert(WAVE_D == 3 * WAVE_U || WAVE_D == 2 * WAVE_U, "a fetch enters the window two or three groups after it was issued");
    constexpr bool THREE = WAVE_D == 3 * WAVE_U;  // a fetched register is read one group (>= the L2 round trip) after the load
    constexpr int R4 = (R + 3) / 4;
    constexpr int H = 32 * R;
    constexpr uint32_t WPL = (5 * R + 3) & ~3u;
    extern __shared__ float4 s_dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* s_tab = s_dyn + (size_t)warp * R4 * NC * 32;
    if(nc_flag && ((*nc_flag != 0) != (NC == 16))) return;
    const uint32_t FULL = 0xffffffffu;
    const int rot = (lane + 31) & 31;
    const f2 ng2 = mk2(g.ng, g.ng), go2 = mk2(g.go, g.go), gs2 = mk2(g.gs, g.gs), ge2 = mk2(g.ge, g.ge);
    const char* tab_lane = reinterpret_cast<const char*>(s_tab) + lane * 16;

    const PairDesc pd = pairs[first];
    if(results[pd.orig].status != 0 || pd.la == 0 || pd.lb == 0) return;
    const uint32_t la = pd.la, lb = pd.lb;
    const float* tab = table + (size_t)(pd.cfg >> CFG_MODEL_SHIFT) * (TABLE_ROWS * TABLE_LD);
    const uint8_t* a = a_all + pd.a_off;
    const uint8_t* b = b_all + pd.b_off;
    uint4* dir = reinterpret_cast<uint4*>(dirs + pd.dir_off);
    const uint32_t nblocks = pipe_nblocks(lb, R);
    const uint32_t nbands = (la + H - 1) / H;
    const uint32_t nsteps = lb + 31;
    float2* bnd = reinterpret_cast<float2*>(bnd_all);  // wave_bnd[band] = bnd + band * 2 * bnd_stride
    (void)last;

    for(;;) {
        uint32_t band = 0;
        if(lane == 0) band = atomicAdd(counter, 1u);
        band = __shfl_sync(FULL, band, 0);
        if(band >= nbands) break;
        COATI_WAVE_MARK(0);  // ticket
        const float2* bin = bnd + (size_t)band * 2 * bnd_stride;
        float2* bout = bnd + (size_t)(band + 1) * 2 * bnd_stride;
        // row above band 0 = top margin row r = 0 (align_pair.cc:88-90)
        if(band == 0) {
            for(uint32_t c = 1 + lane; c <= lb; c += 32) {
                const CellOut o = cell_out<1>(LOWEST, LOWEST, margin_ins<1>(c, g), g);
                st_relaxed_f2_if(true, bnd + c, o.X, o.Y);
            }
        }
        __syncwarp();

        const uint32_t r0 = band * H + lane * R + 1;  // first row of this lane
        // ---- private substitution rows: s_tab[h][nuc][lane] = rows 4h..4h+3 -----------------------
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
        // ---- state at column 0 (left margin, align_pair.cc:84-87) ---------------------------------
        // Every lane runs every step, inside the lattice or not (see `step`): a lane takes this state over on the
        // step it enters the lattice (column 1), so X0 / diag0 live through the first block only.
        float X0[R], diag0;
#pragma unroll
        for(int q = 0; q < R; ++q) X0[q] = margin_del<1>(r0 + q, g) + g.gs;  // X(r, 0): only D is finite; Z(r, 0) = LOWEST
        diag0 = r0 == 1 ? (0.0f + g.ng) + g.ng : margin_del<1>(r0 - 1, g) + g.gs;
        float Xp[R], Zp[R], diagX = diag0;
        uint32_t acc[R][5];
#pragma unroll
        for(int q = 0; q < R; ++q) {
            Xp[q] = X0[q], Zp[q] = LOWEST;
#pragma unroll
            for(int j = 0; j < 5; ++j) acc[q][j] = 0;
        }
        // the lane and row that end on the terminal cell (La, Lb): Viterbi score = X(La, Lb), the max3 of the
        // adjusted terminal scores (align_pair.cc:130-138, 265)
        const uint32_t rr = (la - 1) % H;
        const bool score_lane = band == nbands - 1 && (uint32_t)lane == rr / R;
        __syncwarp();  // s_tab

        // ---- rolling window of the row above: lane j holds column j (mod 32) -----------------------
        float2 win = make_float2(0.f, 0.f);
        // Fetches land in registers of their own (the scoreboard tracks a register for the whole warp: a load into
        // `win` would stall every shuffle that reads it) and are taken over two groups later, when they are due.
        float2 pf0 = make_float2(0.f, 0.f), pf1 = pf0, pf2 = pf0;
        // the column congruent to this lane in [base, base + 32), clamped to lb (columns past lb are never used)
        auto wcol = [&](uint32_t base) { return min(base + (((uint32_t)lane - base) & 31u), lb); };
        auto wmine = [&](uint32_t base) { return (((uint32_t)lane - base) & 31u) < WAVE_U; };
        // Group tg (four steps, which shuffle the columns tg + 3 .. tg + 6 out of the window, each one step before
        // lane 31 hands it to lane 0): those columns, fetched three groups ago into pf0 and found complete one group
        // ago (or re-polled now: the producer was late), enter the window; columns tg + 15 .. tg + 18 are fetched,
        // not waited for; the fetch that is due next is examined -- the vote is issued here and branched on a group
        // later, so neither the loads nor the vote's latency sit on the chain.
        bool late = true;
        auto wgroup = [&](uint32_t tg) {
            const bool mine = wmine(tg + 3);
            if(__builtin_expect(late, 0)) {
                while(__any_sync(FULL, mine && pf0.x != pf0.x)) {
                    if(COATI_WAVE_SLEEP) __nanosleep(COATI_WAVE_SLEEP);
                    if(mine && pf0.x != pf0.x) pf0 = ld_relaxed_f2(bin + wcol(tg + 3));
                }
            }
            win.x = mine ? pf0.x : win.x, win.y = mine ? pf0.y : win.y;
            if(THREE) {
                pf0 = pf1, pf1 = pf2;
                if(wmine(tg + 3 + WAVE_D)) pf2 = ld_relaxed_f2(bin + wcol(tg + 3 + WAVE_D));
            } else {
                pf0 = pf1;
                if(wmine(tg + 3 + WAVE_D)) pf1 = ld_relaxed_f2(bin + wcol(tg + 3 + WAVE_D));
            }
            late = __any_sync(FULL, wmine(tg + 3 + WAVE_U) && pf0.x != pf0.x);
        };
        // lane 31's outgoing registers carry lane 0's inputs, the row above the band: column 1 now, and column
        // t + 2 (wxn, wyn: shuffled out of the window during step t - 1) after step t
        float outX = 0.f, outY = 0.f, wxn, wyn;
        {
            float2 w1 = ld_relaxed_f2(bin + 1 + (lane & 1));  // even lanes: column 1, odd lanes: column 2
            while(__any_sync(FULL, w1.x != w1.x)) {
                if(COATI_WAVE_SLEEP) __nanosleep(COATI_WAVE_SLEEP);
                w1 = ld_relaxed_f2(bin + 1 + (lane & 1));
            }
            COATI_WAVE_MARK(1);  // columns 1 and 2 of the row above seen
            wxn = __shfl_sync(FULL, w1.x, 1), wyn = __shfl_sync(FULL, w1.y, 1);
            const float x1 = __shfl_sync(FULL, w1.x, 0), y1 = __shfl_sync(FULL, w1.y, 0);
            if(lane == 31) outX = x1, outY = y1;
        }
        if(wmine(3)) pf0 = ld_relaxed_f2(bin + wcol(3));  // columns 3 .. 14: groups 0, 1 and 2
        if(wmine(3 + WAVE_U)) pf1 = ld_relaxed_f2(bin + wcol(3 + WAVE_U));
        if(THREE && wmine(3 + 2 * WAVE_U)) pf2 = ld_relaxed_f2(bin + wcol(3 + 2 * WAVE_U));
        // ---- symbols a group ahead, substitution scores one step ahead -------------------------------
        // lane l is at column t - l + 1 on step t: symbol b[t - l]; indices are clamped outside the lattice
        auto sym_idx = [&](uint32_t t) {
            return (uint32_t)min(max((int)t - lane, 0), (int)lb - 1);
        };
        float sv[R4 * 4];
        auto lds_scores = [&](float (&dst)[R4 * 4], uint32_t off) {
#pragma unroll
            for(int h = 0; h < R4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(tab_lane + off + h * (NC * 512));
                dst[4 * h] = v.x, dst[4 * h + 1] = v.y, dst[4 * h + 2] = v.z, dst[4 * h + 3] = v.w;
            }
        };
        lds_scores(sv, ld_symbol_now(b + sym_idx(0)) * 512u);
        // symc[i]: symbol of step tg + i + 1 (the scores step tg + i loads); symn: the same for the next group,
        // loaded at the start of this one
        uint32_t symc[WAVE_U], symn[WAVE_U];
#pragma unroll
        for(uint32_t i = 0; i < WAVE_U; ++i) symc[i] = 0, symn[i] = ld_symbol_now(b + sym_idx(1 + i));
        const uint8_t* b_lane = b - lane;  // b_lane[t] is this lane's symbol on step t

        // ---- one step ----------------------------------------------------------------------------------
        // EVERY lane computes on EVERY step, so a step is one basic block without a per-lane activity test.  A lane
        // that has not entered the lattice yet, or has left it, works on garbage that no lane inside the lattice
        // ever reads: lane l's first real input is lane l - 1's output of the step before (already inside), its
        // decision bits of those steps land on bit positions the traceback never visits, and only what leaves the
        // warp is guarded (the boundary store, the score).  The steps at the two ends of the band are on the
        // critical path of the whole wavefront -- band b + 1 starts 32 + D steps after band b and ends 32 steps
        // after it -- so they must not be slower than the others (a branchy version of them cost 22 000 cycles
        // of lag per band, tools/wave_lag.py).
        //   PHASE 0: first block (t < 32): lanes enter the lattice -> take over the column-0 state at column 1
        //   PHASE 1: all lanes inside, for the symbols fetched ahead too (no clamps, no guards)
        //   PHASE 2: last blocks: lanes leave the lattice -> guarded boundary store, score
        const uint32_t qsel = rr % R;
        float* const score_out = &results[pd.orig].score;
        const uint8_t* psym = b_lane + WAVE_U + 1;           // psym[t]: this lane's symbol on step t + 5 (PHASE 1: no clamp)
        float2* pst = bout + 1 - (ptrdiff_t)lane;   // pst[t]: where lane 31 puts the bottom row on step t
        auto step = [&](auto phase, uint32_t t, uint32_t i) {  // t = tg + i, i the unrolled index within the group
            constexpr int PHASE = decltype(phase)::value;
            const uint32_t u = t - (uint32_t)lane;  // c - 1
            const float recvX = __shfl_sync(FULL, outX, rot);
            const float recvY = __shfl_sync(FULL, outY, rot);
            const float wx = wxn, wy = wyn;         // column t + 2, for lane 0's next step
            wxn = __shfl_sync(FULL, win.x, t + 3);  // source lane (t + 3) mod 32: column t + 3
            wyn = __shfl_sync(FULL, win.y, t + 3);
            float svn[R4 * 4];
            lds_scores(svn, symc[i] * 512u);
            if(PHASE == 0) {
                const bool enter = u == 0;
#pragma unroll
                for(int q = 0; q < R; ++q) Xp[q] = enter ? X0[q] : Xp[q], Zp[q] = enter ? LOWEST : Zp[q];
                diagX = enter ? diag0 : diagX;
            }
            float D = recvY;
            float Mv[R];  // every match score first, from the previous column's X
            Mv[0] = diagX + sv[0];
#pragma unroll
            for(int q = 1; q < R; ++q) Mv[q] = Xp[q - 1] + sv[q];
#pragma unroll
            for(int q = 0; q < R; q += 2) COATI_ROWPAIR_SGN(q)
            diagX = recvX;
            st_relaxed_f2_if(PHASE == 1 ? lane == 31 : (lane == 31 && u < lb), pst + i, Xp[R - 1], D);
            if(PHASE == 2) {
                float score = Xp[0];
#pragma unroll
                for(int q = 1; q < R; ++q) score = (uint32_t)q == qsel ? Xp[q] : score;
                st_f32_if(score_lane && u == lb - 1, score_out, score);
            }
            outX = lane == 31 ? wx : Xp[R - 1];
            outY = lane == 31 ? wy : D;
#pragma unroll
            for(int x = 0; x < R4 * 4; ++x) sv[x] = svn[x];
        };
        auto flush = [&](uint32_t t0) {  // every lane pushed 32 bits since the last flush: step t is bit 31 - t % 32
            uint4* dst = dir + ((size_t)(band * nblocks + (t0 >> 5)) * 32 + lane) * (WPL / 4);
            uint32_t w[WPL];
#pragma unroll
            for(int x = 0; x < (int)WPL; ++x) w[x] = x < 5 * R ? (x % 5 < 4 ? ~acc[x / 5][x % 5] : acc[x / 5][x % 5]) : 0u;
#pragma unroll
            for(int x = 0; x < (int)WPL / 4; ++x)
                dst[x] = make_uint4(w[4 * x], w[4 * x + 1], w[4 * x + 2], w[4 * x + 3]);
        };
        auto block = [&](auto phase, uint32_t t0) {
#pragma unroll 1
            for(uint32_t tg = t0; tg < t0 + 32u; tg += WAVE_U) {
                constexpr int PHASE = decltype(phase)::value;
                wgroup(tg);
#pragma unroll
                for(uint32_t i = 0; i < WAVE_U; ++i) {
                    symc[i] = symn[i];
                    symn[i] = ld_symbol_now(PHASE == 1 ? psym + i : b + sym_idx(tg + WAVE_U + 1 + i));
                }
#pragma unroll
                for(uint32_t i = 0; i < WAVE_U; ++i) step(phase, tg + i, i);
                psym += WAVE_U, pst += WAVE_U;
            }
        };

        // blocks of 32 steps = one word of every decision plane; the flush sits between blocks.  The last block
        // runs its 32 steps too (past step lb + 30 every lane is outside the lattice).  Needs lb >= 34.
        for(uint32_t t0 = 0; t0 < nsteps; t0 += 32) {
            if(t0 == 0) block(std::integral_constant<int, 0>{}, t0);
            else if(t0 + 36u + WAVE_U <= lb) block(std::integral_constant<int, 1>{}, t0);  // inside, and so are the symbols fetched ahead
            else block(std::integral_constant<int, 2>{}, t0);
            flush(t0);
#ifdef COATI_WAVE_TRACE
            if(t0 == 0) COATI_WAVE_MARK(2);                          // 32 steps done
            if(t0 == 992) COATI_WAVE_MARK(3);                        // 1024 steps done
            if(t0 == 4064) COATI_WAVE_MARK(4);                       // 4096 steps done
            if(t0 + 32 >= nsteps) COATI_WAVE_MARK(5);                // band done
#endif
        }
        __syncwarp();  // s_tab is rewritten for the next ticket
    }
}

}  // namespace coati_gpu
