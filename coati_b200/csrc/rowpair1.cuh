// Packed-FADD2 row pair of the K = 1 cell update, shared by the inter-pair fill (viterbi_pipe1.cuh) and the
// intra-pair wavefront (viterbi_wave1.cuh).  See viterbi_pipe.cuh for the X / Y / Z refactoring of
// forward_impl (src/lib/align_pair.cc:94-129) and the exactness argument.
#pragma once

#include "common.cuh"

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
    float lo;
    asm("mov.b64 {%0, _}, %1;" : "=f"(lo) : "l"(a.v));
    return lo;
}
__device__ __forceinline__ float hi2(f2 a) {
    float hi;
    asm("mov.b64 {_, %0}, %1;" : "=f"(hi) : "l"(a.v));
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

// A symbol of the descendant, loaded NOW into a register that is then kept: with a plain `b[i]` the compiler
// re-loads the byte at the point of use instead (the pointer is const __restrict__).
__device__ __forceinline__ uint32_t ld_symbol_now(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
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

}  // namespace coati_gpu
