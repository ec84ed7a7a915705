/* TEST INFRASTRUCTURE ONLY.
 *
 * Host restatement of the three glibc 2.39 (x86-64, FMA ifunc variants) libm functions the
 * reference's Forward/sampling path calls -- the same algorithms coati_b200/csrc/devmath.cuh runs on
 * the device -- plus a checker that compares them with the RUNNING libm over a range of float bit
 * patterns.  Exhaustive runs (every float in [-104.5, 89.5] for expf, every positive float for logf,
 * every float in (-1, 1e30] for log1pf) report zero mismatches on glibc 2.39.
 *   expf  : sysdeps/ieee754/flt-32/e_expf.c   (exp2f table scheme, double arithmetic)
 *   logf  : sysdeps/ieee754/flt-32/e_logf.c
 *   log1pf: sysdeps/ieee754/flt-32/s_log1pf.c (fdlibm float algorithm)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t asuint64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double asdouble(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

static uint64_t T[32];
static int T_ready = 0;
static void init_tab(void) {
    if(T_ready) return;
    for(int i = 0; i < 32; i++) T[i] = asuint64(exp2(i / 32.0)) - ((uint64_t)i << 47);
    T_ready = 1;
}

float orc_port_expf(float x) {
    const double C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6,
                 InvLn2N = 0x1.71547652b82fep+5, SHIFT = 0x1.8p+52;
    init_tab();
    double xd = (double)x;
    uint32_t abstop = (asuint(x) >> 20) & 0x7ff;
    if(abstop >= (asuint(88.0f) >> 20)) {
        if(asuint(x) == asuint(-INFINITY)) return 0.0f;
        if(abstop >= (asuint(INFINITY) >> 20)) return x + x;
        if(x > 0x1.62e42ep6f) return INFINITY;
        if(x < -0x1.9fe368p6f) return 0.0f;
    }
    double z = InvLn2N * xd;
    double kd = z + SHIFT;
    uint64_t ki = asuint64(kd);
    kd -= SHIFT;
    double r = fma(InvLn2N, xd, -kd);
    uint64_t t = T[ki % 32];
    t += ki << (52 - 5);
    double s = asdouble(t);
    z = fma(C0, r, C1);
    double r2 = r * r;
    double y = fma(C2, r, 1.0);
    y = fma(z, r2, y);
    y = y * s;
    return (float)y;
}

static const struct { double invc, logc; } LT[16] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

float orc_port_logf(float x) {
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2,
                 Ln2 = 0x1.62e42fefa39efp-1;
    uint32_t ix = asuint(x);
    if(ix == 0x3f800000) return 0;
    if(ix - 0x00800000 >= 0x7f800000 - 0x00800000) {
        if(ix * 2 == 0) return -INFINITY;
        if(ix == 0x7f800000) return x;
        if((ix & 0x80000000) || ix * 2 >= 0xff000000) return NAN;
        ix = asuint(x * 0x1p23f);
        ix -= 23 << 23;
    }
    uint32_t tmp = ix - 0x3f330000;
    int i = (tmp >> (23 - 4)) % 16;
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000);
    double invc = LT[i].invc, logc = LT[i].logc;
    double z = (double)asfloat(iz);
    double r = fma(z, invc, -1.0);
    double y0 = fma((double)k, Ln2, logc);
    double r2 = r * r;
    double y = fma(A1, r, A2);
    y = fma(A0, r2, y);
    y = fma(y, r2, (y0 + r));
    return (float)y;
}

float orc_port_log1pf(float x) {
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, Lp1 = 6.6666668653e-01f,
                Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f, Lp4 = 2.2222198546e-01f,
                Lp5 = 1.8183572590e-01f, Lp6 = 1.5313838422e-01f, Lp7 = 1.4798198640e-01f;
    float hfsq, f = 0, c = 0, s, z, R, u;
    int32_t k, hx, hu = 0, ax;
    hx = (int32_t)asuint(x);
    ax = hx & 0x7fffffff;
    k = 1;
    if(hx < 0x3ed413d7) {
        if(ax >= 0x3f800000) return x == -1.0f ? -INFINITY : NAN;
        if(ax < 0x31000000) {
            if(ax < 0x24800000) return x;
            return x - x * x * 0.5f;
        }
        if(hx > 0 || hx <= ((int32_t)0xbe95f61f)) {
            k = 0;
            f = x;
            hu = 1;
        }
    }
    if(hx >= 0x7f800000) return x + x;
    if(k != 0) {
        if(hx < 0x5a000000) {
            u = 1.0f + x;
            hu = (int32_t)asuint(u);
            k = (hu >> 23) - 127;
            c = (k > 0) ? 1.0f - (u - x) : x - (u - 1.0f);
            c /= u;
        } else {
            u = x;
            hu = (int32_t)asuint(u);
            k = (hu >> 23) - 127;
            c = 0;
        }
        hu &= 0x007fffff;
        if(hu < 0x3504f7) {
            u = asfloat(hu | 0x3f800000);
        } else {
            k += 1;
            u = asfloat(hu | 0x3f000000);
            hu = (0x00800000 - hu) >> 2;
        }
        f = u - 1.0f;
    }
    hfsq = 0.5f * f * f;
    if(hu == 0) {
        if(f == 0.0f) {
            if(k == 0) return 0.0f;
            c += k * ln2_lo;
            return k * ln2_hi + c;
        }
        R = hfsq * (1.0f - 0.66666666666666666f * f);
        if(k == 0) return f - R;
        return k * ln2_hi - ((R - (k * ln2_lo + c)) - f);
    }
    s = f / (2.0f + f);
    z = s * s;
    R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
    if(k == 0) return f - (hfsq - s * (hfsq + R));
    return k * ln2_hi - ((hfsq - (s * (hfsq + R) + (k * ln2_lo + c))) - f);
}

/* Compare port vs the running libm on bit patterns first, first+stride, ... <= last.
 * op 0: expf, 1: logf, 2: log1pf.  Returns the number of mismatching results (NaN == NaN). */
uint64_t orc_libm_check(int op, uint32_t first, uint32_t last, uint32_t stride, uint64_t* checked) {
    uint64_t bad = 0, n = 0;
    for(uint64_t u = first; u <= last; u += stride) {
        float x = asfloat((uint32_t)u), p, l;
        if(op == 0) p = orc_port_expf(x), l = expf(x);
        else if(op == 1) p = orc_port_logf(x), l = logf(x);
        else p = orc_port_log1pf(x), l = log1pf(x);
        ++n;
        if(asuint(p) != asuint(l) && !(p != p && l != l)) ++bad;
    }
    if(checked) *checked = n;
    return bad;
}
