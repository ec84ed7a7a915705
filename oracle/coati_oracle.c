/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's marginal Gotoh hot path (COATi,
 * CartwrightLab/coati).  It is the checker for tests/, __graft_entry__.smoke() and the
 * `cpu_baseline` leg of bench.py; the product (coati_b200/) never includes, links or calls
 * it.  Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 *
 * Parity pin: tests/test_oracle_vs_ref.py checks this file bit-for-bit against the reference
 * itself (oracle/_ref/libcoati_ref.so = the reference's align_pair.cc + contrib/random
 * compiled unmodified) on randomised pairs, and tests/test_oracle_golden.py checks it against
 * the reference's own known-answer tests (align_marginal.cc:149-240, 477-509, 598-723;
 * utils.cc:532-586, 971-1094, 1168-1227) through the committed fixtures in tests/golden/.
 *
 * All arithmetic is float32, round-to-nearest, no FMA contraction (-ffp-contract=off), with
 * glibc libm for logf/log1pf/expf exactly as the reference uses them.
 */
#include "coati_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOWEST (-FLT_MAX) /* semiring.hpp:82-84,114-116: zero() = numeric_limits<float>::lowest() */

static inline float fmax2(float x, float y) { return (x < y) ? y : x; } /* std::max */

/* ---- utils.hpp:134-146: float log1p_exp --------------------------------------------- */
float orc_log1p_exp(float x) {
    if(x <= -16.0f) return expf(x);
    if(x <= 8.0f) return log1pf(expf(x));
    if(x <= 14.5f) return x + expf(-x);
    return x;
}

/* the same over an array (tests compare device code with the host libm on whole ranges of floats) */
void orc_log1p_exp_array(const float* in, float* out, size_t n) {
    for(size_t x = 0; x < n; ++x) out[x] = orc_log1p_exp(in[x]);
}

/* ---- utils.hpp:152-156: log_sum_exp ------------------------------------------------- */
float orc_log_sum_exp(float a, float b) {
    float x = fmax2(a, b);
    float y = -fabsf(a - b);
    return x + orc_log1p_exp(y);
}

static inline float plus2(int sr, float x, float y) {
    return sr == ORC_TROPICAL ? fmax2(x, y) : orc_log_sum_exp(x, y); /* semiring.hpp:65-71,97-102 */
}
static inline float plus3(int sr, float x, float y, float z) { return plus2(sr, plus2(sr, x, y), z); }

typedef struct {
    float ng, gs, go, ge;
} gapc_t;

/* align_pair.cc:66-69 (and :253-256): log(1-g), log(1-e), log(g), log(e) */
static gapc_t gap_consts(float g, float e) {
    gapc_t c;
    c.ng = log1pf(-g);
    c.gs = log1pf(-e);
    c.go = logf(g);
    c.ge = logf(e);
    return c;
}

/* semiring.hpp:76-78,109-111: power(x, n) = x * float(n) */
static inline float powerf(float x, size_t n) { return x * (float)n; }

/* ---- align_pair.cc:62-139: forward_impl<S,W> ---------------------------------------- */
int orc_fill(int sr, const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const float* table,
             float g, float e, size_t k, float* mch, float* del, float* ins, float* const* trans) {
    if(k == 0 || !mch || !del || !ins) return ORC_E_ARG;
    const gapc_t c = gap_consts(g, e);
    const size_t start = k - 1;
    const size_t len_a = la + k, len_b = lb + k;
    const size_t n = len_a * len_b;
#define AT(m, i, j) (m)[(i) * len_b + (j)]
    for(size_t x = 0; x < n; ++x) mch[x] = del[x] = ins[x] = LOWEST; /* :79 resize(..., lowest) */
    if(trans)
        for(int t = 0; t < 8; ++t)
            for(size_t x = 0; x < n; ++x) trans[t][x] = LOWEST;

    AT(mch, start, start) = 0.0f; /* :82 S::one() */
    for(size_t i = start + k; i < len_a; i += k) /* :84-87 */
        AT(del, i, start) = (c.ng + c.go) + powerf(c.ge, i - 1);
    for(size_t j = start + k; j < len_b; j += k) /* :88-90 */
        AT(ins, start, j) = c.go + powerf(c.ge, j - 1);
    if(trans) { /* :91 init_margins(): del_del = del; ins_ins = ins (align_pair.hpp:108-111) */
        memcpy(trans[ORC_DEL_DEL], del, n * sizeof(float));
        memcpy(trans[ORC_INS_INS], ins, n * sizeof(float));
    }

    const float gk1 = powerf(c.ge, k - 1), gk = powerf(c.ge, k);
    for(size_t i = k; i < len_a; ++i) { /* :94-129 */
        for(size_t j = k; j < len_b; ++j) {
            float s = table[(size_t)a[i - k] * 15 + b[j - k]];
            float m2m = ((AT(mch, i - 1, j - 1) + c.ng) + c.ng) + s;
            float d2m = (AT(del, i - 1, j - 1) + c.gs) + s;
            float i2m = ((AT(ins, i - 1, j - 1) + c.gs) + c.ng) + s;
            float m2d = ((AT(mch, i - k, j) + c.ng) + c.go) + gk1;
            float i2d = ((AT(ins, i - k, j) + c.gs) + c.go) + gk1;
            float d2d = AT(del, i - k, j) + gk;
            float m2i = (AT(mch, i, j - k) + c.go) + gk1;
            float i2i = AT(ins, i, j - k) + gk;
            AT(mch, i, j) = plus3(sr, m2m, d2m, i2m);
            AT(del, i, j) = plus3(sr, m2d, d2d, i2d);
            AT(ins, i, j) = plus2(sr, m2i, i2i);
            if(trans) { /* align_pair.hpp:94-103 save_values */
                AT(trans[ORC_MCH_MCH], i, j) = m2m;
                AT(trans[ORC_MCH_DEL], i, j) = m2d;
                AT(trans[ORC_MCH_INS], i, j) = m2i;
                AT(trans[ORC_DEL_MCH], i, j) = d2m;
                AT(trans[ORC_DEL_DEL], i, j) = d2d;
                AT(trans[ORC_INS_MCH], i, j) = i2m;
                AT(trans[ORC_INS_DEL], i, j) = i2d;
                AT(trans[ORC_INS_INS], i, j) = i2i;
            }
        }
    }
    /* :130-138 terminal state */
    AT(mch, len_a - 1, len_b - 1) = (AT(mch, len_a - 1, len_b - 1) + c.ng) + c.ng;
    AT(ins, len_a - 1, len_b - 1) = (AT(ins, len_a - 1, len_b - 1) + c.gs) + c.ng;
    AT(del, len_a - 1, len_b - 1) = AT(del, len_a - 1, len_b - 1) + c.gs;
    return ORC_OK;
}

/* align_pair.cc:210-221 max_mdi; 0 = MATCH, 1 = DELETION, 2 = INSERTION */
static int max_mdi(float m, float d, float i) {
    int st = 0;
    float val = m;
    if(d > val) {
        val = d;
        st = 1;
    }
    if(i > val) return 2;
    return st;
}
/* align_pair.cc:230-232 max_mi */
static int max_mi(float m, float i) { return m > i ? 0 : 2; }

static void reverse(char* s, size_t n) {
    for(size_t x = 0; x < n / 2; ++x) {
        char t = s[x];
        s[x] = s[n - 1 - x];
        s[n - 1 - x] = t;
    }
}

/* ---- align_pair.cc:249-303: traceback<tropical> ------------------------------------- */
int orc_traceback(const float* mch, const float* del, const float* ins, size_t la, size_t lb,
                  const char* anc, const char* des, float g, float e, size_t k, char* out_a,
                  char* out_b, size_t* out_len, float* score) {
    const gapc_t c = gap_consts(g, e);
    const size_t len_b = lb + k;
    size_t i = la + k - 1, j = lb + k - 1, n = 0;
    const size_t cap = la + lb;
    if(score) *score = fmax2(fmax2(AT(mch, i, j), AT(del, i, j)), AT(ins, i, j)); /* :265 */
    int m = max_mdi(AT(mch, i, j), AT(del, i, j), AT(ins, i, j));                   /* :266 */
    while(j > (k - 1) || i > (k - 1)) {                                             /* :268 */
        if(m == 0) {
            if(n + 1 > cap || i == 0 || j == 0) return ORC_E_ARG;
            out_a[n] = anc[i - k];
            out_b[n] = des[j - k];
            ++n;
            i--;
            j--;
            m = max_mdi((AT(mch, i, j) + c.ng) + c.ng, AT(del, i, j) + c.gs,
                        (AT(ins, i, j) + c.gs) + c.ng);
        } else if(m == 1) {
            if(n + k > cap || i < k) return ORC_E_ARG;
            for(size_t r = i; r > i - k; r--) {
                out_a[n] = anc[r - k];
                out_b[n] = '-';
                ++n;
            }
            i -= k;
            m = max_mdi((AT(mch, i, j) + c.ng) + c.go, AT(del, i, j) + c.ge,
                        (AT(ins, i, j) + c.gs) + c.go);
        } else {
            if(n + k > cap || j < k) return ORC_E_ARG;
            for(size_t q = j; q > j - k; q--) {
                out_a[n] = '-';
                out_b[n] = des[q - k];
                ++n;
            }
            j -= k;
            m = max_mi(AT(mch, i, j) + c.go, AT(ins, i, j) + c.ge);
        }
    }
    reverse(out_a, n); /* :301-302 */
    reverse(out_b, n);
    out_a[n] = 0;
    out_b[n] = 0;
    if(out_len) *out_len = n;
    return ORC_OK;
}

/* The decisions traceback (align_pair.cc:275-296) would take when it LANDS on cell (i,j),
 * for each state it could have arrived from.  This is the byte stream the CUDA fill emits. */
int orc_directions(const float* mch, const float* del, const float* ins, size_t la, size_t lb,
                   float g, float e, size_t k, uint8_t* dirs) {
    const gapc_t c = gap_consts(g, e);
    const size_t len_a = la + k, len_b = lb + k;
    for(size_t i = 0; i < len_a; ++i)
        for(size_t j = 0; j < len_b; ++j) {
            float M = AT(mch, i, j), D = AT(del, i, j), I = AT(ins, i, j);
            if(i == len_a - 1 && j == len_b - 1) { dirs[i * len_b + j] = 0; continue; } /* adjusted */
            int x = max_mdi((M + c.ng) + c.ng, D + c.gs, (I + c.gs) + c.ng);
            int y = max_mdi((M + c.ng) + c.go, D + c.ge, (I + c.gs) + c.go);
            int z = max_mi(M + c.go, I + c.ge);
            dirs[i * len_b + j] = (uint8_t)(x | (y << 2) | ((z ? 1 : 0) << 4));
        }
    return ORC_OK;
}
#undef AT

/* viterbi_mem + traceback_viterbi as marg_alignment calls them (align_marginal.cc:69-80) */
int orc_viterbi(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const char* anc,
                const char* des, const float* table, float g, float e, size_t k, char* out_a,
                char* out_b, size_t* out_len, float* score) {
    size_t n = (la + k) * (lb + k);
    float* buf = (float*)malloc(3 * n * sizeof(float));
    if(!buf) return ORC_E_NOMEM;
    int rc = orc_fill(ORC_TROPICAL, a, la, b, lb, table, g, e, k, buf, buf + n, buf + 2 * n, NULL);
    if(rc == ORC_OK)
        rc = orc_traceback(buf, buf + n, buf + 2 * n, la, lb, anc, des, g, e, k, out_a, out_b,
                           out_len, score);
    free(buf);
    return rc;
}

/* ---- contrib/random/random.hpp ------------------------------------------------------- */
typedef unsigned __int128 u128;
#define MCG_MULT 0xda942042e4dd58b5ULL /* :90 */

void orc_rng_set_state(uint64_t state[2], uint64_t lo, uint64_t hi) { /* :131-134 state | 1 */
    state[0] = lo | 1u;
    state[1] = hi;
}
uint64_t orc_rng_bits(uint64_t state[2]) { /* :107,122-125 */
    u128 s = ((u128)state[1] << 64) | state[0];
    s *= MCG_MULT;
    state[0] = (uint64_t)s;
    state[1] = (uint64_t)(s >> 64);
    return state[1];
}
float orc_rng_f24(uint64_t state[2]) { /* :213-216 */
    int64_t n = (int64_t)(orc_rng_bits(state) >> 40);
    return n / 16777216.0f;
}

/* :334-358 hash_impl_t: Weyl-sequence multilinear hash */
static void mlhash(uint64_t init, const uint32_t* in, size_t nin, uint32_t* out, size_t nout) {
    const uint64_t INC = 0x9e3779b97f4a7c15ULL;
    uint64_t w = init;
    for(size_t o = 0; o < nout; ++o) {
        w += INC;
        uint64_t sum = w;
        for(size_t x = 0; x < nin; ++x) {
            w += INC;
            sum += w * in[x];
        }
        w += INC;
        sum += w * 1;
        out[o] = (uint32_t)(sum >> 32);
    }
}

/* :366-398 SeedSeq<8>::Seed/Generate + :408-413 Random::Seed(SeedSeq) + :101-105 memcpy */
void orc_rng_seed_u32(const uint32_t* seeds, size_t n, uint64_t state[2]) {
    uint32_t inner[8], outw[4];
    mlhash(0x3423da0b87484307ULL, seeds, n, inner, 8);
    mlhash(0xdf8b06c40fa44478ULL, inner, 8, outw, 4);
    uint64_t lo = (uint64_t)outw[0] | ((uint64_t)outw[1] << 32);
    uint64_t hi = (uint64_t)outw[2] | ((uint64_t)outw[3] << 32);
    orc_rng_set_state(state, lo, hi);
}

uint32_t orc_fnv1(const char* s, size_t n) { /* :465-472 (char is signed on x86-64) */
    uint32_t h = 2166136261U;
    for(size_t x = 0; x < n; ++x) h = (h * 16777619U) ^ (uint32_t)(int)(signed char)s[x];
    return h;
}

/* std::from_chars<int32_t>(…, 10) consuming the whole string (:530-535) */
static int parse_i32(const char* s, size_t n, int32_t* out) {
    size_t x = 0;
    int neg = 0;
    if(x < n && s[x] == '-') {
        neg = 1;
        ++x;
    }
    if(x >= n) return 0;
    int64_t v = 0;
    int overflow = 0;
    for(; x < n; ++x) {
        if(s[x] < '0' || s[x] > '9') return 0;
        v = v * 10 + (s[x] - '0');
        if(v > 4294967296LL) overflow = 1, v = 4294967296LL;
    }
    if(neg) v = -v;
    if(overflow || v < INT32_MIN || v > INT32_MAX) return 0;
    *out = (int32_t)v;
    return 1;
}

void orc_rng_seed_strings(const char* const* seeds, size_t n, uint64_t state[2]) { /* :523-540 */
    uint32_t* u = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    for(size_t x = 0; x < n; ++x) {
        int32_t v;
        size_t len = strlen(seeds[x]);
        u[x] = parse_i32(seeds[x], len, &v) ? (uint32_t)v : orc_fnv1(seeds[x], len);
    }
    orc_rng_seed_u32(u, n, state);
    free(u);
}

/* ---- align_pair.cc:336-385: sample_mdi / sample_mi ---------------------------------- */
static int sample_mdi(float lm, float ld, float li, float p, float* logp) {
    float m = expf(lm), d = expf(ld), n = expf(li);
    float scale = m + d + n;
    p *= scale;
    int st;
    float sc;
    if(p < m) {
        st = 0;
        sc = lm;
    } else if(p < d + m) {
        st = 1;
        sc = ld;
    } else {
        st = 2;
        sc = li;
    }
    *logp = sc - logf(scale);
    return st;
}
static int sample_mi(float lm, float li, float p, float* logp) {
    float m = expf(lm), n = expf(li);
    float scale = m + n;
    p *= scale;
    int st;
    float sc;
    if(p < m) {
        st = 0;
        sc = lm;
    } else {
        st = 2;
        sc = li;
    }
    *logp = sc - logf(scale);
    return st;
}

/* ---- align_pair.cc:401-458: sampleback ---------------------------------------------- */
int orc_sampleback(const float* mch, const float* del, const float* ins, float* const* trans,
                   size_t la, size_t lb, const char* anc, const char* des, size_t k,
                   uint64_t state[2], char* out_a, char* out_b, size_t* out_len, float* score) {
    const size_t len_b = lb + k;
#define AT(m, i, j) (m)[(i) * len_b + (j)]
    size_t i = la + k - 1, j = lb + k - 1, n = 0;
    const size_t cap = la + lb;
    float sc = 0.0f, lp;
    float w = fmax2(fmax2(AT(mch, i, j), AT(del, i, j)), AT(ins, i, j));
    int pick = sample_mdi(AT(mch, i, j) - w, AT(del, i, j) - w, AT(ins, i, j) - w,
                          orc_rng_f24(state), &lp);
    sc += lp;
    while(j > (k - 1) || i > (k - 1)) {
        if(pick == 0) {
            if(n + 1 > cap || i == 0 || j == 0) return ORC_E_ARG;
            out_a[n] = anc[i - k];
            out_b[n] = des[j - k];
            ++n;
            w = AT(mch, i, j);
            pick = sample_mdi(AT(trans[ORC_MCH_MCH], i, j) - w, AT(trans[ORC_DEL_MCH], i, j) - w,
                              AT(trans[ORC_INS_MCH], i, j) - w, orc_rng_f24(state), &lp);
            sc += lp;
            i--;
            j--;
        } else if(pick == 1) {
            if(n + k > cap || i < k) return ORC_E_ARG;
            for(size_t r = i; r > i - k; r--) {
                out_a[n] = anc[r - k];
                out_b[n] = '-';
                ++n;
            }
            w = AT(del, i, j);
            pick = sample_mdi(AT(trans[ORC_MCH_DEL], i, j) - w, AT(trans[ORC_DEL_DEL], i, j) - w,
                              AT(trans[ORC_INS_DEL], i, j) - w, orc_rng_f24(state), &lp);
            sc += lp;
            i -= k;
        } else {
            if(n + k > cap || j < k) return ORC_E_ARG;
            for(size_t q = j; q > j - k; q--) {
                out_a[n] = '-';
                out_b[n] = des[q - k];
                ++n;
            }
            w = AT(ins, i, j);
            pick = sample_mi(AT(trans[ORC_MCH_INS], i, j) - w, AT(trans[ORC_INS_INS], i, j) - w,
                             orc_rng_f24(state), &lp);
            sc += lp;
            j -= k;
        }
    }
#undef AT
    reverse(out_a, n);
    reverse(out_b, n);
    out_a[n] = 0;
    out_b[n] = 0;
    if(out_len) *out_len = n;
    if(score) *score = sc;
    return ORC_OK;
}

/* forward + n x sampleback sharing one RNG stream (align_marginal.cc:585-593).
 * out_a/out_b: n rows of stride la+lb+1.  loglik = plus(M,D,I) at the (adjusted) terminal. */
int orc_sample(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const char* anc,
               const char* des, const float* table, float g, float e, size_t k, uint64_t state[2],
               size_t n, char* out_a, char* out_b, size_t* out_len, float* scores, float* loglik) {
    size_t cells = (la + k) * (lb + k);
    float* buf = (float*)malloc(11 * cells * sizeof(float));
    if(!buf) return ORC_E_NOMEM;
    float* trans[8];
    for(int t = 0; t < 8; ++t) trans[t] = buf + (3 + t) * cells;
    int rc = orc_fill(ORC_LOG, a, la, b, lb, table, g, e, k, buf, buf + cells, buf + 2 * cells, trans);
    if(rc == ORC_OK && loglik) {
        size_t last = cells - 1;
        *loglik = plus3(ORC_LOG, buf[last], buf[cells + last], buf[2 * cells + last]);
    }
    size_t stride = la + lb + 1;
    for(size_t s = 0; rc == ORC_OK && s < n; ++s)
        rc = orc_sampleback(buf, buf + cells, buf + 2 * cells, trans, la, lb, anc, des, k, state,
                            out_a + s * stride, out_b + s * stride, out_len ? out_len + s : NULL,
                            scores ? scores + s : NULL);
    free(buf);
    return rc;
}

/* ---- sequence prep: src/lib/utils.cc ------------------------------------------------ */
/* utils.hpp:54-61 nt16_table: IUPAC code of an ASCII symbol, 16 = invalid, '-' = 15 */
static uint8_t nt16(unsigned char ch) {
    switch(ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case 'R': case 'r': return 4;
    case 'Y': case 'y': return 5;
    case 'M': case 'm': return 6;
    case 'K': case 'k': return 7;
    case 'S': case 's': return 8;
    case 'W': case 'w': return 9;
    case 'B': case 'b': return 10;
    case 'D': case 'd': return 11;
    case 'H': case 'h': return 12;
    case 'V': case 'v': return 13;
    case 'N': case 'n': return 14;
    case '-': return 15;
    default: return 16;
    }
}

/* utils.cc:72-85 cod_int: -1 on anything outside ACGTUacgtu */
int orc_cod_int(const char* codon) {
    int v = 0;
    for(int x = 0; x < 3; ++x) {
        uint8_t c = nt16((unsigned char)codon[x]);
        if(c > 3) return -1;
        v = (v << 2) | c;
    }
    return v;
}

/* utils.cc:1144-1165 (-1: out of range, -2: stop codon) */
int orc_cod64_to_61(int cod) {
    if(cod < 0 || cod > 63) return -1;
    if(cod == 48 || cod == 50 || cod == 56) return -2;
    if(cod < 48) return cod;
    if(cod == 49) return 48;
    if(cod < 57) return cod - 2;
    return cod - 3;
}
/* utils.cc:1195-1211 */
int orc_cod61_to_64(int cod) {
    if(cod < 0 || cod > 60) return -1;
    if(cod < 48) return cod;
    if(cod == 48) return 49;
    if(cod < 54) return cod + 2;
    return cod + 3;
}
/* utils.cc:738-749 get_nuc */
int orc_get_nuc(int cod61, int pos) {
    int c = orc_cod61_to_64(cod61);
    if(c < 0) return -1;
    return (c >> (4 - 2 * pos)) & 3;
}

/* utils.cc:496-520 marginal_seq_encoding, ancestor half */
int orc_encode_anc(const char* anc, size_t n, uint8_t* out) {
    if(n % 3 != 0) return ORC_E_ARG;
    for(size_t i = 0; i < n; i += 3) {
        int cod = orc_cod_int(anc + i);
        if(cod == -1) return ORC_E_AMBIGUOUS;
        if(cod == 48 || cod == 50 || cod == 56) return ORC_E_STOP;
        cod = orc_cod64_to_61(cod) * 3;
        out[i] = (uint8_t)cod;
        out[i + 1] = (uint8_t)(cod + 1);
        out[i + 2] = (uint8_t)(cod + 2);
    }
    return ORC_OK;
}
/* utils.cc:522-526, descendant half (codes 15/16 are emitted unchecked by the reference) */
void orc_encode_des(const char* des, size_t n, uint8_t* out) {
    for(size_t i = 0; i < n; ++i) {
        unsigned char ch = (unsigned char)des[i];
        out[i] = ch < 128 ? nt16(ch) : 16;
    }
}
/* utils.cc:945-967 trim_end_stops predicate */
int orc_has_end_stop(const char* seq, size_t n) {
    if(n < 3) return 0;
    int cod = orc_cod_int(seq + n - 3);
    return cod == 48 || cod == 50 || cod == 56;
}
/* utils.cc:1049 */
float orc_end_stop_gap_score(float g, float e) { return logf(g * e * e); }

/* ---- align_marginal.cc:373-473 alignment_score (+ utils.cc:847-935 process_alignment) */
int orc_alignment_score(const char* aln_a, const char* aln_b, size_t n, const float* table,
                        float g, float e, size_t k, float* score_out) {
    char* rows[2];
    rows[0] = (char*)malloc(2 * (n + 1));
    if(!rows[0]) return ORC_E_NOMEM;
    rows[1] = rows[0] + n + 1;
    memcpy(rows[0], aln_a, n);
    memcpy(rows[1], aln_b, n);
    int stop[2] = {0, 0};
    /* utils.cc:868-899: replace a terminal stop codon (last 3 non-gap symbols) by gaps */
    for(int s = 0; s < 2; ++s) {
        long pos[3];
        long p = (long)n - 1;
        int found = 0;
        for(int q = 2; q >= 0; --q) {
            while(p >= 0 && rows[s][p] == '-') --p;
            if(p < 0) break;
            pos[q] = p--;
            ++found;
        }
        if(found < 3) continue;
        char cod[3] = {rows[s][pos[0]], rows[s][pos[1]], rows[s][pos[2]]};
        int c = orc_cod_int(cod);
        if(c == 48 || c == 50 || c == 56) {
            stop[s] = 1;
            rows[s][pos[0]] = rows[s][pos[1]] = rows[s][pos[2]] = '-';
        }
    }
    /* utils.cc:901-913 expanded cigar, :916-917 strip gaps */
    char* cigar = (char*)malloc(n + 1);
    char* sa = (char*)malloc(2 * (n + 1));
    uint8_t* enc = (uint8_t*)malloc(2 * (n + 1));
    int rc = ORC_OK;
    if(!cigar || !sa || !enc) {
        rc = ORC_E_NOMEM;
        goto done;
    }
    {
        char* sb = sa + n + 1;
        size_t nc = 0, na = 0, nb = 0;
        for(size_t x = 0; x < n; ++x) {
            char ca = rows[0][x], cb = rows[1][x];
            if(ca != '-' && cb != '-') cigar[nc++] = 'M';
            else if(ca != '-') cigar[nc++] = 'D';
            else if(cb != '-') cigar[nc++] = 'I';
            if(ca != '-') sa[na++] = ca;
            if(cb != '-') sb[nb++] = cb;
        }
        if(na % 3 != 0 || na % k != 0 || nb % k != 0) { /* utils.cc:924-935 */
            rc = ORC_E_ARG;
            goto done;
        }
        uint8_t* ea = enc;
        uint8_t* eb = enc + n + 1;
        rc = orc_encode_anc(sa, na, ea);
        if(rc != ORC_OK) goto done;
        orc_encode_des(sb, nb, eb);

        const gapc_t c = gap_consts(g, e);
        int gap_state = 0;
        float score = 0.f;
        size_t nins = 0, ndel = 0, apos = 0, bpos = 0;
        for(size_t x = 0; x < nc; ++x) { /* align_marginal.cc:398-450 */
            if(!gap_state) {
                if(cigar[x] == 'I') {
                    nins++, bpos++, gap_state = 1;
                } else if(cigar[x] == 'D') {
                    ndel++, apos++, gap_state = 1;
                } else {
                    score = ((score + c.ng) + c.ng) + table[(size_t)ea[apos] * 15 + eb[bpos]];
                    apos++, bpos++;
                }
            } else {
                if(cigar[x] == 'I') {
                    nins++, bpos++;
                } else if(cigar[x] == 'D') {
                    ndel++, apos++;
                } else {
                    if(nins == 0)
                        score = (((score + c.ng) + c.go) + powerf(c.ge, ndel - 1)) + c.gs;
                    else if(ndel == 0)
                        score = (((score + c.go) + powerf(c.ge, nins - 1)) + c.gs) + c.ng;
                    else
                        score = ((((score + c.go) + c.go) + powerf(c.ge, nins + ndel - 2)) + c.gs) + c.gs;
                    score = score + table[(size_t)ea[apos] * 15 + eb[bpos]];
                    nins = ndel = 0;
                    gap_state = 0;
                    apos++, bpos++;
                }
            }
        }
        if(!gap_state) { /* :454-470 terminal */
            score = (score + c.ng) + c.ng;
        } else if(nins == 0) {
            score = (((score + c.ng) + c.go) + powerf(c.ge, ndel - 1)) + c.gs;
        } else if(ndel == 0) {
            score = (((score + c.go) + powerf(c.ge, nins - 1)) + c.gs) + c.ng;
        } else {
            score = (((((score + c.go) + c.go) + powerf(c.ge, nins + ndel - 2)) + c.gs) + c.gs) + c.ng;
        }
        if(stop[0] != stop[1]) score += orc_end_stop_gap_score(g, e); /* utils.cc:1049-1062 */
        *score_out = score;
    }
done:
    free(enc);
    free(sa);
    free(cigar);
    free(rows[0]);
    return rc;
}
