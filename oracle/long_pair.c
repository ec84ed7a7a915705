/* TEST INFRASTRUCTURE ONLY -- exact validators for pairs too long for the full-matrix oracle.
 *
 * The full restatement (coati_oracle.c: orc_fill) and the reference itself keep three (La+k) x (Lb+k)
 * float matrices: 19 GB at 40k, 307 GB at 160k.  These two functions give the same float32 bits with
 * O(La + Lb) memory:
 *
 *   orc_path_score     re-scores a returned alignment through forward_impl's own recurrence
 *                      (/root/reference/src/lib/align_pair.cc:81-91 margins, :97-124 body, :130-138
 *                      terminal) in its own left-to-right association.  The value of the Viterbi
 *                      optimum IS the chain of float additions along the arg-max path (max selects one
 *                      of its operands unchanged), so for k = 1 -- where traceback's comparison
 *                      expressions (:275-296) are the fill's own terms -- the result must equal the
 *                      Viterbi score bit for bit.  O(La + Lb) time.
 *   orc_viterbi_score  score-only Viterbi fill with rolling rows (same cell expressions, same order of
 *                      the max folds), k = 1, rows cut into one strip per thread that run as a
 *                      pipeline over column blocks.  Gives max(M', D', I') of the adjusted terminal
 *                      cell (:130-138, :265) = the score traceback<S> reports.
 *
 * Both are checked bit for bit against orc_fill / orc_viterbi in tests/test_oracle_long.py.
 */
#include "coati_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define LOWEST (-FLT_MAX)

static inline float fmax2(float x, float y) { return (x < y) ? y : x; } /* std::max */

typedef struct {
    float ng, gs, go, ge;
} gapc_t;

static gapc_t gap_consts(float g, float e) { /* align_pair.cc:66-69 */
    gapc_t c;
    c.ng = log1pf(-g);
    c.gs = log1pf(-e);
    c.go = logf(g);
    c.ge = logf(e);
    return c;
}

/* ------------------------------------------------------------------------------------------------
 * orc_path_score: aln_a / aln_b are the two alignment rows (n columns, '-' = gap); a / b the encoded
 * sequences (utils.cc:496-528) the rows were made from.  Walks the path from the origin and applies,
 * per move, the one transition term of forward_impl that the move uses:
 *   ->M  m2m = ((M+ng)+ng)+s   d2m = (D+gs)+s        i2m = ((I+gs)+ng)+s        (:98-102)
 *   ->D  m2d = ((M+ng)+go)+gk1 d2d = D+gk            i2d = ((I+gs)+go)+gk1      (:106-112)
 *   ->I  m2i = (M+go)+gk1      i2i = I+gk            (no D->I, :30-43)          (:115-118)
 * margin cells hold their initial values instead (:84-90).  Returns ORC_E_ARG when the rows are not
 * an alignment of a and b in units of k, or use a D->I transition. */
int orc_path_score(const char* aln_a, const char* aln_b, size_t n, const uint8_t* a, size_t la,
                   const uint8_t* b, size_t lb, const float* table, float g, float e, size_t k,
                   float* score) {
    if(k == 0 || !score) return ORC_E_ARG;
    const gapc_t c = gap_consts(g, e);
    const float gk1 = c.ge * (float)(k - 1), gk = c.ge * (float)k; /* semiring.hpp:109-111 */
    const size_t start = k - 1;
    size_t i = start, j = start; /* matrix indices (align_pair.cc:72-79) */
    int st = 0;                  /* 0 M, 1 D, 2 I */
    float v = 0.0f;              /* mch(start, start) = one() (:82) */
    size_t x = 0;
    while(x < n) {
        const int ga = aln_a[x] == '-', gb = aln_b[x] == '-';
        if(ga && gb) return ORC_E_ARG;
        if(!ga && !gb) { /* MATCH: one column */
            if(i + 1 - k >= la || j + 1 - k >= lb) return ORC_E_ARG;
            const float s = table[(size_t)a[i + 1 - k] * 15 + b[j + 1 - k]];
            if(st == 0) v = ((v + c.ng) + c.ng) + s;
            else if(st == 1) v = (v + c.gs) + s;
            else v = ((v + c.gs) + c.ng) + s;
            ++i, ++j, ++x;
            st = 0;
        } else if(gb) { /* DELETION: k columns (a, '-') */
            for(size_t q = 0; q < k; ++q)
                if(x + q >= n || aln_b[x + q] != '-' || aln_a[x + q] == '-') return ORC_E_ARG;
            if(i + k - start > la) return ORC_E_ARG;
            if(j == start) { /* left margin (:84-87): the cell holds its initial value */
                if(st == 2) return ORC_E_ARG;
                v = (c.ng + c.go) + c.ge * (float)(i + k - 1);
            } else if(st == 0) v = ((v + c.ng) + c.go) + gk1;
            else if(st == 1) v = v + gk;
            else v = ((v + c.gs) + c.go) + gk1;
            i += k, x += k;
            st = 1;
        } else { /* INSERTION: k columns ('-', b) */
            for(size_t q = 0; q < k; ++q)
                if(x + q >= n || aln_a[x + q] != '-' || aln_b[x + q] == '-') return ORC_E_ARG;
            if(j + k - start > lb) return ORC_E_ARG;
            if(st == 1) return ORC_E_ARG; /* insertion-before-deletion rule */
            if(i == start) v = c.go + c.ge * (float)(j + k - 1); /* top margin (:88-90) */
            else if(st == 0) v = (v + c.go) + gk1;
            else v = v + gk;
            j += k, x += k;
            st = 2;
        }
    }
    if(i - start != la || j - start != lb) return ORC_E_ARG;
    /* terminal state (:130-138) */
    if(st == 0) v = (v + c.ng) + c.ng;
    else if(st == 1) v = v + c.gs;
    else v = (v + c.gs) + c.ng;
    *score = v;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * orc_viterbi_score: rolling-row Viterbi (k = 1).  Rows 1..la are cut into `threads` strips; strip t
 * sweeps its rows over column blocks of CB columns and hands its bottom row to strip t + 1 through
 * bnd[t + 1]; done[t] counts the column blocks strip t has published. */
enum { CB = 2048 };

typedef struct {
    const uint8_t *a, *b;
    size_t la, lb;
    const float* table;
    gapc_t c;
    int nthreads;
    size_t nblocks;
    float** bnd;          /* bnd[t]: row above strip t, 3 * (lb + 1) floats: M, D, I planes */
    atomic_size_t* done;  /* done[t]: column blocks of bnd[t + 1] that are complete */
} shared_t;

typedef struct {
    shared_t* sh;
    int t;
    int rc;
} job_t;

static void* strip_main(void* arg) {
    job_t* job = (job_t*)arg;
    shared_t* sh = job->sh;
    const int t = job->t;
    const gapc_t c = sh->c;
    const size_t la = sh->la, lb = sh->lb;
    const size_t r0 = la * (size_t)t / (size_t)sh->nthreads + 1;        /* first row of the strip */
    const size_t r1 = la * (size_t)(t + 1) / (size_t)sh->nthreads;      /* last row (inclusive) */
    const size_t nrows = r1 >= r0 ? r1 - r0 + 1 : 0;
    const size_t ld = lb + 1;
    const float *topM = sh->bnd[t], *topD = topM + ld, *topI = topD + ld;
    float *botM = sh->bnd[t + 1], *botD = botM + ld, *botI = botD + ld;
    /* state of every strip row at the last column of the previous block */
    float* leftM = (float*)malloc(3 * (nrows + 1) * sizeof(float));
    float* prev = (float*)malloc(6 * (CB + 1) * sizeof(float));
    if(!leftM || !prev) {
        free(leftM), free(prev);
        job->rc = ORC_E_NOMEM;
        /* keep the pipeline alive so the other strips terminate */
        atomic_store_explicit(&sh->done[t], sh->nblocks, memory_order_release);
        return NULL;
    }
    float *leftD = leftM + nrows + 1, *leftI = leftD + nrows + 1;
    float *pM = prev, *pD = pM + CB + 1, *pI = pD + CB + 1, *qM = pI + CB + 1, *qD = qM + CB + 1,
          *qI = qD + CB + 1;
    /* column 0 (align_pair.cc:84-87): only del is finite; row r0 - 1 is slot 0 */
    for(size_t x = 0; x <= nrows; ++x) {
        const size_t r = r0 - 1 + x;
        leftM[x] = r == 0 ? 0.0f : LOWEST;
        leftD[x] = r == 0 ? LOWEST : (c.ng + c.go) + c.ge * (float)(r - 1);
        leftI[x] = LOWEST;
    }
    for(size_t blk = 0; blk < sh->nblocks; ++blk) {
        const size_t c0 = blk * CB + 1, c1 = c0 + CB - 1 < lb ? c0 + CB - 1 : lb; /* columns [c0, c1] */
        const size_t w = c1 - c0 + 1;
        if(t > 0)
            while(atomic_load_explicit(&sh->done[t - 1], memory_order_acquire) <= blk) sched_yield();
        /* prev row = row above the strip over columns c0-1 .. c1 (slot 0 = column c0 - 1) */
        pM[0] = leftM[0], pD[0] = leftD[0], pI[0] = leftI[0];
        memcpy(pM + 1, topM + c0, w * sizeof(float));
        memcpy(pD + 1, topD + c0, w * sizeof(float));
        memcpy(pI + 1, topI + c0, w * sizeof(float));
        leftM[0] = pM[w], leftD[0] = pD[w], leftI[0] = pI[w];
        for(size_t x = 1; x <= nrows; ++x) {
            const size_t r = r0 - 1 + x;
            const float* trow = sh->table + (size_t)sh->a[r - 1] * 15;
            const uint8_t* bb = sh->b + (c0 - 1);
            float cm = leftM[x], ci = leftI[x]; /* M, I of this row at the previous column */
            qM[0] = cm, qD[0] = leftD[x], qI[0] = ci;
            for(size_t y = 1; y <= w; ++y) {
                const float s = trow[bb[y - 1]];
                const float m2m = ((pM[y - 1] + c.ng) + c.ng) + s; /* :98-102 */
                const float d2m = (pD[y - 1] + c.gs) + s;
                const float i2m = ((pI[y - 1] + c.gs) + c.ng) + s;
                const float m2d = ((pM[y] + c.ng) + c.go) + c.ge * 0.0f; /* :106-112, gk1 = ge * 0 */
                const float i2d = ((pI[y] + c.gs) + c.go) + c.ge * 0.0f;
                const float d2d = pD[y] + c.ge * 1.0f;
                const float m2i = (cm + c.go) + c.ge * 0.0f; /* :115-118 */
                const float i2i = ci + c.ge * 1.0f;
                cm = fmax2(fmax2(m2m, d2m), i2m); /* :119-121 */
                ci = fmax2(m2i, i2i);
                qM[y] = cm;
                qD[y] = fmax2(fmax2(m2d, d2d), i2d);
                qI[y] = ci;
            }
            leftM[x] = qM[w], leftD[x] = qD[w], leftI[x] = qI[w];
            float* sw;
            sw = pM, pM = qM, qM = sw;
            sw = pD, pD = qD, qD = sw;
            sw = pI, pI = qI, qI = sw;
        }
        /* bottom row of the strip over this block */
        if(nrows) {
            memcpy(botM + c0, pM + 1, w * sizeof(float));
            memcpy(botD + c0, pD + 1, w * sizeof(float));
            memcpy(botI + c0, pI + 1, w * sizeof(float));
        } else {
            memcpy(botM + c0, topM + c0, w * sizeof(float));
            memcpy(botD + c0, topD + c0, w * sizeof(float));
            memcpy(botI + c0, topI + c0, w * sizeof(float));
        }
        atomic_store_explicit(&sh->done[t], blk + 1, memory_order_release);
    }
    /* column 0 of a boundary row is only ever read as the terminal cell of a pair with lb == 0 */
    if(sh->nblocks == 0) botM[0] = leftM[nrows], botD[0] = leftD[nrows], botI[0] = leftI[nrows];
    free(leftM);
    free(prev);
    job->rc = ORC_OK;
    return NULL;
}

int orc_viterbi_score(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const float* table,
                      float g, float e, size_t k, int threads, float* score) {
    if(k != 1 || !score || threads < 1) return ORC_E_ARG;
    if((size_t)threads > la) threads = la ? (int)la : 1;
    shared_t sh;
    sh.a = a, sh.b = b, sh.la = la, sh.lb = lb, sh.table = table;
    sh.c = gap_consts(g, e);
    sh.nthreads = threads;
    sh.nblocks = (lb + CB - 1) / CB;
    const size_t ld = lb + 1;
    sh.bnd = (float**)calloc((size_t)threads + 1, sizeof(float*));
    sh.done = (atomic_size_t*)calloc((size_t)threads, sizeof(atomic_size_t));
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    pthread_t* tid = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    int rc = ORC_OK;
    if(!sh.bnd || !sh.done || !jobs || !tid) rc = ORC_E_NOMEM;
    for(int t = 0; rc == ORC_OK && t <= threads; ++t) {
        sh.bnd[t] = (float*)malloc(3 * ld * sizeof(float));
        if(!sh.bnd[t]) rc = ORC_E_NOMEM;
    }
    if(rc == ORC_OK) {
        /* top margin row (:82, :88-90) */
        float *M = sh.bnd[0], *D = M + ld, *I = D + ld;
        for(size_t j = 0; j <= lb; ++j) {
            M[j] = j == 0 ? 0.0f : LOWEST;
            D[j] = LOWEST;
            I[j] = j == 0 ? LOWEST : sh.c.go + sh.c.ge * (float)(j - 1);
        }
        for(int t = 0; t < threads; ++t) {
            atomic_init(&sh.done[t], 0);
            jobs[t].sh = &sh, jobs[t].t = t, jobs[t].rc = ORC_OK;
        }
        int started = 0;
        for(int t = 0; t < threads; ++t) {
            if(pthread_create(&tid[t], NULL, strip_main, &jobs[t]) != 0) {
                /* run the remaining strips on this thread, in order */
                for(int u = t; u < threads; ++u) strip_main(&jobs[u]);
                break;
            }
            ++started;
        }
        for(int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
        for(int t = 0; t < threads; ++t)
            if(jobs[t].rc != ORC_OK) rc = jobs[t].rc;
        if(rc == ORC_OK) {
            const float *bM = sh.bnd[threads], *bD = bM + ld, *bI = bD + ld;
            /* terminal adjust (:130-138) and score = max(M, D, I) (:265) */
            const float m = (bM[lb] + sh.c.ng) + sh.c.ng;
            const float i = (bI[lb] + sh.c.gs) + sh.c.ng;
            const float d = bD[lb] + sh.c.gs;
            *score = fmax2(fmax2(m, d), i);
        }
    }
    if(sh.bnd)
        for(int t = 0; t <= threads; ++t) free(sh.bnd[t]);
    free(sh.bnd), free(sh.done), free(jobs), free(tid);
    return rc;
}
