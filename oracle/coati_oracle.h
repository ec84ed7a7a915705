/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's marginal Gotoh hot path.
 * Nothing under coati_b200/ may include, link or call this.  See coati_oracle.c. */
#ifndef COATI_ORACLE_H
#define COATI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_E_AMBIGUOUS -1 /* "Ambiguous nucleotides in ancestor/reference." */
#define ORC_E_STOP -2      /* "Early stop codon in ancestor/reference." */
#define ORC_E_ARG -3
#define ORC_E_NOMEM -4

enum { ORC_TROPICAL = 0, ORC_LOG = 1 };

/* order of the 8 transition matrices in `trans` (align_pair.hpp:94-103) */
enum {
    ORC_MCH_MCH = 0, ORC_MCH_DEL, ORC_MCH_INS, ORC_DEL_MCH,
    ORC_DEL_DEL, ORC_INS_MCH, ORC_INS_DEL, ORC_INS_INS
};

float orc_log1p_exp(float x);
float orc_log_sum_exp(float a, float b);
void orc_log1p_exp_array(const float* in, float* out, size_t n);

int orc_fill(int semiring, const uint8_t* a, size_t la, const uint8_t* b, size_t lb,
             const float* table, float g, float e, size_t k, float* mch, float* del, float* ins,
             float* const* trans);
int orc_traceback(const float* mch, const float* del, const float* ins, size_t la, size_t lb,
                  const char* anc, const char* des, float g, float e, size_t k, char* out_a,
                  char* out_b, size_t* out_len, float* score);
int orc_viterbi(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const char* anc,
                const char* des, const float* table, float g, float e, size_t k, char* out_a,
                char* out_b, size_t* out_len, float* score);
/* 1 byte per cell: the argmax decisions traceback would take AT that cell:
 * bits 0-1 = next state after a MATCH step lands here, bits 2-3 after a DELETION step,
 * bit 4 = after an INSERTION step (0 = MATCH, 1 = INSERTION); states 0=M 1=D 2=I. */
int orc_directions(const float* mch, const float* del, const float* ins, size_t la, size_t lb,
                   float g, float e, size_t k, uint8_t* dirs);

int orc_sampleback(const float* mch, const float* del, const float* ins, float* const* trans,
                   size_t la, size_t lb, const char* anc, const char* des, size_t k,
                   uint64_t state[2], char* out_a, char* out_b, size_t* out_len, float* score);
int orc_sample(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const char* anc,
               const char* des, const float* table, float g, float e, size_t k, uint64_t state[2],
               size_t n, char* out_a, char* out_b, size_t* out_len, float* scores, float* loglik);

/* RNG (contrib/random/random.hpp) */
void orc_rng_set_state(uint64_t state[2], uint64_t lo, uint64_t hi);
uint64_t orc_rng_bits(uint64_t state[2]);
float orc_rng_f24(uint64_t state[2]);
void orc_rng_seed_u32(const uint32_t* seeds, size_t n, uint64_t state[2]);
void orc_rng_seed_strings(const char* const* seeds, size_t n, uint64_t state[2]);
uint32_t orc_fnv1(const char* s, size_t n);

/* sequence prep (src/lib/utils.cc) */
int orc_cod_int(const char* codon);
int orc_cod64_to_61(int cod);
int orc_cod61_to_64(int cod);
int orc_get_nuc(int cod61, int pos);
int orc_encode_anc(const char* anc, size_t n, uint8_t* out);
void orc_encode_des(const char* des, size_t n, uint8_t* out);
int orc_has_end_stop(const char* seq, size_t n);
float orc_end_stop_gap_score(float g, float e);

/* alignment_score (align_marginal.cc:373-473); seqs are the gapped alignment rows. */
int orc_alignment_score(const char* aln_a, const char* aln_b, size_t n, const float* table,
                        float g, float e, size_t k, float* score);

/* long_pair.c: exact O(La + Lb)-memory validators for pairs the full matrices cannot hold.
 * orc_path_score: the alignment rows re-scored through forward_impl's own transition terms
 * (align_pair.cc:81-138); orc_viterbi_score: score-only rolling-row Viterbi, k = 1, `threads` strips. */
int orc_path_score(const char* aln_a, const char* aln_b, size_t n, const uint8_t* a, size_t la,
                   const uint8_t* b, size_t lb, const float* table, float g, float e, size_t k,
                   float* score);
int orc_viterbi_score(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const float* table,
                      float g, float e, size_t k, int threads, float* score);

#ifdef __cplusplus
}
#endif
#endif
