/*
 * coati_gpu.h -- C ABI of the B200-native marginal Gotoh hot path (libcoati_gpu.so).
 *
 * The reference (CartwrightLab/coati) has no plugin/FFI seam; its only boundary for this path is
 * the C++ free-function family in src/include/coati/align_pair.hpp:157-182, called from
 * src/lib/align_marginal.cc:71,80 (alignpair), :586,590 (sample) and src/lib/align_msa.cc:307-308.
 * Each entry point below names the reference call(s) it replaces.  Plain pointers and sizes only;
 * no C++ or torch types.  All functions return COATI_GPU_OK (0) or a negative error code;
 * coati_gpu_strerror() gives the message the C++ wrapper rethrows with.
 *
 * Conventions
 *   - a   : ancestor encoded as codon61*3+phase in [0,183)   (utils.cc:496-520)
 *   - b   : descendant encoded as IUPAC code in [0,15)        (utils.cc:522-526, utils.hpp:54-61)
 *   - anc/des : the raw (case-preserved, end-stop-trimmed) symbols the alignment rows are made of
 *   - table : 183 x 15 float32 row-major log-odds (mutation_coati.cc:164-202)
 *   - scores are float32 and bit-identical to the reference's for Viterbi
 *   - a context is bound to one GPU and is not thread-safe; use one per host thread / device
 *   - there is NO CPU fallback: every entry point fails with COATI_GPU_E_CUDA when no device runs it
 */
#ifndef COATI_GPU_H
#define COATI_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COATI_GPU_OK 0
#define COATI_GPU_E_CUDA -1      /* CUDA runtime/driver failure or no usable device            */
#define COATI_GPU_E_ARG -2       /* invalid argument (null pointer, k == 0, model not set ...) */
#define COATI_GPU_E_NOMEM -3     /* device or host allocation failed (reference: std::bad_alloc) */
#define COATI_GPU_E_SYMBOL -4    /* encoded symbol outside the table (a >= 183 or b >= 15)      */
#define COATI_GPU_E_LENGTH -5    /* La % k != 0 or Lb % k != 0 (unreachable terminal cell)      */
#define COATI_GPU_E_AMBIGUOUS -6 /* "Ambiguous nucleotides in ancestor/reference."              */
#define COATI_GPU_E_STOP -7      /* "Early stop codon in ancestor/reference."                   */
#define COATI_GPU_E_INTERNAL -8  /* traceback left the lattice (would be UB in the reference)   */

typedef struct coati_gpu_ctx coati_gpu_ctx;
typedef struct coati_gpu_batch coati_gpu_batch;
typedef struct coati_gpu_forward_t coati_gpu_forward_t;

/* ---- context ------------------------------------------------------------------------------- */
int coati_gpu_init(int device, coati_gpu_ctx** ctx);
void coati_gpu_shutdown(coati_gpu_ctx* ctx);
const char* coati_gpu_strerror(int code);
/* last CUDA error string seen by this context (diagnostics) */
const char* coati_gpu_last_cuda_error(coati_gpu_ctx* ctx);
/* the cudaStream_t all work of this context is enqueued on (for external CUDA-event timing) */
void* coati_gpu_stream(coati_gpu_ctx* ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t coati_gpu_launch_count(coati_gpu_ctx* ctx);
/* bytes the batch calls of this context moved host->device and device->host since creation (bench.py's
 * h2d_bytes_per_step / d2h_bytes_per_step): copies as enqueued, and -- where the rows are written straight into
 * page-locked caller arenas, see coati_gpu_host_alloc -- the row bytes actually written (length + terminator) */
void coati_gpu_transfer_bytes(coati_gpu_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
int coati_gpu_device_info(coati_gpu_ctx* ctx, int* sm_count, int* clock_khz, size_t* free_bytes,
                          size_t* total_bytes);

/* ---- model -----------------------------------------------------------------------------------
 * Replaces what forward_impl/traceback read from `alignment_t`: aln.subst_matrix, aln.gap.open,
 * aln.gap.extend, aln.gap.len (align_pair.cc:66-72, 253-256).  log(1-g), log(1-e), log(g), log(e),
 * log(e)*(k-1), log(e)*k are derived on the host with libm exactly as align_pair.cc:66-69 does. */
int coati_gpu_set_model(coati_gpu_ctx* ctx, const float* table, float gap_open, float gap_extend,
                        uint32_t gap_len);

/* Several substitution models at once (tables: n_models x 183 x 15), gap parameters shared: the msa
 * driver aligns every leaf with its own branch length (align_msa.cc:285-318: set_subst per leaf). */
int coati_gpu_set_models(coati_gpu_ctx* ctx, uint32_t n_models, const float* tables, float gap_open,
                         float gap_extend, uint32_t gap_len);

/* ---- Viterbi, one pair ------------------------------------------------------------------------
 * = viterbi_mem (align_pair.cc:195-198) + traceback_viterbi (:319-323) as marg_alignment calls
 * them (align_marginal.cc:69-80).  out_a/out_b: caller buffers of >= La+Lb+1 bytes (NUL added). */
int coati_gpu_viterbi(coati_gpu_ctx* ctx, const uint8_t* a, size_t La, const uint8_t* b, size_t Lb,
                      const char* anc, const char* des, char* out_a, char* out_b, size_t* out_len,
                      float* score);

/* ---- Viterbi, batch of independent pairs (CSR) ------------------------------------------------
 * The same two reference calls per pair, for npairs pairs (the batch the msa driver issues per
 * leaf, align_msa.cc:285-318, and BASELINE configs 4/5).  Pair p owns a_all[a_off[p]..a_off[p+1])
 * and b_all[b_off[p]..b_off[p+1]); anc_all/des_all share those offsets.  Outputs for pair p start
 * at byte a_off[p] + b_off[p] + p of out_a/out_b (capacity La+Lb+1, NUL-terminated); out_len[p],
 * score[p], status[p] (COATI_GPU_OK or a per-pair error) are in input order.
 * Returns COATI_GPU_OK when the batch ran, even if some pairs carry a non-zero status. */
int coati_gpu_viterbi_batch(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all,
                            const uint64_t* a_off, const uint8_t* b_all, const uint64_t* b_off,
                            const char* anc_all, const char* des_all, char* out_a, char* out_b,
                            uint64_t* out_len, float* score, int32_t* status);

/* Same, every pair aligned under its own model of coati_gpu_set_models (model_idx[p] < n_models). */
int coati_gpu_viterbi_batch_models(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all,
                                   const uint64_t* a_off, const uint8_t* b_all, const uint64_t* b_off,
                                   const char* anc_all, const char* des_all, const uint32_t* model_idx,
                                   char* out_a, char* out_b, uint64_t* out_len, float* score,
                                   int32_t* status);

/* ---- alignpair, batch of RAW pairs -----------------------------------------------------------------
 * = marg_alignment (align_marginal.cc:44-88) per pair, minus file I/O: the length checks of
 * process_marginal (La % 3, La % k, Lb % k, on the untrimmed sequences, utils.cc:819-837), trim_end_stops
 * (utils.cc:945-967), marginal_seq_encoding (utils.cc:496-528, here on the device), viterbi_mem +
 * traceback_viterbi, restore_end_stops (utils.cc:1044-1063).  Inputs are the raw symbols only (CSR);
 * outputs as for coati_gpu_viterbi_batch (rows at anc_off[p] + des_off[p] + p).  Per-pair status:
 * COATI_GPU_E_LENGTH / E_AMBIGUOUS / E_STOP carry the reference's exception messages; E_SYMBOL marks a
 * descendant symbol outside the IUPAC table (undefined behaviour upstream). */
int coati_gpu_alignpair_batch(coati_gpu_ctx* ctx, size_t npairs, const char* anc_all,
                              const uint64_t* anc_off, const char* des_all, const uint64_t* des_off,
                              char* out_a, char* out_b, uint64_t* out_len, float* score,
                              int32_t* status);

/* ---- the same call over several GPUs of one box (north_star (4)) ----------------------------
 * The reference is single-threaded and has no analogue; its natural batch client is the per-leaf loop of
 * align_msa.cc:285-318.  ctxs: one context per device, all with the same model set.  The batch is cut into
 * contiguous chunks, ordered heaviest first (longest-processing-time on the sum of La * Lb), and one host
 * thread per context pipelines whatever chunk it pops next from the shared queue; results land in the
 * caller's arenas in input order, exactly as coati_gpu_alignpair_batch leaves them (no collective). */
int coati_gpu_multi_alignpair_batch(coati_gpu_ctx* const* ctxs, int n_ctx, size_t npairs, const char* anc_all,
                                    const uint64_t* anc_off, const char* des_all, const uint64_t* des_off,
                                    char* out_a, char* out_b, uint64_t* out_len, float* score,
                                    int32_t* status);
/* For callers that run one PROCESS per device: the same chunks given to n_shards shards by greedy
 * longest-processing-time (deterministic: every process computes the same plan from the offsets), and the
 * alignpair call restricted to a shard's ranges [first[j], last[j]) of the one CSR batch.  plan_shards
 * returns the number of ranges written, 0 if max_ranges is too small (2 * (npairs / 8192 + n_shards + 2) always suffices). */
size_t coati_gpu_plan_shards(size_t npairs, const uint64_t* a_off, const uint64_t* b_off, uint32_t n_shards,
                             size_t max_ranges, uint64_t* range_first, uint64_t* range_last,
                             uint32_t* range_shard);
int coati_gpu_alignpair_batch_ranges(coati_gpu_ctx* ctx, size_t npairs, const char* anc_all,
                                     const uint64_t* anc_off, const char* des_all, const uint64_t* des_off,
                                     char* out_a, char* out_b, uint64_t* out_len, float* score, int32_t* status,
                                     size_t n_ranges, const uint64_t* first, const uint64_t* last);

/* Page-locked host memory for the arenas of the batch calls (they overlap copies and kernels only from and
 * to pinned memory): allocate here, or register memory the caller already owns.  The memory is also mapped for the
 * device: a call that is one share of a multi-device batch (coati_gpu_multi_alignpair_batch with several
 * contexts, coati_gpu_alignpair_batch_ranges on part of a batch) has the GPU write the rows into such arenas
 * itself, used bytes only; bytes of a slot beyond a row's terminator are then left as the caller had them
 * (with the copy they are scratch).  COATI_GPU_ROWS_DIRECT=1 / 0 forces that / the copy for every call. */
void* coati_gpu_host_alloc(size_t bytes);
void coati_gpu_host_free(void* p);
int coati_gpu_host_register(void* p, size_t bytes);
int coati_gpu_host_unregister(void* p);

/* Staged form of the same call, so the device-resident part can be timed alone:
 *   create  : host-side plan (length-binned LPT order, direction-buffer chunks) + device buffers
 *   upload  : H2D of sequences        run : fill + traceback kernels only (async on the stream)
 *   download: D2H of rows/scores      destroy
 * `run` may be repeated; it recomputes everything from the device-resident inputs. */
int coati_gpu_batch_create(coati_gpu_ctx* ctx, size_t npairs, const uint64_t* a_off,
                           const uint64_t* b_off, coati_gpu_batch** batch);
int coati_gpu_batch_upload(coati_gpu_batch* batch, const uint8_t* a_all, const uint8_t* b_all,
                           const char* anc_all, const char* des_all);
int coati_gpu_batch_run(coati_gpu_batch* batch);
int coati_gpu_batch_download(coati_gpu_batch* batch, char* out_a, char* out_b, uint64_t* out_len,
                             float* score, int32_t* status);
/* counters of the last run: lattice cells filled, direction bytes written, kernels launched */
int coati_gpu_batch_stats(coati_gpu_batch* batch, uint64_t* cells, uint64_t* dir_bytes,
                          uint64_t* launches, uint64_t* chunks);
/* device time of the last run split by kernel family, from CUDA events recorded on the batch's
 * streams (synchronises them): fill = every fill launch; traceback = the walks of each chunk (for long
 * pairs including their segment-parallel row expansion); compact = the row expansion of each chunk;
 * fill_launches = fill kernels launched */
int coati_gpu_batch_timing(coati_gpu_batch* batch, double* fill_ms, double* traceback_ms,
                           double* compact_ms, uint64_t* fill_launches);
/* device-resident outputs of the last run (for a NCCL gather of per-rank results): the two row
 * arenas (out_bytes each, same layout as coati_gpu_viterbi_batch's out_a/out_b) and the per-pair
 * result records {float term[3]; float score; u32 len; u32 start; i32 status; u32 pad}. */
int coati_gpu_batch_device_buffers(coati_gpu_batch* batch, void** out_a, void** out_b,
                                   uint64_t* out_bytes, void** results, uint64_t* result_bytes);
void coati_gpu_batch_destroy(coati_gpu_batch* batch);

/* ---- Forward fill + seeded stochastic sampleback ----------------------------------------------
 * coati_gpu_forward   = forward (align_pair.cc:149-152: forward_impl<semiring::log, align_pair_work_t>).
 *                       The handle owns the device-resident state matrices (opaque work object: no
 *                       caller of the reference ever reads align_pair_work_t).
 * coati_gpu_sampleback = n consecutive sampleback calls (align_pair.cc:401-458) on one RNG stream, as
 *                       marg_sample's loop does (align_marginal.cc:589-593).  rng_state is the raw
 *                       128-bit Lehmer64Fast state {low, high 64 bits} (random.hpp:99,115 GetState /
 *                       Seed(state_type)), in-out, so a host `Random` stays in lock-step.
 *                       out_a/out_b: n rows of stride La+Lb+1 bytes, NUL-terminated; scores[n].
 * coati_gpu_forward_terminal: the adjusted terminal M, D, I (align_pair.cc:130-138); the forward
 *                       log-likelihood is log_sum_exp of the three.
 * A handle belongs to the model it was filled under: after coati_gpu_set_model(s) on its context,
 * coati_gpu_sampleback on an older handle returns COATI_GPU_E_ARG (fill again). */
int coati_gpu_forward(coati_gpu_ctx* ctx, const uint8_t* a, size_t La, const uint8_t* b, size_t Lb,
                      coati_gpu_forward_t** handle);
int coati_gpu_forward_terminal(coati_gpu_forward_t* handle, float term[3], float* fill_ms);
int coati_gpu_sampleback(coati_gpu_forward_t* handle, const char* anc, const char* des,
                         uint64_t rng_state[2], size_t n, char* out_a, char* out_b, size_t* out_len,
                         float* scores, float* sample_ms);
void coati_gpu_forward_free(coati_gpu_forward_t* handle);

/* Batch forms (the throughput path of `sample` / of a per-leaf sampling driver): forward for npairs pairs
 * (CSR, as coati_gpu_viterbi_batch), one handle for all matrices; the adjusted terminal M, D, I (3 per pair)
 * and the forward log-likelihood log_sum_exp(log_sum_exp(M, D), I) per pair; and n consecutive samplebacks
 * per pair, every pair on its own RNG stream (rng_states: 2 x uint64 per pair, in-out).  Rows of sample s
 * of pair p start at byte n * (a_off[p] + b_off[p] + p) + s * (La_p + Lb_p + 1), offsets relative to the
 * first pair's, NUL-terminated; out_len / scores are [p * n + s]. */
int coati_gpu_forward_batch(coati_gpu_ctx* ctx, size_t npairs, const uint8_t* a_all, const uint64_t* a_off,
                            const uint8_t* b_all, const uint64_t* b_off, coati_gpu_forward_t** handle);
int coati_gpu_forward_batch_terminal(coati_gpu_forward_t* handle, float* term, float* loglik,
                                     float* fill_ms);
int coati_gpu_sampleback_batch(coati_gpu_forward_t* handle, const char* anc_all, const char* des_all,
                               uint64_t* rng_states, size_t n, char* out_a, char* out_b,
                               uint64_t* out_len, float* scores, float* sample_ms);
/* parity aids: the three state matrices in lattice coordinates, (La+1) x (Lb+1) row-major (the
 * reference's (La+k) x (Lb+k) matrices without their k-1 padding rows/columns); and the device
 * twins of libm (op 0: expf, 1: logf, 2: log1pf, 3: log1p_exp of utils.hpp:134-146, 4: the branch-free
 * log1p_exp the banded Forward kernel uses, defined for x <= 0) on an array. */
int coati_gpu_forward_matrices(coati_gpu_forward_t* handle, float* mch, float* del, float* ins);
int coati_gpu_libm_eval(coati_gpu_ctx* ctx, int op, const float* in, float* out, size_t n);

/* ---- debugging / parity aid -------------------------------------------------------------------
 * Fill one pair and return the decision byte of every body cell, row-major La x Lb
 * (bits 0-1: next state after a MATCH step lands on the cell, bits 2-3: after a DELETION step,
 * bit 4: after an INSERTION step; 0 = M, 1 = D, 2 = I), decoded from whichever packed stream the
 * fill kernel for this pair emits in place of the reference's three score matrices, plus the
 * Viterbi score.  For cell-level parity tests against traceback's expressions (align_pair.cc:275-296). */
int coati_gpu_viterbi_directions(coati_gpu_ctx* ctx, const uint8_t* a, size_t La, const uint8_t* b,
                                 size_t Lb, uint8_t* dirs, float* score);

#ifdef __cplusplus
}
#endif
#endif /* COATI_GPU_H */
