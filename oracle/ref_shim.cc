// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// C-ABI shim over the UNMODIFIED reference sources.  This file holds no reference
// code: it #includes the reference headers where they lie under /root/reference and
// is linked (oracle/Makefile) against the reference's own translation units
//   src/lib/align_pair.cc, contrib/random/random.cpp, contrib/fstlib/*.cc
// compiled in place.  Output: oracle/_ref/libcoati_ref.so (git-ignored).
//
// It exists so tests/ can pin oracle/coati_oracle.c (the C restatement) and the CUDA
// path against the reference's real `viterbi_mem`/`traceback_viterbi`/`forward`/
// `sampleback` (src/include/coati/align_pair.hpp:157-182), and so bench.py can time
// the reference CPU path (`cpu_baseline.kind == "reference"`).
#include <coati/align_pair.hpp>

#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

using coati::seq_view_t;

void load_model(coati::alignment_t& aln, const float* table, float g, float e, size_t k) {
    aln.subst_matrix = coati::Matrixf(183, 15, table, table + 183 * 15);
    aln.gap.open = g;
    aln.gap.extend = e;
    aln.gap.len = k;
}

void copy_matrix(const coati::Matrixf& m, float* out) {
    if(out == nullptr) return;
    for(size_t i = 0; i < m.rows(); ++i)
        for(size_t j = 0; j < m.cols(); ++j) out[i * m.cols() + j] = m(i, j);
}

void copy_out(coati::alignment_t& aln, char* out_a, char* out_b, size_t* out_len, float* score) {
    const std::string& sa = aln.data.seqs[0];
    const std::string& sb = aln.data.seqs[1];
    if(out_a) { std::memcpy(out_a, sa.data(), sa.size()); out_a[sa.size()] = 0; }
    if(out_b) { std::memcpy(out_b, sb.data(), sb.size()); out_b[sb.size()] = 0; }
    if(out_len) *out_len = sa.size();
    if(score) *score = aln.data.score;
}

fragmites::random::Random make_rng(const uint64_t state[2]) {
    __uint128_t s = (static_cast<__uint128_t>(state[1]) << 64) | state[0];
    fragmites::random::Random r;
    r.Seed(s);  // Lehmer64Fast::Seed(state_type) -> SetState (state | 1)
    return r;
}

void save_rng(const fragmites::random::Random& r, uint64_t state[2]) {
    __uint128_t s = r.GetState();
    state[0] = static_cast<uint64_t>(s);
    state[1] = static_cast<uint64_t>(s >> 64);
}

}  // namespace

extern "C" {

// viterbi_mem + traceback_viterbi (align_marginal.cc:69-80).  a/b: encoded, anc/des raw.
// mch/del/ins (optional): (la+k)*(lb+k) row-major dumps of the three score matrices.
int coati_ref_viterbi(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const char* anc,
                      const char* des, const float* table, float g, float e, size_t k,
                      char* out_a, char* out_b, size_t* out_len, float* score, float* mch,
                      float* del, float* ins) {
    try {
        coati::alignment_t aln;
        load_model(aln, table, g, e, k);
        coati::align_pair_work_mem_t work;
        coati::viterbi_mem(work, seq_view_t(a, la), seq_view_t(b, lb), aln);
        copy_matrix(work.mch, mch);
        copy_matrix(work.del, del);
        copy_matrix(work.ins, ins);
        coati::traceback_viterbi(work, std::string(anc, la), std::string(des, lb), aln, k);
        copy_out(aln, out_a, out_b, out_len, score);
    } catch(const std::bad_alloc&) {
        return -2;
    } catch(const std::exception&) {
        return -1;
    }
    return 0;
}

// forward (11-matrix log-semiring fill).  mats (optional): 11 pointers in the member order
// of align_pair_work_t (align_pair.hpp:47-57): mch del ins mch_mch mch_del mch_ins del_mch
// del_del ins_mch ins_del ins_ins.
int coati_ref_forward(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const float* table,
                      float g, float e, size_t k, float** mats) {
    try {
        coati::alignment_t aln;
        load_model(aln, table, g, e, k);
        coati::align_pair_work_t work;
        coati::forward(work, seq_view_t(a, la), seq_view_t(b, lb), aln);
        if(mats) {
            const coati::Matrixf* src[11] = {&work.mch,     &work.del,     &work.ins,     &work.mch_mch,
                                             &work.mch_del, &work.mch_ins, &work.del_mch, &work.del_del,
                                             &work.ins_mch, &work.ins_del, &work.ins_ins};
            for(int m = 0; m < 11; ++m) copy_matrix(*src[m], mats[m]);
        }
    } catch(const std::bad_alloc&) {
        return -2;
    } catch(const std::exception&) {
        return -1;
    }
    return 0;
}

// forward + n x sampleback with a shared RNG stream (align_marginal.cc:585-593).
// out_a/out_b: n rows of stride (la+lb+1) bytes; state = raw Lehmer state {lo, hi}, in-out.
int coati_ref_sample(const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const char* anc,
                     const char* des, const float* table, float g, float e, size_t k,
                     uint64_t state[2], size_t n, char* out_a, char* out_b, size_t* out_len,
                     float* scores, double* seconds_fill, double* seconds_sample) {
    try {
        coati::alignment_t aln;
        load_model(aln, table, g, e, k);
        auto rng = make_rng(state);
        coati::align_pair_work_t work;
        auto t0 = std::chrono::steady_clock::now();
        coati::forward(work, seq_view_t(a, la), seq_view_t(b, lb), aln);
        auto t1 = std::chrono::steady_clock::now();
        std::string sa(anc, la), sb(des, lb);
        size_t stride = la + lb + 1;
        for(size_t s = 0; s < n; ++s) {
            coati::sampleback(work, sa, sb, aln, k, rng);
            copy_out(aln, out_a ? out_a + s * stride : nullptr, out_b ? out_b + s * stride : nullptr,
                     out_len ? out_len + s : nullptr, scores ? scores + s : nullptr);
        }
        auto t2 = std::chrono::steady_clock::now();
        if(seconds_fill) *seconds_fill = std::chrono::duration<double>(t1 - t0).count();
        if(seconds_sample) *seconds_sample = std::chrono::duration<double>(t2 - t1).count();
        save_rng(rng, state);
    } catch(const std::bad_alloc&) {
        return -2;
    } catch(const std::exception&) {
        return -1;
    }
    return 0;
}

// string_seed_seq + Random::Seed(SeedSeq) (random.hpp:408-413, 523-540; coati-sample.cc).
void coati_ref_seed(const char* const* seeds, size_t n, uint64_t state[2]) {
    std::vector<std::string> v(seeds, seeds + n);
    auto ss = fragmites::random::string_seed_seq(v.begin(), v.end());
    fragmites::random::Random r;
    r.Seed(ss);
    save_rng(r, state);
}

uint64_t coati_ref_rng_bits(uint64_t state[2]) {
    auto r = make_rng(state);
    uint64_t u = r.bits();
    save_rng(r, state);
    return u;
}

float coati_ref_rng_f24(uint64_t state[2]) {
    auto r = make_rng(state);
    float f = r.f24();
    save_rng(r, state);
    return f;
}

// Threaded batch driver used ONLY as the CPU baseline in bench.py: `threads` workers pull
// pairs from a shared counter; every pair runs the reference viterbi_mem + traceback_viterbi.
// CSR layout: pair p has a = a_all[a_off[p] .. a_off[p+1]) etc.; anc/des share the offsets.
// Outputs: out_a/out_b arenas at offset a_off[p] + b_off[p] + p (capacity la+lb+1 per pair).
// Returns wall seconds, or a negative number on failure.
double coati_ref_viterbi_batch(size_t npairs, const uint8_t* a_all, const uint64_t* a_off,
                               const uint8_t* b_all, const uint64_t* b_off, const char* anc_all,
                               const char* des_all, const float* table, float g, float e, size_t k,
                               int threads, char* out_a, char* out_b, uint64_t* out_len,
                               float* scores) {
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto worker = [&]() {
        coati::alignment_t aln;
        load_model(aln, table, g, e, k);
        for(;;) {
            size_t p = next.fetch_add(1);
            if(p >= npairs) break;
            try {
                size_t la = a_off[p + 1] - a_off[p], lb = b_off[p + 1] - b_off[p];
                coati::align_pair_work_mem_t work;
                coati::viterbi_mem(work, seq_view_t(a_all + a_off[p], la),
                                   seq_view_t(b_all + b_off[p], lb), aln);
                coati::traceback_viterbi(work, std::string(anc_all + a_off[p], la),
                                         std::string(des_all + b_off[p], lb), aln, k);
                size_t o = a_off[p] + b_off[p] + p;
                size_t len = 0;
                copy_out(aln, out_a ? out_a + o : nullptr, out_b ? out_b + o : nullptr, &len,
                         scores ? scores + p : nullptr);
                if(out_len) out_len[p] = len;
            } catch(...) {
                failed = 1;
            }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for(int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for(auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if(failed) return -1.0;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
