// Host-side C++17 layer above the C ABI: COATi's library surface for the marginal path
// (`coati::alignment_t` in; `marg_alignment` / `marg_sample` / `alignment_score` semantics and
// FASTA / PHYLIP / JSON out).  The dynamic programs run ONLY through
// include/coati_gpu.h -- there is no CPU implementation of the hot path in here.
//
// Mirrors (paths relative to the reference):
//   src/include/coati/structs.hpp:37-132, data.hpp:44-77      gap_t, alignment_t, data_t, sample_t, args_t
//   src/lib/utils.cc:72-85, 496-528, 595-618, 738-749, 789-838, 945-967, 1044-1063, 1144-1211
//   src/lib/mutation_coati.cc:49-125, 164-354   mg94_p, marginal_p, ambiguous_*_p, gtr_q
//   src/lib/mutation_ecm.cc:151-184             ecm_p
//   src/lib/align_marginal.cc:44-88, 373-473, 536-594   marg_alignment, alignment_score, marg_sample
//   src/lib/fasta.cc:39-76,182-191  phylip.cc:194-217  json.cc:37-42,163-227  io.cc:184-222,316-346
//   contrib/random/random.hpp:80-136, 334-413, 465-472, 523-540   Lehmer64Fast, SeedSeq, string_seed_seq
#pragma once

#include <array>
#include <cstdint>
#include <iosfwd>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

struct coati_gpu_ctx;

namespace coati {

using float_t = float;

struct gap_t {
    std::size_t len{1};
    float_t open{0.001f};
    float_t extend{1.0f - 1.0f / 6.0f};
};

enum struct AmbiguousNucs { SUM, BEST };
enum struct MarginalSubst { SUM, MAX };

struct data_t {
    std::string path;
    std::vector<std::string> names;
    std::vector<std::string> seqs;
    float_t score{0.f};
    std::vector<std::string> stops;
    std::size_t size() const {
        if(names.size() != seqs.size()) throw std::invalid_argument("Different number of sequences and names.");
        return names.size();
    }
};

// 183 x 15 row-major log-odds table (the reference's `Matrixf subst_matrix`)
struct subst_table_t {
    std::vector<float_t> v;
    float_t operator()(std::size_t row, std::size_t col) const { return v[row * 15 + col]; }
    bool empty() const { return v.empty(); }
};

struct alignment_t {
    data_t data;
    std::string model{"mar-mg"};
    float_t br_len{0.0133f};
    float_t omega{0.2f};
    std::vector<float_t> pi{0.308f, 0.185f, 0.199f, 0.308f};
    std::string refs;
    bool rev{false};
    std::string rate;  // --sub: path to a CSV codon rate matrix (io.cc:48-88)
    gap_t gap;
    std::vector<float_t> sigma{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // false (default): sigma is ignored by the marginal models exactly as upstream (utils.cc:606); true: the GTR
    // rates reach mg94_p (mutation_coati.cc:317-354) -- a deliberate, documented deviation
    bool use_sigma{false};
    subst_table_t subst_matrix;
    std::string output;
    bool score{false};
    AmbiguousNucs amb = AmbiguousNucs::SUM;
    MarginalSubst sub = MarginalSubst::SUM;
    bool is_marginal() const { return model == "mar-mg" || model == "mar-ecm" || !rate.empty(); }
    std::string& seq(std::size_t i) { return data.seqs[i]; }
};

struct sample_t {
    std::size_t sample_size{1};
    std::vector<std::string> seeds{{""}};
};

struct args_t {
    alignment_t aln;
    sample_t sample;
};

// ---- codon helpers, encoding, stop codons (utils.cc) --------------------------------------------
int cod_int(std::string_view codon);
int cod64_to_61(int cod);
int cod61_to_64(int cod);
uint8_t get_nuc(uint8_t cod, int pos);
using sequence_pair_t = std::vector<std::basic_string<unsigned char>>;
sequence_pair_t marginal_seq_encoding(std::string_view anc, std::string_view des);
void order_ref(alignment_t& aln);
void process_marginal(alignment_t& aln);
void trim_end_stops(data_t& data);
void restore_end_stops(data_t& data, const gap_t& gap);

// ---- substitution models --------------------------------------------------------------------------
using matrix61_t = std::vector<float_t>;  // 61 x 61 row-major, P(i, j) = P(codon i -> codon j)
std::array<float_t, 16> gtr_q(const std::vector<float_t>& pi, const std::vector<float_t>& sigma);
matrix61_t mg94_p(float br_len, float omega, const std::vector<float_t>& nuc_freqs,
                  const std::vector<float_t>& sigma = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f});
matrix61_t ecm_p(float br_len, float omega);
subst_table_t marginal_p(const matrix61_t& P, const std::vector<float_t>& pi, AmbiguousNucs amb,
                         MarginalSubst msub);
void set_subst(alignment_t& aln);
matrix61_t parse_matrix_csv(const std::string& file);  // io.cc:48-88
// float Pade scaling-and-squaring matrix exponential (the algorithm of Eigen 3.4 MatrixBase::exp())
void expm61(const matrix61_t& A, matrix61_t& out);

// ---- RNG (contrib/random/random.hpp) --------------------------------------------------------------
struct random_t {
    uint64_t lo{0x9f57c403d06c42fcull | 1ull}, hi{0};
    void Seed(uint64_t state_lo, uint64_t state_hi) { lo = state_lo | 1ull, hi = state_hi; }
    void Seed(const std::vector<uint32_t>& seeds);           // SeedSeq<8> + Random::Seed(SeedSeq)
    void Seed(const std::vector<std::string>& seed_strings); // string_seed_seq
    uint64_t bits();
    float f24();
};
uint32_t str_crushto32(std::string_view s);

// ---- I/O ------------------------------------------------------------------------------------------
struct file_type_t {
    std::string path, type_ext;
};
file_type_t extract_file_type(std::string path);
data_t read_fasta(std::istream& in);
data_t read_phylip(std::istream& in);
data_t read_json(std::istream& in);
data_t read_input(alignment_t& aln);
void write_fasta(const data_t& d, std::ostream& out);
void write_phylip(const data_t& d, std::ostream& out);
void write_json(const data_t& d, std::ostream& out);
void write_json(const data_t& d, std::ostream& out, std::size_t iter, std::size_t sample_size);
void write_output(alignment_t& aln);
std::string json_number(float v);  // shortest round-trip of the float widened to double

// ---- drivers (GPU through the C ABI) ----------------------------------------------------------------
class gpu_context {
   public:
    explicit gpu_context(int device = 0);
    ~gpu_context();
    gpu_context(const gpu_context&) = delete;
    gpu_context& operator=(const gpu_context&) = delete;
    void set_model(const alignment_t& aln);
    coati_gpu_ctx* handle() { return h_; }

   private:
    coati_gpu_ctx* h_{nullptr};
};
[[noreturn]] void rethrow_gpu_error(int code);

// viterbi_mem + traceback_viterbi on the GPU; fills aln.data.seqs / score like traceback<S> does
void viterbi_align(gpu_context& ctx, const sequence_pair_t& enc, const std::string& anc,
                   const std::string& des, alignment_t& aln);
bool marg_alignment(alignment_t& aln, gpu_context& ctx);
void marg_sample(alignment_t& aln, std::size_t sample_size, random_t& rand, gpu_context& ctx);
float alignment_score(alignment_t& aln, const subst_table_t& p_marg);
std::string process_alignment(alignment_t& aln);

}  // namespace coati
