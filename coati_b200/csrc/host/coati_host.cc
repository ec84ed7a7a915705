// Host-side C++17 layer of coati-b200 (see coati_host.hpp): COATi's library surface for the marginal path.
// Every function cites the reference lines whose BEHAVIOUR it keeps (signatures, exception texts and output
// bytes are pinned by "drop-in"); the code is this repository's own.  The DP itself is only ever run through
// the C ABI (include/coati_gpu.h).
#include "coati_host.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>

#include "../../../include/coati_gpu.h"

namespace coati {

namespace {
#include "ecm_data.inc"

// IUPAC code of a symbol: A C G T/U R Y M K S W B D H V N '-' -> 0..15, else 16 (utils.hpp:54-61)
uint8_t nt16(unsigned char ch) {
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

// amino acid of each sense codon (utils.hpp:66-70 amino_group, as letters)
const char kAmino[62] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVYYSSSSCWCLFLF";

bool is_stop(int cod) { return cod == 48 || cod == 50 || cod == 56; }

// utils.hpp:134-156 (host libm, as the reference)
float log1p_exp(float x) {
    if(x <= -16.0f) return ::expf(x);
    if(x <= 8.0f) return ::log1pf(::expf(x));
    if(x <= 14.5f) return x + ::expf(-x);
    return x;
}
float log_sum_exp(float a, float b) {
    const float x = std::max(a, b);
    const float y = -std::fabs(a - b);
    return x + log1p_exp(y);
}
}  // namespace

// ---- utils.cc:72-85 ---------------------------------------------------------------------------------
int cod_int(std::string_view codon) {
    if(codon.size() < 3) return -1;
    int v = 0;
    for(int x = 0; x < 3; ++x) {
        const uint8_t c = nt16(static_cast<unsigned char>(codon[x]));
        if(c > 3) return -1;
        v = (v << 2) | c;
    }
    return v;
}

// ---- utils.cc:1144-1165 -----------------------------------------------------------------------------
int cod64_to_61(int cod) {
    if(cod < 0 || cod > 63)
        throw std::out_of_range("Codon index " + std::to_string(cod) + " is out of range [0-63].");
    if(is_stop(cod)) throw std::invalid_argument("Stop codon not expected in cod64_to_61");
    if(cod < 48) return cod;
    if(cod == 49) return 48;
    if(cod < 57) return cod - 2;
    return cod - 3;
}
// ---- utils.cc:1195-1211 -----------------------------------------------------------------------------
int cod61_to_64(int cod) {
    if(cod < 0 || cod > 60)
        throw std::out_of_range("Codon index " + std::to_string(cod) + " is out of range [0-60].");
    if(cod < 48) return cod;
    if(cod == 48) return 49;
    if(cod < 54) return cod + 2;
    return cod + 3;
}
// ---- utils.cc:738-749 -------------------------------------------------------------------------------
uint8_t get_nuc(uint8_t cod, int pos) {
    if(cod > 61) throw std::out_of_range("Codon out of range for list without stop codons.");
    const int c = cod61_to_64(cod);
    return static_cast<uint8_t>((c >> (4 - 2 * pos)) & 3);
}

// ---- sequence preparation ---------------------------------------------------------------------------
// One classification of an ancestor codon serves the encoder here and the error reporting of the batch entry
// points (the device kernel encode_pairs_kernel applies the same rule per codon).
namespace {
enum class codon_kind { sense, ambiguous, stop };
struct codon_code {
    codon_kind kind;
    int index61;
};
codon_code classify_codon(std::string_view codon) {
    const int c64 = cod_int(codon);
    if(c64 < 0) return {codon_kind::ambiguous, -1};
    if(is_stop(c64)) return {codon_kind::stop, -1};
    return {codon_kind::sense, cod64_to_61(c64)};
}
// a trailing stop codon of `seq` (case-insensitive, U = T), removed from it and returned; "" if there is none
std::string split_end_stop(std::string& seq) {
    if(seq.size() < 3) return {};
    const std::string_view tail(seq.data() + seq.size() - 3, 3);
    if(!is_stop(cod_int(tail))) return {};
    std::string stop(tail);
    seq.resize(seq.size() - 3);
    return stop;
}
}  // namespace

// Behaviour of utils.cc:496-528: ancestor codon -> 3 * cod61 + phase; the first codon that is not a sense codon
// decides the exception; descendant symbol -> IUPAC code (16 for anything unknown).
sequence_pair_t marginal_seq_encoding(std::string_view anc, std::string_view des) {
    sequence_pair_t enc(2);
    auto& a = enc[0];
    auto& b = enc[1];
    a.resize(anc.size() / 3 * 3 + (anc.size() % 3 ? 3 : 0));
    size_t w = 0;
    for(size_t pos = 0; pos < anc.size(); pos += 3) {
        const codon_code cc = classify_codon(anc.substr(pos, 3));
        switch(cc.kind) {
        case codon_kind::ambiguous: throw std::invalid_argument("Ambiguous nucleotides in ancestor/reference.");
        case codon_kind::stop: throw std::invalid_argument("Early stop codon in ancestor/reference.");
        case codon_kind::sense: break;
        }
        for(int phase = 0; phase < 3; ++phase) a[w++] = static_cast<unsigned char>(3 * cc.index61 + phase);
    }
    b.resize(des.size());
    std::transform(des.begin(), des.end(), b.begin(),
                   [](char ch) { return static_cast<unsigned char>(ch) < 128 ? nt16(static_cast<unsigned char>(ch)) : 16; });
    return enc;
}

// Behaviour of utils.cc:789-803: make the reference sequence the first one.
void order_ref(alignment_t& aln) {
    auto& names = aln.data.names;
    const bool first_is_ref = names[0] == aln.refs;
    const bool second_is_ref = names[1] == aln.refs;
    if(first_is_ref) return;
    if(!second_is_ref && !aln.rev) throw std::invalid_argument("Name of reference sequence not found.");
    std::swap(names[0], names[1]);
    std::swap(aln.data.seqs[0], aln.data.seqs[1]);
}

// Behaviour of utils.cc:945-967: strip one terminal stop codon per sequence and remember it.
void trim_end_stops(data_t& data) {
    const size_t n = data.size();
    for(size_t i = 0; i < n; ++i) data.stops.push_back(split_end_stop(data.seqs[i]));
}

// Behaviour of utils.cc:1044-1063: put the stops back; a stop facing no stop is aligned to a gap codon and
// charged one gap of three nucleotides.
void restore_end_stops(data_t& data, const gap_t& gap) {
    if(data.stops.size() != 2) throw std::runtime_error("Error restoring end stop codons.");
    const std::string& s0 = data.stops[0];
    const std::string& s1 = data.stops[1];
    const bool lone = s0.size() != s1.size() && (s0.empty() || s1.empty());
    if(s0.size() != s1.size() && !lone) return;  // (unreachable: a stop is 0 or 3 symbols)
    data.seqs[0] += lone && s0.empty() ? std::string("---") : s0;
    data.seqs[1] += lone && s1.empty() ? std::string("---") : s1;
    if(lone) data.score += ::logf(gap.open * gap.extend * gap.extend);
}

// Behaviour of utils.cc:809-838: the checks marg_alignment makes before aligning (lengths are tested on the
// untrimmed sequences), then the stop trimming.
void process_marginal(alignment_t& aln) {
    if(aln.data.size() != 2) throw std::invalid_argument("Exactly two sequences required.");
    const bool reorder = !aln.refs.empty() || aln.rev;
    if(reorder) order_ref(aln);
    const size_t unit = aln.gap.len;
    const size_t la = aln.seq(0).length(), lb = aln.seq(1).length();
    if(la % 3 || la % unit)
        throw std::invalid_argument("Length of reference sequence must be multiple of 3 and gap unit length.");
    if(lb % unit)
        throw std::invalid_argument("Length of descendant sequence must be multiple of gap unit length.");
    trim_end_stops(aln.data);
}

// ---- mutation_coati.cc:317-354 ----------------------------------------------------------------------
std::array<float_t, 16> gtr_q(const std::vector<float_t>& pi, const std::vector<float_t>& sigma) {
    if(std::any_of(sigma.cbegin(), sigma.cend(), [](float_t f) { return f < 0.f || f > 1.f; }))
        throw std::invalid_argument("Sigma values must be in range [0,1].");
    std::array<float_t, 16> q{};
    auto at = [&](int i, int j) -> float_t& { return q[i * 4 + j]; };
    at(0, 1) = at(1, 0) = sigma[0];
    at(0, 2) = at(2, 0) = sigma[1];
    at(0, 3) = at(3, 0) = sigma[2];
    at(1, 2) = at(2, 1) = sigma[3];
    at(1, 3) = at(3, 1) = sigma[4];
    at(2, 3) = at(3, 2) = sigma[5];
    for(int i = 0; i < 4; ++i)
        for(int j = 0; j < 4; ++j) at(i, j) *= pi[j];
    at(0, 0) = -(at(0, 1) + at(0, 2) + at(0, 3));
    at(1, 1) = -(at(1, 0) + at(1, 2) + at(1, 3));
    at(2, 2) = -(at(2, 0) + at(2, 1) + at(2, 3));
    at(3, 3) = -(at(3, 0) + at(3, 1) + at(3, 2));
    return q;
}

// ---- matrix exponential: the algorithm of Eigen 3.4 unsupported/MatrixFunctions for float --------
// (Higham 2005 scaling and squaring; Pade degree 3 / 5 / 7 chosen from the 1-norm; (V - U) X = V + U
// solved by partial-pivoting LU; result squared `squarings` times).  Eigen is not vendored by the
// reference nor present in this image, so bit-parity of P with the reference is not attainable; the
// reference's own tolerance against its golden mg94P is 1e-5 relative.
namespace {
constexpr int N61 = 61;
using mat = std::vector<float>;
void matmul(const mat& A, const mat& B, mat& C) {
    C.assign(N61 * N61, 0.f);
    for(int i = 0; i < N61; ++i)
        for(int k = 0; k < N61; ++k) {
            const float a = A[i * N61 + k];
            if(a == 0.f) continue;
            for(int j = 0; j < N61; ++j) C[i * N61 + j] += a * B[k * N61 + j];
        }
}
// X = A^-1 B by LU with partial pivoting (A, B overwritten)
void lu_solve(mat& A, mat& B) {
    for(int c = 0; c < N61; ++c) {
        int piv = c;
        for(int r = c + 1; r < N61; ++r)
            if(std::fabs(A[r * N61 + c]) > std::fabs(A[piv * N61 + c])) piv = r;
        if(piv != c)
            for(int j = 0; j < N61; ++j) {
                std::swap(A[c * N61 + j], A[piv * N61 + j]);
                std::swap(B[c * N61 + j], B[piv * N61 + j]);
            }
        const float d = A[c * N61 + c];
        for(int r = c + 1; r < N61; ++r) {
            const float f = A[r * N61 + c] / d;
            if(f == 0.f) continue;
            for(int j = c; j < N61; ++j) A[r * N61 + j] -= f * A[c * N61 + j];
            for(int j = 0; j < N61; ++j) B[r * N61 + j] -= f * B[c * N61 + j];
        }
    }
    for(int r = N61 - 1; r >= 0; --r) {
        for(int j = 0; j < N61; ++j) {
            float s = B[r * N61 + j];
            for(int k = r + 1; k < N61; ++k) s -= A[r * N61 + k] * B[k * N61 + j];
            B[r * N61 + j] = s / A[r * N61 + r];
        }
    }
}
}  // namespace

void expm61(const matrix61_t& arg, matrix61_t& out) {
    float l1 = 0.f;
    for(int j = 0; j < N61; ++j) {
        float s = 0.f;
        for(int i = 0; i < N61; ++i) s += std::fabs(arg[i * N61 + j]);
        l1 = std::max(l1, s);
    }
    mat A = arg, A2, A4, A6, tmp(N61 * N61), U, V(N61 * N61);
    int squarings = 0;
    auto poly = [&](std::initializer_list<std::pair<float, const mat*>> terms, float ident, mat& dst) {
        dst.assign(N61 * N61, 0.f);
        for(const auto& t : terms)
            for(int x = 0; x < N61 * N61; ++x) dst[x] += t.first * (*t.second)[x];
        for(int i = 0; i < N61; ++i) dst[i * N61 + i] += ident;
    };
    if(l1 < 4.258730016922831e-001f) {
        matmul(A, A, A2);
        poly({{1.f, &A2}}, 60.f, tmp);
        matmul(A, tmp, U);
        poly({{12.f, &A2}}, 120.f, V);
    } else if(l1 < 1.880152677804762e+000f) {
        matmul(A, A, A2);
        matmul(A2, A2, A4);
        poly({{1.f, &A4}, {420.f, &A2}}, 15120.f, tmp);
        matmul(A, tmp, U);
        poly({{30.f, &A4}, {3360.f, &A2}}, 30240.f, V);
    } else {
        const float maxnorm = 3.925724783138660f;
        std::frexp(l1 / maxnorm, &squarings);
        if(squarings < 0) squarings = 0;
        const float scale = std::ldexp(1.0f, -squarings);
        for(float& x : A) x *= scale;
        matmul(A, A, A2);
        matmul(A2, A2, A4);
        matmul(A4, A2, A6);
        poly({{1.f, &A6}, {1512.f, &A4}, {277200.f, &A2}}, 8648640.f, tmp);
        matmul(A, tmp, U);
        poly({{56.f, &A6}, {25200.f, &A4}, {1995840.f, &A2}}, 17297280.f, V);
    }
    mat numer(N61 * N61), denom(N61 * N61);
    for(int x = 0; x < N61 * N61; ++x) {
        numer[x] = U[x] + V[x];
        denom[x] = -U[x] + V[x];
    }
    lu_solve(denom, numer);
    for(int s = 0; s < squarings; ++s) {
        matmul(numer, numer, tmp);
        numer = tmp;
    }
    out = numer;
}

// ---- mutation_coati.cc:49-125 -----------------------------------------------------------------------
matrix61_t mg94_p(float br_len, float omega, const std::vector<float_t>& nuc_freqs,
                  const std::vector<float_t>& sigma) {
    if(br_len <= 0) throw std::out_of_range("Branch length must be positive.");
    std::array<float_t, 16> nuc_q;
    if(std::any_of(sigma.cbegin(), sigma.cend(), [](float_t f) { return f > 0.f; })) {
        nuc_q = gtr_q(nuc_freqs, sigma);
    } else {  // Yang (1994)
        nuc_q = {-0.818f, 0.132f, 0.586f, 0.1f,   0.221f, -1.349f, 0.231f, 0.897f,
                 0.909f,  0.215f, -1.322f, 0.198f, 0.1f,   0.537f,  0.128f, -0.765f};
    }
    matrix61_t Q(N61 * N61, 0.f);
    float d = 0.0f;
    for(uint8_t i = 0; i < 61; i++) {
        const float Pi = nuc_freqs[get_nuc(i, 0)] * nuc_freqs[get_nuc(i, 1)] * nuc_freqs[get_nuc(i, 2)];
        float rowSum = 0.0f;
        for(uint8_t j = 0; j < 61; j++) {
            float q = 0.f;
            if(i != j) {
                int ndiff = 0, x = 0, y = 0;
                for(int p = 2; p >= 0; --p)
                    if(get_nuc(i, p) != get_nuc(j, p)) {
                        ++ndiff;
                        x = get_nuc(i, p), y = get_nuc(j, p);  // ends on the first differing position
                    }
                if(ndiff == 1) {
                    const float w = (kAmino[i] == kAmino[j]) ? 1.f : omega;
                    q = w * nuc_q[x * 4 + y];
                }
            }
            Q[i * N61 + j] = q;
            rowSum += q;
        }
        Q[i * N61 + i] = -rowSum;
        d += Pi * rowSum;
    }
    const float scale = br_len / d;
    for(float& q : Q) q *= scale;
    matrix61_t P;
    expm61(Q, P);
    return P;
}

// ---- mutation_ecm.cc:151-184 ------------------------------------------------------------------------
matrix61_t ecm_p(float br_len, float omega) {
    if(br_len <= 0) throw std::out_of_range("Branch length must be positive.");
    auto exch = [](int i, int j) {
        if(i == j) return 0.f;
        if(i < j) std::swap(i, j);
        return kEcmExchLower[i * (i - 1) / 2 + j];
    };
    matrix61_t Q(N61 * N61, 0.f);
    float d = 0.0f;
    for(int i = 0; i < 61; i++) {
        float rowSum = 0.0f;
        for(int j = 0; j < 61; j++) {
            if(i == j) continue;
            float q = exch(i, j) * kEcmPi[j] * 1.f;  // k(i, j, 0) == 1
            if(kAmino[i] != kAmino[j]) q = q * omega;
            Q[i * N61 + j] = q;
            rowSum += q;
        }
        Q[i * N61 + i] = -rowSum;
        d += kEcmPi[i] * rowSum;
    }
    const float scale = br_len / d;
    for(float& q : Q) q *= scale;
    matrix61_t P;
    expm61(Q, P);
    return P;
}

// ---- mutation_coati.cc:164-306 ----------------------------------------------------------------------
subst_table_t marginal_p(const matrix61_t& P, const std::vector<float_t>& pi, AmbiguousNucs amb,
                         MarginalSubst msub) {
    subst_table_t t;
    t.v.assign(183 * 15, 0.f);
    for(size_t cod = 0; cod < 61; cod++)
        for(int nuc = 0; nuc < 4; nuc++)
            for(int pos = 0; pos < 3; pos++) {
                float marg = 0.f;
                for(uint8_t i = 0; i < 61; i++) {
                    const float v = (get_nuc(i, pos) == nuc ? P[cod * N61 + i] : 0.0f);
                    if(msub == MarginalSubst::SUM) marg += v;
                    else if(v > marg) marg = v;
                }
                t.v[(cod * 3 + pos) * 15 + nuc] = ::logf(marg / pi[nuc]);
            }
    static const int groups[11][4] = {{0, 2, -1, -1}, {1, 3, -1, -1}, {0, 1, -1, -1}, {2, 3, -1, -1},
                                      {1, 2, -1, -1}, {0, 3, -1, -1}, {1, 2, 3, -1},  {0, 2, 3, -1},
                                      {0, 1, 3, -1},  {0, 1, 2, -1},  {0, 1, 2, 3}};  // R Y M K S W B D H V N
    for(size_t row = 0; row < 183; ++row)
        for(int g = 0; g < 11; ++g) {
            float acc = t.v[row * 15 + groups[g][0]];
            for(int x = 1; x < 4 && groups[g][x] >= 0; ++x) {
                const float v = t.v[row * 15 + groups[g][x]];
                acc = amb == AmbiguousNucs::SUM ? log_sum_exp(acc, v) : std::max(acc, v);
            }
            t.v[row * 15 + 4 + g] = acc;
        }
    return t;
}

// ---- utils.cc:595-618 (marginal models) ---------------------------------------------------------------
// ---- io.cc:48-88 ------------------------------------------------------------------------------------
matrix61_t parse_matrix_csv(const std::string& file) {
    std::ifstream input(file);
    if(!input.good()) throw std::invalid_argument("Error opening file " + file + ".");
    std::string line;
    std::getline(input, line);
    const float br_len = std::stof(line);
    matrix61_t Q(N61 * N61, 0.f);
    int count = 0;
    while(std::getline(input, line)) {
        std::stringstream ss(line);
        std::string c0, c1, val;
        std::getline(ss, c0, ',');
        std::getline(ss, c1, ',');
        std::getline(ss, val, ',');
        const int cod0 = cod64_to_61(cod_int(c0)), cod1 = cod64_to_61(cod_int(c1));
        Q[cod0 * N61 + cod1] = std::stof(val);
        count++;
    }
    if(count != 3721) throw std::invalid_argument("Error reading substitution rate CSV file. Exiting!");
    for(float& q : Q) q *= br_len;
    matrix61_t P;
    expm61(Q, P);
    return P;
}

void set_subst(alignment_t& aln) {
    if(!aln.rate.empty()) {
        aln.model = "user_marg_model";
        aln.subst_matrix = marginal_p(parse_matrix_csv(aln.rate), aln.pi, aln.amb, aln.sub);
    } else if(aln.model == "mar-ecm") {
        // NB marginalised with the caller's pi (MG94 default), not ecm_pi -- as upstream (:603-604)
        aln.subst_matrix = marginal_p(ecm_p(aln.br_len, aln.omega), aln.pi, aln.amb, aln.sub);
    } else if(aln.model == "mar-mg") {
        // Upstream, -x/--sigma is parsed but never reaches this path: set_subst calls mg94_p without it
        // (utils.cc:606), so the GTR rates of mutation_coati.cc:317-354 are dead code from the CLI.  Default:
        // the same (drop-in).  aln.use_sigma (--gtr on the CLI) is the documented deviation SURVEY 8(f)-4 asks
        // for: the six rates go to mg94_p, which builds the nucleotide matrix with gtr_q when any is positive.
        const matrix61_t P = aln.use_sigma ? mg94_p(aln.br_len, aln.omega, aln.pi, aln.sigma)
                                           : mg94_p(aln.br_len, aln.omega, aln.pi);
        aln.subst_matrix = marginal_p(P, aln.pi, aln.amb, aln.sub);
    } else {
        throw std::invalid_argument("Mutation model unknown.");
    }
}

// ---- RNG: contrib/random/random.hpp -----------------------------------------------------------------
namespace {
void mlhash(uint64_t init, const uint32_t* in, size_t nin, uint32_t* out, size_t nout) {  // :334-358
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
        out[o] = static_cast<uint32_t>(sum >> 32);
    }
}
}  // namespace

uint32_t str_crushto32(std::string_view s) {  // :465-472 FNV-1 (char is signed on x86-64)
    uint32_t h = 2166136261U;
    for(char c : s) h = (h * 16777619U) ^ static_cast<uint32_t>(static_cast<int>(c));
    return h;
}
void random_t::Seed(const std::vector<uint32_t>& seeds) {  // :366-413
    uint32_t inner[8], outw[4];
    mlhash(0x3423da0b87484307ULL, seeds.data(), seeds.size(), inner, 8);
    mlhash(0xdf8b06c40fa44478ULL, inner, 8, outw, 4);
    Seed(static_cast<uint64_t>(outw[0]) | (static_cast<uint64_t>(outw[1]) << 32),
         static_cast<uint64_t>(outw[2]) | (static_cast<uint64_t>(outw[3]) << 32));
}
void random_t::Seed(const std::vector<std::string>& strs) {  // :523-540
    std::vector<uint32_t> u;
    for(const std::string& s : strs) {
        int32_t value = 0;
        auto [p, ec] = std::from_chars(s.data(), s.data() + s.size(), value, 10);
        if(ec == std::errc() && p == s.data() + s.size()) u.push_back(static_cast<uint32_t>(value));
        else u.push_back(str_crushto32(s));
    }
    Seed(u);
}
uint64_t random_t::bits() {  // :107,122-125
    unsigned __int128 s = (static_cast<unsigned __int128>(hi) << 64) | lo;
    s *= 0xda942042e4dd58b5ULL;
    lo = static_cast<uint64_t>(s);
    hi = static_cast<uint64_t>(s >> 64);
    return hi;
}
float random_t::f24() { return static_cast<int64_t>(bits() >> 40) / 16777216.0f; }  // :213-216

// ---- I/O ----------------------------------------------------------------------------------------------
file_type_t extract_file_type(std::string path) {  // utils.cc:632-649
    const char* ws = " \f\n\r\t\v";
    const auto b = path.find_first_not_of(ws);
    if(b == std::string::npos) return {"", ""};
    path = path.substr(b, path.find_last_not_of(ws) - b + 1);
    const auto colon = path.find_first_of(':');
    if(colon != std::string::npos && colon > 1) return {path.substr(colon + 1), "." + path.substr(0, colon)};
    // std::filesystem::path::extension semantics
    const auto slash = path.find_last_of('/');
    const std::string file = slash == std::string::npos ? path : path.substr(slash + 1);
    const auto dot = file.find_last_of('.');
    std::string ext;
    if(dot != std::string::npos && dot != 0 && file != "." && file != "..") ext = file.substr(dot);
    return {path, ext};
}

data_t read_fasta(std::istream& in) {  // fasta.cc:39-76
    data_t fasta;
    std::string line, name, content;
    while(in.good()) {
        std::getline(in, line);
        if(line.empty() || line[0] == ';') continue;
        if(line[0] == '>') {
            if(!name.empty()) {
                fasta.seqs.push_back(content);
                name.clear();
            }
            name = line.substr(1);
            if(name.empty()) throw std::invalid_argument("Input fasta file contains a sequence without a name.");
            fasta.names.push_back(name);
            content.clear();
        } else if(!name.empty()) {
            line.erase(std::remove_if(line.begin(), line.end(), [](unsigned char c) { return std::isspace(c); }),
                       line.end());
            content += line;
        }
    }
    if(!name.empty()) fasta.seqs.push_back(content);
    return fasta;
}

// PHYLIP as COATi reads and writes it: a header "<count> <columns>", then interleaved blocks; the first block
// carries the names in a 10-character field followed by 50 columns, later blocks 60 columns per row.
namespace {
constexpr size_t kPhyName = 10, kPhyFirst = 50, kPhyBlock = 60;
std::string without_space(std::string_view s) {
    std::string o;
    o.reserve(s.size());
    std::copy_if(s.begin(), s.end(), std::back_inserter(o), [](unsigned char c) { return !std::isspace(c); });
    return o;
}
}  // namespace

// Behaviour of phylip.cc:37-97.
data_t read_phylip(std::istream& in) {
    std::string tok_count, tok_width;
    in >> tok_count >> tok_width;
    const size_t count = static_cast<size_t>(std::stoi(tok_count));
    (void)std::stoi(tok_width);  // the declared width is validated as a number and otherwise ignored upstream
    data_t d;
    d.names.assign(count, {});
    d.seqs.assign(count, {});
    std::string line;
    // named rows: the rest of the header line reads as one empty line, which a row may skip once
    for(size_t row = 0; row < count; ++row) {
        std::getline(in, line);
        if(line.empty()) std::getline(in, line);
        const std::string_view v(line);
        d.names[row] = without_space(v.substr(0, kPhyName));
        d.seqs[row] = v.size() > kPhyName ? without_space(v.substr(kPhyName)) : std::string();
    }
    // continuation rows go round-robin to the sequences; blank lines separate blocks and carry nothing
    for(size_t row = 0; in.good();) {
        std::getline(in, line);
        if(line.empty()) continue;
        d.seqs[row % count] += without_space(line);
        ++row;
    }
    return d;
}

// json.cc:44-79: {"alignment": {name: seq, ...}, "score": x} -- a minimal reader for exactly the
// documents COATi writes (string values without escapes other than \" and \\)
data_t read_json(std::istream& in) {
    std::stringstream buf;
    buf << in.rdbuf();
    const std::string t = buf.str();
    data_t d;
    size_t i = t.find("\"alignment\"");
    if(i == std::string::npos) throw std::invalid_argument("Invalid JSON input: no \"alignment\" object.");
    i = t.find('{', i);
    const size_t end = t.find('}', i);
    if(i == std::string::npos || end == std::string::npos) throw std::invalid_argument("Invalid JSON input.");
    auto next_string = [&](size_t& pos, std::string& out) -> bool {
        const size_t q0 = t.find('"', pos);
        if(q0 == std::string::npos || q0 > end) return false;
        out.clear();
        size_t q = q0 + 1;
        for(; q < t.size() && t[q] != '"'; ++q) {
            if(t[q] == '\\' && q + 1 < t.size()) ++q;
            out.push_back(t[q]);
        }
        pos = q + 1;
        return true;
    };
    size_t pos = i + 1;
    std::string key, val;
    while(next_string(pos, key)) {
        if(!next_string(pos, val)) throw std::invalid_argument("Invalid JSON input.");
        d.names.push_back(key);
        d.seqs.push_back(val);
    }
    const size_t sc = t.find("\"score\"", end);
    if(sc == std::string::npos) throw std::invalid_argument("Invalid JSON input: no \"score\".");
    d.score = std::stof(t.substr(t.find(':', sc) + 1));
    return d;
}

data_t read_input(alignment_t& aln) {  // io.cc:184-222
    file_type_t t = aln.data.path.empty() ? file_type_t{"-", ".json"} : extract_file_type(aln.data.path);
    std::ifstream infile;
    std::istream* pin = &std::cin;
    if(!(t.path.empty() || t.path == "-")) {
        infile.open(t.path);
        if(!infile) throw std::invalid_argument("Opening input file " + aln.data.path + " failed.");
        pin = &infile;
    }
    data_t d;
    if(t.type_ext == ".fa" || t.type_ext == ".fasta") d = read_fasta(*pin);
    else if(t.type_ext == ".phy") d = read_phylip(*pin);
    else if(t.type_ext == ".json") d = read_json(*pin);
    else throw std::invalid_argument("Invalid input " + aln.data.path + ".");
    d.path = aln.data.path;
    return d;
}

void write_fasta(const data_t& d, std::ostream& out) {  // fasta.cc:182-191
    for(size_t i = 0; i < d.size(); i++) {
        out << ">" << d.names[i] << std::endl;
        for(size_t j = 0; j < d.seqs[i].size(); j += 60) out << d.seqs[i].substr(j, 60) << std::endl;
    }
}
// Behaviour of phylip.cc:194-217 (byte for byte: names cut or padded to 10, a blank line after every block).
void write_phylip(const data_t& d, std::ostream& out) {
    const size_t count = d.size(), columns = d.seqs[0].length();
    out << count << " " << columns << std::endl;
    for(size_t from = 0, width = kPhyFirst; from == 0 || from < columns; from += width, width = kPhyBlock) {
        for(size_t row = 0; row < count; ++row) {
            if(from == 0) {
                std::string field = d.names[row].substr(0, kPhyName);
                field.resize(kPhyName, ' ');
                out << field;
            }
            out << d.seqs[row].substr(std::min(from, d.seqs[row].size()), width) << std::endl;
        }
        out << std::endl;
    }
}
std::string json_number(float v) {  // nlohmann::json dump of a float stored as double
    const double dv = static_cast<double>(v);
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, dv);
    std::string s(buf, r.ptr);
    if(s.find_first_of(".e") == std::string::npos && s.find("inf") == std::string::npos &&
       s.find("nan") == std::string::npos)
        s += ".0";
    return s;
}
namespace {
std::string json_escape(const std::string& s) {
    std::string o;
    for(char c : s) {
        if(c == '"' || c == '\\') o.push_back('\\');
        o.push_back(c);
    }
    return o;
}
void json_object(const data_t& d, std::ostream& out) {  // json.cc:37-42 + std::setw(2) dump
    out << "{\n  \"alignment\": {\n";
    for(size_t i = 0; i < d.size(); ++i)
        out << "    \"" << json_escape(d.names[i]) << "\": \"" << d.seqs[i] << "\"" << (i + 1 < d.size() ? ",\n" : "\n");
    out << "  },\n  \"score\": " << json_number(d.score) << "\n}";
}
}  // namespace
void write_json(const data_t& d, std::ostream& out) {  // json.cc:163-168
    json_object(d, out);
    out << std::endl;
}
void write_json(const data_t& d, std::ostream& out, size_t iter, size_t sample_size) {  // json.cc:211-227
    if(iter == 0) out << "[" << std::endl;
    json_object(d, out);
    if(iter < sample_size - 1) out << "," << std::endl;
    else out << std::endl << "]" << std::endl;
}
void write_output(alignment_t& aln) {  // io.cc:316-346
    file_type_t t = aln.output.empty() ? file_type_t{"-", ".json"} : extract_file_type(aln.output);
    std::ofstream outfile;
    std::ostream* pout = &std::cout;
    if(t.path != "-") {
        outfile.open(t.path);
        pout = &outfile;
    }
    if(t.type_ext == ".fa" || t.type_ext == ".fasta") write_fasta(aln.data, *pout);
    else if(t.type_ext == ".phy") write_phylip(aln.data, *pout);
    else if(t.type_ext == ".json") write_json(aln.data, *pout);
    else throw std::invalid_argument("Invalid output format " + t.type_ext + ".");
}

// ---- GPU drivers ----------------------------------------------------------------------------------------
void rethrow_gpu_error(int code) {
    switch(code) {
    case COATI_GPU_E_NOMEM: throw std::bad_alloc();
    case COATI_GPU_E_AMBIGUOUS:
    case COATI_GPU_E_STOP:
    case COATI_GPU_E_LENGTH:
    case COATI_GPU_E_SYMBOL:
    case COATI_GPU_E_ARG: throw std::invalid_argument(coati_gpu_strerror(code));
    default: throw std::runtime_error(coati_gpu_strerror(code));
    }
}
gpu_context::gpu_context(int device) {
    if(int rc = coati_gpu_init(device, &h_)) rethrow_gpu_error(rc);
}
gpu_context::~gpu_context() { coati_gpu_shutdown(h_); }
void gpu_context::set_model(const alignment_t& aln) {
    if(aln.subst_matrix.v.size() != 183 * 15) throw std::invalid_argument("Substitution matrix not set.");
    if(int rc = coati_gpu_set_model(h_, aln.subst_matrix.v.data(), aln.gap.open, aln.gap.extend,
                                    static_cast<uint32_t>(aln.gap.len)))
        rethrow_gpu_error(rc);
}

void viterbi_align(gpu_context& ctx, const sequence_pair_t& enc, const std::string& anc,
                   const std::string& des, alignment_t& aln) {
    std::string ra(enc[0].size() + enc[1].size() + 1, '\0'), rb(ra);
    size_t n = 0;
    float score = 0.f;
    if(int rc = coati_gpu_viterbi(ctx.handle(), enc[0].data(), enc[0].size(), enc[1].data(), enc[1].size(),
                                  anc.data(), des.data(), ra.data(), rb.data(), &n, &score))
        rethrow_gpu_error(rc);
    ra.resize(n);
    rb.resize(n);
    aln.data.seqs.clear();
    aln.data.seqs.push_back(std::move(ra));
    aln.data.seqs.push_back(std::move(rb));
    aln.data.score = score;
}

// ---- align_marginal.cc:44-88 ------------------------------------------------------------------------------
bool marg_alignment(alignment_t& aln, gpu_context& ctx) {
    aln.data = read_input(aln);
    set_subst(aln);
    if(aln.score) {
        std::cout << alignment_score(aln, aln.subst_matrix) << std::endl;
        return true;
    }
    process_marginal(aln);
    const std::string anc = aln.seq(0), des = aln.seq(1);
    const sequence_pair_t seq_pair = marginal_seq_encoding(anc, des);
    try {
        ctx.set_model(aln);
        viterbi_align(ctx, seq_pair, anc, des, aln);
    } catch(const std::bad_alloc&) {
        std::cerr << "ERROR: sequences to align exceed available memory." << std::endl;
        return false;  // upstream returns EXIT_FAILURE (== true) here by mistake (:75)
    }
    restore_end_stops(aln.data, aln.gap);
    write_output(aln);
    return true;
}

// ---- utils.cc:847-935 -------------------------------------------------------------------------------------
std::string process_alignment(alignment_t& aln) {
    if(aln.data.size() != 2) throw std::invalid_argument("Exactly two sequences required.");
    if(!aln.refs.empty() || aln.rev) order_ref(aln);
    size_t len_a = aln.data.seqs[0].length(), len_b = aln.data.seqs[1].length();
    if(len_a != len_b) throw std::invalid_argument("For alignment scoring both sequences must have equal length.");
    for(size_t i = 0; i < 2; ++i) {
        std::string& seq = aln.data.seqs[i];
        long pos[3], p = static_cast<long>(seq.size()) - 1;
        int found = 0;
        for(int q = 2; q >= 0; --q) {
            while(p >= 0 && seq[p] == '-') --p;
            if(p < 0) break;
            pos[q] = p--;
            ++found;
        }
        if(found < 3) {
            aln.data.stops.emplace_back("");
            continue;
        }
        const std::string last{seq[pos[0]], seq[pos[1]], seq[pos[2]]};
        if(is_stop(cod_int(last))) {
            aln.data.stops.push_back(last);
            seq[pos[0]] = seq[pos[1]] = seq[pos[2]] = '-';
        } else {
            aln.data.stops.emplace_back("");
        }
    }
    std::string cigar;
    cigar.reserve(len_a);
    for(size_t i = 0; i < len_a; ++i) {
        const char a = aln.data.seqs[0][i], b = aln.data.seqs[1][i];
        if(a != '-' && b != '-') cigar.push_back('M');
        else if(a != '-') cigar.push_back('D');
        else if(b != '-') cigar.push_back('I');
    }
    for(std::string& s : aln.data.seqs) s.erase(std::remove(s.begin(), s.end(), '-'), s.end());
    len_a = aln.seq(0).length();
    len_b = aln.seq(1).length();
    if(len_a % 3 != 0 || len_a % aln.gap.len != 0)
        throw std::invalid_argument("Length of reference sequence must be multiple of 3 and gap unit length.");
    if(len_b % aln.gap.len != 0)
        throw std::invalid_argument("Length of descendant sequence must be multiple of gap unit length.");
    return cigar;
}

// ---- align_marginal.cc:373-473 ------------------------------------------------------------------------------
float alignment_score(alignment_t& aln, const subst_table_t& p_marg) {
    const std::string cigar = process_alignment(aln);
    const sequence_pair_t sp = marginal_seq_encoding(aln.data.seqs[0], aln.data.seqs[1]);
    const float no_gap = ::log1pf(-aln.gap.open), gap_stop = ::log1pf(-aln.gap.extend);
    const float gap_open = ::logf(aln.gap.open), gap_extend = ::logf(aln.gap.extend);
    auto power = [&](size_t n) { return gap_extend * static_cast<float>(n); };
    bool in_gap = false;
    float score = 0.f;
    size_t nins = 0, ndel = 0, apos = 0, bpos = 0;
    auto close_gap = [&](bool terminal) {
        if(nins == 0) score = (((score + no_gap) + gap_open) + power(ndel - 1)) + gap_stop;
        else if(ndel == 0) score = (((score + gap_open) + power(nins - 1)) + gap_stop) + no_gap;
        else {
            score = ((((score + gap_open) + gap_open) + power(nins + ndel - 2)) + gap_stop) + gap_stop;
            if(terminal) score = score + no_gap;
        }
    };
    for(char op : cigar) {
        if(!in_gap) {
            if(op == 'I') nins++, bpos++, in_gap = true;
            else if(op == 'D') ndel++, apos++, in_gap = true;
            else {
                score = ((score + no_gap) + no_gap) + p_marg(sp[0][apos], sp[1][bpos]);
                apos++, bpos++;
            }
        } else {
            if(op == 'I') nins++, bpos++;
            else if(op == 'D') ndel++, apos++;
            else {
                close_gap(false);
                score = score + p_marg(sp[0][apos], sp[1][bpos]);
                nins = ndel = 0;
                in_gap = false;
                apos++, bpos++;
            }
        }
    }
    if(!in_gap) score = (score + no_gap) + no_gap;
    else close_gap(true);
    aln.data.score = score;
    restore_end_stops(aln.data, aln.gap);
    return aln.data.score;
}

// ---- align_marginal.cc:536-594 ------------------------------------------------------------------------------
void marg_sample(alignment_t& aln, size_t sample_size, random_t& rand, gpu_context& ctx) {
    aln.data = read_input(aln);
    if(aln.data.size() != 2) throw std::invalid_argument("Exactly two sequences required.");
    std::ofstream outfile;
    std::ostream* pout = &std::cout;
    if(!(aln.output.empty() || aln.output == "-")) {
        outfile.open(aln.output);
        if(!outfile) throw std::invalid_argument("Opening output file " + aln.output + " failed.");
        pout = &outfile;
    }
    const size_t len_a = aln.seq(0).length();
    if(len_a % 3 != 0 || len_a % aln.gap.len != 0)
        throw std::invalid_argument("Length of reference sequence must be multiple of 3.");
    if(aln.seq(1).length() % aln.gap.len != 0)
        throw std::invalid_argument("Length of descendant sequence must be multiple of " +
                                    std::to_string(aln.gap.len) + ".");
    trim_end_stops(aln.data);
    const std::string anc = aln.seq(0), des = aln.seq(1);
    const sequence_pair_t sp = marginal_seq_encoding(anc, des);
    set_subst(aln);
    ctx.set_model(aln);

    coati_gpu_forward_t* fw = nullptr;
    if(int rc = coati_gpu_forward(ctx.handle(), sp[0].data(), sp[0].size(), sp[1].data(), sp[1].size(), &fw))
        rethrow_gpu_error(rc);
    const size_t stride = sp[0].size() + sp[1].size() + 1;
    std::vector<char> ra(sample_size * stride + 1), rb(sample_size * stride + 1);
    std::vector<size_t> len(sample_size);
    std::vector<float> score(sample_size);
    uint64_t st[2] = {rand.lo, rand.hi};
    const int rc = coati_gpu_sampleback(fw, anc.data(), des.data(), st, sample_size, ra.data(), rb.data(),
                                        len.data(), score.data(), nullptr);
    coati_gpu_forward_free(fw);
    if(rc) rethrow_gpu_error(rc);
    rand.Seed(st[0], st[1]);  // the caller's stream advances exactly as upstream's would
    const std::vector<std::string> stops = aln.data.stops;
    for(size_t i = 0; i < sample_size; ++i) {
        aln.data.seqs = {std::string(&ra[i * stride], len[i]), std::string(&rb[i * stride], len[i])};
        aln.data.score = score[i];
        aln.data.stops = stops;
        restore_end_stops(aln.data, aln.gap);
        write_json(aln.data, *pout, i, sample_size);
    }
}

}  // namespace coati

// ---- C entry points for the ctypes tests (host logic only; no DP here) -------------------------------------
extern "C" {

// 183 x 15 table for a marginal model.  model: 0 = mar-mg, 1 = mar-ecm; amb: 0 SUM 1 BEST; msub: 0 SUM 1 MAX
int coati_host_marginal_table(int model, float br_len, float omega, const float* pi, int amb, int msub,
                              float* out) {
    try {
        coati::alignment_t aln;
        aln.model = model == 0 ? "mar-mg" : "mar-ecm";
        aln.br_len = br_len;
        aln.omega = omega;
        aln.pi.assign(pi, pi + 4);
        aln.amb = amb ? coati::AmbiguousNucs::BEST : coati::AmbiguousNucs::SUM;
        aln.sub = msub ? coati::MarginalSubst::MAX : coati::MarginalSubst::SUM;
        coati::set_subst(aln);
        std::memcpy(out, aln.subst_matrix.v.data(), 183 * 15 * sizeof(float));
    } catch(...) {
        return -1;
    }
    return 0;
}

// the same with the GTR rates wired through (alignment_t::use_sigma), mar-mg only
int coati_host_marginal_table_gtr(float br_len, float omega, const float* pi, const float* sigma, float* out) {
    try {
        coati::alignment_t aln;
        aln.br_len = br_len;
        aln.omega = omega;
        aln.pi.assign(pi, pi + 4);
        aln.sigma.assign(sigma, sigma + 6);
        aln.use_sigma = true;
        coati::set_subst(aln);
        std::memcpy(out, aln.subst_matrix.v.data(), 183 * 15 * sizeof(float));
    } catch(...) {
        return -1;
    }
    return 0;
}

// write_phylip of two rows, read back with read_phylip: returns 0 when names (cut to 10) and rows survive
int coati_host_phylip_roundtrip(const char* n0, const char* s0, const char* n1, const char* s1, char* text,
                                size_t cap) {
    try {
        coati::data_t d;
        d.names = {n0, n1};
        d.seqs = {s0, s1};
        std::ostringstream os;
        coati::write_phylip(d, os);
        const std::string t = os.str();
        if(t.size() + 1 > cap) return -2;
        std::memcpy(text, t.c_str(), t.size() + 1);
        std::istringstream is(t);
        const coati::data_t r = coati::read_phylip(is);
        if(r.seqs != d.seqs) return -3;
        if(r.names[0] != d.names[0].substr(0, 10) || r.names[1] != d.names[1].substr(0, 10)) return -4;
    } catch(...) {
        return -1;
    }
    return 0;
}

int coati_host_mg94_p(float br_len, float omega, const float* pi, const float* sigma, float* out) {
    try {
        std::vector<float> s(6, 0.f);
        if(sigma) s.assign(sigma, sigma + 6);
        const auto P = coati::mg94_p(br_len, omega, std::vector<float>(pi, pi + 4), s);
        std::memcpy(out, P.data(), 61 * 61 * sizeof(float));
    } catch(...) {
        return -1;
    }
    return 0;
}

int coati_host_gtr_q(const float* pi, const float* sigma, float* out) {
    try {
        const auto q = coati::gtr_q(std::vector<float>(pi, pi + 4), std::vector<float>(sigma, sigma + 6));
        std::memcpy(out, q.data(), 16 * sizeof(float));
    } catch(...) {
        return -1;
    }
    return 0;
}

// 0 ok, -6 ambiguous, -7 early stop (codes of include/coati_gpu.h)
int coati_host_encode(const char* anc, size_t la, const char* des, size_t lb, uint8_t* a, uint8_t* b) {
    try {
        const auto sp = coati::marginal_seq_encoding(std::string_view(anc, la), std::string_view(des, lb));
        std::memcpy(a, sp[0].data(), sp[0].size());
        std::memcpy(b, sp[1].data(), sp[1].size());
    } catch(const std::invalid_argument& e) {
        return std::strstr(e.what(), "Ambiguous") ? COATI_GPU_E_AMBIGUOUS : COATI_GPU_E_STOP;
    }
    return 0;
}

void coati_host_seed(const char* const* seeds, size_t n, uint64_t state[2]) {
    coati::random_t r;
    r.Seed(std::vector<std::string>(seeds, seeds + n));
    state[0] = r.lo;
    state[1] = r.hi;
}

int coati_host_alignment_score(const char* a, const char* b, const float* table, float g, float e, size_t k,
                               float* score) {
    try {
        coati::alignment_t aln;
        aln.data.names = {"A", "B"};
        aln.data.seqs = {a, b};
        aln.gap.open = g, aln.gap.extend = e, aln.gap.len = k;
        coati::subst_table_t t;
        t.v.assign(table, table + 183 * 15);
        *score = coati::alignment_score(aln, t);
    } catch(...) {
        return -1;
    }
    return 0;
}

// io.cc:48-88: P = expm(Q * t)^T from a "cod,cod,rate" CSV; out = 61 x 61 row-major
int coati_host_parse_matrix_csv(const char* path, float* out) {
    try {
        const coati::matrix61_t P = coati::parse_matrix_csv(path);
        std::copy(P.begin(), P.end(), out);
    } catch(const std::invalid_argument&) {
        return -2;
    } catch(...) {
        return -1;
    }
    return 0;
}

// read_input + write_output (io.cc:184-222, 316-346): format chosen by the extensions, as the CLI does
int coati_host_convert(const char* in_path, const char* out_path) {
    try {
        coati::alignment_t aln;
        aln.data.path = in_path;
        aln.data = coati::read_input(aln);
        aln.output = out_path;
        coati::write_output(aln);
    } catch(const std::invalid_argument&) {
        return -2;
    } catch(...) {
        return -1;
    }
    return 0;
}

void coati_host_json_number(float v, char* out, size_t cap) {
    const std::string s = coati::json_number(v);
    std::strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
}

}  // extern "C"
