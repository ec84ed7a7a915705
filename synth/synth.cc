// Seeded synthetic codon-sequence pairs for the benchmark workloads.
// Host-side, multi-threaded, deterministic per pair: pair p draws from splitmix64(seed, p) only,
// so any subset or sharding of the batch reproduces the same sequences.
//
//   ancestor   : n codons, uniform over the 61 sense codons (never a stop codon)
//   descendant : ancestor with per-nucleotide substitution prob `sub`, and codon-unit indels
//                (length 3 * Geom(mean 2) nt) opened with prob `indel` per codon for each of
//                insertion and deletion  =>  Lb % 3 == 0, symbols in ACGT only
//   lengths    : C5 = bins {150,300,600,1200,2400} nt with weights {40,30,20,8,2} %
//                C4 = n ~ U[100,1000] codons (300-3000 nt)
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

SplitMix pair_rng(uint64_t seed, uint64_t p) {
    SplitMix r{seed * 0x9e3779b97f4a7c15ull + p * 0xd1342543de82ef95ull + 0x632be59bd9b4e019ull};
    r.next();
    return r;
}

const char NUC[4] = {'A', 'C', 'G', 'T'};

// 61-codon index -> 64-codon index (skipping TAA=48, TAG=50, TGA=56)
inline int cod61_to_64(int c) { return c < 48 ? c : c == 48 ? 49 : c < 54 ? c + 2 : c + 3; }

uint32_t draw_codons(SplitMix& r, int workload) {
    if(workload == 5) {
        const double u = r.unit();
        return u < 0.40 ? 50 : u < 0.70 ? 100 : u < 0.90 ? 200 : u < 0.98 ? 400 : 800;
    }
    return 100 + r.below(901);  // C4: U[100, 1000]
}

uint32_t geom_mean2(SplitMix& r) {  // support 1,2,..., mean 2
    uint32_t n = 1;
    while(r.unit() < 0.5 && n < 64) ++n;
    return n;
}

// One pair.  If anc == nullptr only the lengths are produced (first pass).
void make_pair(uint64_t seed, uint64_t p, int workload, double sub, double indel, uint32_t* la_out,
               uint32_t* lb_out, char* anc, char* des, uint8_t* a, uint8_t* b) {
    SplitMix r = pair_rng(seed, p);
    const uint32_t ncod = draw_codons(r, workload);
    uint32_t ia = 0, ib = 0;
    uint32_t del_left = 0;
    for(uint32_t x = 0; x < ncod; ++x) {
        const int c61 = (int)r.below(61), c64 = cod61_to_64(c61);
        const int nuc[3] = {(c64 >> 4) & 3, (c64 >> 2) & 3, c64 & 3};
        if(anc) {
            for(int q = 0; q < 3; ++q) {
                anc[ia + q] = NUC[nuc[q]];
                a[ia + q] = (uint8_t)(3 * c61 + q);
            }
        }
        ia += 3;
        // insertion before this codon
        if(r.unit() < indel) {
            const uint32_t len = 3 * geom_mean2(r);
            for(uint32_t q = 0; q < len; ++q) {
                const int n = (int)r.below(4);
                if(des) des[ib] = NUC[n], b[ib] = (uint8_t)n;
                ++ib;
            }
        }
        if(del_left == 0 && r.unit() < indel) del_left = geom_mean2(r);
        if(del_left > 0) {
            --del_left;
            continue;
        }
        for(int q = 0; q < 3; ++q) {
            int n = nuc[q];
            if(r.unit() < sub) n = (int)r.below(4);
            if(des) des[ib] = NUC[n], b[ib] = (uint8_t)n;
            ++ib;
        }
    }
    if(ib == 0) {  // never emit an empty descendant
        for(int q = 0; q < 3; ++q) {
            if(des) des[ib] = 'A', b[ib] = 0;
            ++ib;
        }
    }
    *la_out = ia;
    *lb_out = ib;
}

template <class F>
void parallel_for(uint64_t n, int threads, F f) {
    threads = std::max(1, threads);
    std::vector<std::thread> pool;
    const uint64_t chunk = (n + threads - 1) / threads;
    for(int t = 0; t < threads; ++t) {
        const uint64_t lo = t * chunk, hi = std::min(n, lo + chunk);
        if(lo >= hi) break;
        pool.emplace_back([=] { for(uint64_t p = lo; p < hi; ++p) f(p); });
    }
    for(auto& th : pool) th.join();
}

}  // namespace

extern "C" {

// Pass 1: lengths of pairs [first, first + n) -> CSR offsets a_off/b_off (n + 1 entries each).
void coati_synth_offsets(uint64_t seed, uint64_t first, uint64_t n, int workload, double sub,
                         double indel, int threads, uint64_t* a_off, uint64_t* b_off) {
    std::vector<uint32_t> la(n), lb(n);
    parallel_for(n, threads, [&](uint64_t p) {
        make_pair(seed, first + p, workload, sub, indel, &la[p], &lb[p], nullptr, nullptr, nullptr,
                  nullptr);
    });
    a_off[0] = b_off[0] = 0;
    for(uint64_t p = 0; p < n; ++p) {
        a_off[p + 1] = a_off[p] + la[p];
        b_off[p + 1] = b_off[p] + lb[p];
    }
}

// Pass 2: raw symbols (anc/des) and their encodings (a: codon61*3+phase, b: 0..3) into arenas.
void coati_synth_fill(uint64_t seed, uint64_t first, uint64_t n, int workload, double sub,
                      double indel, int threads, const uint64_t* a_off, const uint64_t* b_off,
                      char* anc_all, char* des_all, uint8_t* a_all, uint8_t* b_all) {
    parallel_for(n, threads, [&](uint64_t p) {
        uint32_t la, lb;
        make_pair(seed, first + p, workload, sub, indel, &la, &lb, anc_all + a_off[p],
                  des_all + b_off[p], a_all + a_off[p], b_all + b_off[p]);
    });
}

}  // extern "C"
