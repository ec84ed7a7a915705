// coati-gpu: `alignpair` and `sample` verbs of COATi for the marginal models, running the dynamic
// programs on the GPU through libcoati_gpu.so.  Option names follow the reference's tables
// (src/lib/utils.cc:93-161 alignpair, :328-380 sample); only marginal models are accepted.
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "coati_host.hpp"

namespace {
struct parser {
    std::vector<std::string> a;
    size_t i = 0;
    bool more() const { return i < a.size(); }
    bool is_flag(const std::string& s, const char* sh, const char* lg) const { return s == sh || s == lg; }
    std::string value(const std::string& opt) {
        if(i >= a.size()) throw std::invalid_argument(opt + ": 1 required");
        return a[i++];
    }
};

int usage() {
    std::cerr << "usage: coati-gpu alignpair|sample input.fasta [-m mar-mg|mar-ecm] [-t time] [-g gap-open]\n"
                 "       [-e gap-extend] [-w omega] [-p A C G T] [-k gap-len] [-o output] [-r ref | -v] [-s]\n"
                 "       [-a SUM|BEST] [--marginal-sub SUM|MAX] [-d device]   (sample: -n size -s seeds...)\n";
    return 1;
}
}  // namespace

int main(int argc, char* argv[]) {
    if(argc < 2) return usage();
    const std::string verb = argv[1];
    if(verb != "alignpair" && verb != "sample") return usage();
    const bool sampling = verb == "sample";
    coati::args_t args;
    int device = 0;
    bool seeds_given = false;
    parser p;
    for(int x = 2; x < argc; ++x) p.a.emplace_back(argv[x]);
    try {
        while(p.more()) {
            const std::string o = p.a[p.i++];
            if(p.is_flag(o, "-m", "--model")) args.aln.model = p.value(o);
            else if(o == "--sub") args.aln.rate = p.value(o);
            else if(p.is_flag(o, "-t", "--time")) args.aln.br_len = std::stof(p.value(o));
            else if(p.is_flag(o, "-r", "--ref")) args.aln.refs = p.value(o);
            else if(p.is_flag(o, "-v", "--rev-ref")) args.aln.rev = true;
            else if(p.is_flag(o, "-o", "--output")) args.aln.output = p.value(o);
            else if(p.is_flag(o, "-g", "--gap-open")) args.aln.gap.open = std::stof(p.value(o));
            else if(p.is_flag(o, "-e", "--gap-extend")) args.aln.gap.extend = std::stof(p.value(o));
            else if(p.is_flag(o, "-w", "--omega")) args.aln.omega = std::stof(p.value(o));
            else if(p.is_flag(o, "-k", "--gap-len")) args.aln.gap.len = std::stoul(p.value(o));
            else if(p.is_flag(o, "-d", "--device")) device = std::stoi(p.value(o));
            else if(p.is_flag(o, "-p", "--pi")) {
                for(int q = 0; q < 4; ++q) args.aln.pi[q] = std::stof(p.value(o));
            } else if(p.is_flag(o, "-x", "--sigma")) {
                for(int q = 0; q < 6; ++q) args.aln.sigma[q] = std::stof(p.value(o));
            } else if(o == "--gtr") {
                // not an upstream option: let -x/--sigma reach the marginal table (alignment_t::use_sigma)
                args.aln.use_sigma = true;
            } else if(p.is_flag(o, "-b", "--base-error")) {
                p.value(o);  // parsed and unused on the marginal path, as upstream
            } else if(p.is_flag(o, "-a", "--ambiguous")) {
                std::string v = p.value(o);
                for(char& c : v) c = static_cast<char>(std::toupper(c));
                if(v == "SUM") args.aln.amb = coati::AmbiguousNucs::SUM;
                else if(v == "BEST") args.aln.amb = coati::AmbiguousNucs::BEST;
                else throw std::invalid_argument("--ambiguous: SUM or BEST");
            } else if(o == "--marginal-sub") {
                std::string v = p.value(o);
                for(char& c : v) c = static_cast<char>(std::toupper(c));
                if(v == "SUM") args.aln.sub = coati::MarginalSubst::SUM;
                else if(v == "MAX") args.aln.sub = coati::MarginalSubst::MAX;
                else throw std::invalid_argument("--marginal-sub: SUM or MAX");
            } else if(sampling && p.is_flag(o, "-n", "--sample-size")) {
                args.sample.sample_size = std::stoul(p.value(o));
            } else if(sampling && p.is_flag(o, "-s", "--seed")) {
                if(!seeds_given) args.sample.seeds.clear();
                seeds_given = true;
                while(p.more() && (p.a[p.i].empty() || p.a[p.i][0] != '-')) args.sample.seeds.push_back(p.a[p.i++]);
            } else if(!sampling && p.is_flag(o, "-s", "--score")) {
                args.aln.score = true;
            } else if(!o.empty() && o[0] == '-' && o != "-") {
                throw std::invalid_argument("The following argument was not expected: " + o);
            } else {
                args.aln.data.path = o;
            }
        }
        if(args.aln.data.path.empty()) throw std::invalid_argument("input is required");
        if(!args.aln.refs.empty() && args.aln.rev) throw std::invalid_argument("--rev-ref excludes --ref");
        if(args.aln.br_len <= 0 || args.aln.gap.open <= 0 || args.aln.gap.extend <= 0 || args.aln.omega <= 0)
            throw std::invalid_argument("Number less or equal to 0");
    } catch(const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return usage() + 105;
    }
    if(!args.aln.is_marginal()) {
        std::cerr << "ERROR: only the marginal models mar-mg and mar-ecm run on the GPU." << std::endl;
        return EXIT_FAILURE;
    }
    try {
        coati::gpu_context ctx(device);
        if(sampling) {
            coati::random_t rand;
            rand.Seed(args.sample.seeds);
            coati::marg_sample(args.aln, args.sample.sample_size, rand, ctx);
        } else if(!coati::marg_alignment(args.aln, ctx)) {
            return EXIT_FAILURE;
        }
    } catch(const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;  // coati-alignpair.cc:39-49
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
