// TEST INFRASTRUCTURE ONLY.
//
// Shows that veryfasttree::B200Operations<P> (veryfasttree_b200/csrc/B200Operations.h) satisfies the
// reference's compile-time plug-in surface: the reference's OWN NeighbourJoining<P, Operations>
// template (included from /root/reference/src, nothing copied) is instantiated once with
// BasicOperations and once with B200Operations, both run the NJ phase on the same alignment, and the
// two trees (topology + branch lengths, printNJInternal) must be identical -- the ten per-element
// primitives of the new backend are arithmetically the reference's `-ext NONE`.
//
// usage: plugin_probe <fasta> <nt|aa>     exit code 0 iff identical for float and double
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

#include "operations/BasicOperations.h"
#include "../veryfasttree_b200/csrc/B200Operations.h"
#include "NeighbourJoining.h"

using namespace veryfasttree;

template<typename P, template<class> class Ops>
static std::string njTree(const std::string &fasta, bool aa) {
    Options options;
    options.verbose = 0; options.showProgress = false; options.threads = 1; options.diskComputing = false;
    options.nCodes = aa ? 20 : 4; options.useMatrix = aa;
    options.codesString = aa ? Constants::codesStringAA : Constants::codesStringNT;
    options.doublePrecision = sizeof(P) == 8;
    options.fPostTotalTolerance = sizeof(P) == 8 ? Constants::fPostTotalToleranceDouble : Constants::fPostTotalToleranceFloat;
    omp_set_num_threads(1);
    std::ifstream in(fasta);
    std::ostringstream lg;
    Alignment aln(options, in, lg);
    aln.readAlignment();
    std::vector<std::string> seqs = aln.seqs;
    typedef Ops<P> op_t;
    DistanceMatrix<P, op_t::ALIGNMENT> dmat{};
    TransitionMatrix<P, op_t::ALIGNMENT> transmat;
    if (aa) { dmat.matrixBLOSUM45(); dmat.setupDistanceMatrix(options, lg); }
    ProgressReport progress(false, 0, false);
    std::vector<std::string> cons;
    std::unique_ptr<DiskMemory> d1, d2;
    NeighbourJoining<P, Ops> nj(options, lg, progress, seqs, aln.nPos, cons, dmat, transmat, d1, d2);
    nj.fastNJ();
    std::ostringstream out;
    nj.printNJInternal(out, true);
    return out.str();
}

int main(int argc, char **argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: plugin_probe <fasta> <nt|aa>\n"); return 2; }
    bool aa = std::string(argv[2]) == "aa";
    int bad = 0;
    bad += njTree<double, BasicOperations>(argv[1], aa) != njTree<double, B200Operations>(argv[1], aa);
    bad += njTree<float, BasicOperations>(argv[1], aa) != njTree<float, B200Operations>(argv[1], aa);
    std::printf("plugin_probe: %s\n", bad ? "DIFFERENT" : "IDENTICAL");
    return bad;
}
