// B200FastNJ.h -- the one call a maintainer adds to the reference: the NJ + TopHits phase through the B200 library.
//
// VeryFastTreeImpl.tcc:140 calls `nj.fastNJ()`.  With `-ext B200` (integration/reference_b200.patch) that line becomes
// `b200FastNJ(nj)`: the leaf codes the reference's own seqsToProfiles produced (NeighbourJoining.tcc:382-534) and the tables
// its own setupDistanceMatrix produced (DistanceMatrix.tcc:102-153) go to vft_nj_build (include/vft_b200.h), the tree comes
// back into the reference's own arrays (parent / child / branchlength / root / maxnode, NeighbourJoining.h:294-299), and the
// reference's own recomputeProfiles (NeighbourJoining.tcc:3482-3506) rebuilds the internal profiles the later phases (NNI,
// SPR, ML) read.  Everything after that line is the unmodified reference.
//
// Compiled inside the reference's translation units (it needs the NeighbourJoining template) with -fno-access-control, as the
// members it fills are private; a maintainer would make it a member function instead.
#ifndef VERYFASTTREE_B200FASTNJ_H
#define VERYFASTTREE_B200FASTNJ_H

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vft_b200.h"

namespace veryfasttree {

template<typename Precision, template<class> class Operations>
void b200FastNJ(NeighbourJoining<Precision, Operations> &nj) {
    const int64_t N = nj.nSeqs, L = nj.nPos, M = 2 * N;
    const Options &o = nj.options;
    if (!o.constraintsFile.empty() || o.slow) throw std::invalid_argument("B200 backend: -constraints and -slow are not supported");
    std::vector<uint8_t> codes((size_t) (N * L));
    for (int64_t i = 0; i < N; i++)
        for (int64_t p = 0; p < L; p++) codes[(size_t) (i * L + p)] = (uint8_t) nj.profiles[i].codes[p];
    vft_config cfg{};
    cfg.nSeqs = N; cfg.nPos = L; cfg.nCodes = (int32_t) o.nCodes; cfg.precision = (int32_t) (8 * sizeof(Precision));
    cfg.useMatrix = o.useMatrix ? 1 : 0; cfg.reduction = VFT_REDUCE_SCALAR;   /* the lane order of B200Operations' own primitives (== -ext NONE), which the later phases use */ cfg.device = 0; cfg.fPostTotalTolerance = o.fPostTotalTolerance;
    vft_nj_options opt;
    vft_nj_default_options(&opt);
    opt.tophitsMult = o.tophitsMult; opt.tophitsClose = o.tophitsClose; opt.topvisibleMult = o.topvisibleMult; opt.tophitsRefresh = o.tophitsRefresh;
    opt.staleOutLimit = o.staleOutLimit; opt.fResetOutProfile = o.fResetOutProfile; opt.nResetOutProfile = (int32_t) o.nResetOutProfile;
    opt.bionj = o.bionj ? 1 : 0;
    std::vector<Precision> d(400), ev(20), et(20), cf(400);
    const void *tables[4] = {d.data(), ev.data(), et.data(), cf.data()};
    if (o.useMatrix) {
        const auto &dm = nj.distanceMatrix;
        for (int i = 0; i < 20; i++) {
            ev[(size_t) i] = dm.eigenval[i]; et[(size_t) i] = dm.eigentot[i];
            for (int j = 0; j < 20; j++) { d[(size_t) (i * 20 + j)] = dm.distances[i][j]; cf[(size_t) (i * 20 + j)] = dm.codeFreq[i][j]; }
        }
    }
    std::vector<int64_t> parent((size_t) M), child((size_t) (3 * M));
    std::vector<int32_t> nChild((size_t) M);
    std::vector<Precision> bl((size_t) M);
    vft_nj_result res{};
    res.parent = parent.data(); res.nChild = nChild.data(); res.child = child.data(); res.branchlength = bl.data();
    const int rc = vft_nj_build(&cfg, &opt, codes.data(), o.useMatrix ? tables : nullptr, &res);
    if (rc != VFT_OK) throw std::invalid_argument(std::string("B200 backend: ") + vft_last_error());
    for (int64_t i = 0; i < res.maxnode; i++) {
        nj.parent[i] = parent[(size_t) i];
        nj.child[i].nChild = nChild[(size_t) i];
        for (int k = 0; k < 3; k++) nj.child[i].child[k] = child[(size_t) (3 * i + k)];
        nj.branchlength[i] = bl[(size_t) i];
    }
    nj.maxnode = res.maxnode;
    nj.root = res.root;
    nj.recomputeProfiles(nj.distanceMatrix);
}

}  // namespace veryfasttree
#endif
