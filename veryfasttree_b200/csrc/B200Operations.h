// B200Operations.h -- the reference-facing C++ side of the drop-in boundary.
//
// The reference selects its arithmetic backend through a compile-time policy class
// `Operations<Precision>` (/root/reference/src/operations/BasicOperations.h:16-39): a by-value member
// of NeighbourJoining (NeighbourJoining.h:256) with ten per-element primitives, an ALIGNMENT and an
// Allocator.  `veryfasttree::B200Operations<P>` is that class for the B200 path.  It has two halves:
//
//  1. The ten primitives, same names / argument meaning / in-place aliasing rules as
//     BasicOperations.  They are what is left of the per-element call sites once the hot loops are
//     batched (NJ.tcc:782,825,828,849,857,916,1359,2034,2035,2333,2377-2389,2426), they run on the
//     host and are arithmetically identical to `-ext NONE` (left-to-right sums, separate mul/add).
//     A GPU cannot live behind a 4..20-element call (the reference's own CudaOperations.cu shows
//     what happens when it tries), so these are deliberately NOT device calls.
//
//  2. The batched entry points: thin, exception-translating wrappers over the C-ABI of
//     include/vft_b200.h, one per loop of NeighbourJoining.tcc that is replaced (setBestHit,
//     transferBestHits/uniqueBestHits, setOutDistance, averageProfile, outProfile/updateOutProfile).
//     INTEGRATION.md shows the handful of call-site changes in NeighbourJoining.tcc.
//
// Errors: the C-ABI returns status codes; the wrappers throw std::invalid_argument, which is what
// main.cpp:673-678 catches (the reference's CudaOperations throws runtime_error, which main does
// not catch).  There is no CPU fallback: without a device the constructor of the context throws.
#ifndef VERYFASTTREE_B200OPERATIONS_H
#define VERYFASTTREE_B200OPERATIONS_H

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>

#include "../../include/vft_b200.h"

namespace veryfasttree {

// minimal aligned allocator (the reference uses boost::alignment::aligned_allocator; any
// std-conforming allocator with the same alignment works for its std::vector members)
template<typename T, std::size_t Align>
struct B200AlignedAllocator {
    typedef T value_type;
    B200AlignedAllocator() noexcept {}
    template<class U> B200AlignedAllocator(const B200AlignedAllocator<U, Align> &) noexcept {}
    template<class U> struct rebind { typedef B200AlignedAllocator<U, Align> other; };
    T *allocate(std::size_t n) {
        if (n == 0) return nullptr;
        void *p = nullptr;
        if (posix_memalign(&p, Align < sizeof(void *) ? sizeof(void *) : Align, n * sizeof(T)) != 0) throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, std::size_t) noexcept { std::free(p); }
    template<class U> bool operator==(const B200AlignedAllocator<U, Align> &) const noexcept { return true; }
    template<class U> bool operator!=(const B200AlignedAllocator<U, Align> &) const noexcept { return false; }
};

template<typename Precision>
class B200Operations {
public:
    // 32 keeps nCodeSize (= alignsz(nCodes, ALIGNMENT/sizeof(P)), NJ.tcc:221-222) equal to the AVX2
    // backend's, so host-side Profile vectors have the stride the AVX2-ordered goldens were made with
    static constexpr int ALIGNMENT = 32;
    using Allocator = B200AlignedAllocator<Precision, ALIGNMENT>;
    typedef Precision numeric_t;

    // ---- 1. the per-element surface (BasicOperations.h:20-39) --------------------------------
    inline void vector_multiply(numeric_t f1[], numeric_t f2[], int64_t n, numeric_t fOut[]) {
        for (int64_t k = 0; k < n; k++) fOut[k] = f1[k] * f2[k];
    }

    inline numeric_t vector_multiply_sum(numeric_t f1[], numeric_t f2[], int64_t n) {
        numeric_t acc = 0;
        for (int64_t k = 0; k < n; k++) acc += f1[k] * f2[k];
        return acc;
    }

    inline numeric_t vector_multiply3_sum(numeric_t f1[], numeric_t f2[], numeric_t f3[], int64_t n) {
        numeric_t acc = 0;
        for (int64_t k = 0; k < n; k++) acc += f1[k] * f2[k] * f3[k];
        return acc;
    }

    inline numeric_t vector_dot_product_rot(numeric_t f1[], numeric_t f2[], numeric_t fBy[], int64_t n) {
        numeric_t a = 0, b = 0;
        for (int64_t k0 = 0; k0 < n; k0 += 4) {           // the reference interleaves in groups of four
            for (int64_t k = k0; k < k0 + 4; k++) a += f1[k] * fBy[k];
            for (int64_t k = k0; k < k0 + 4; k++) b += f2[k] * fBy[k];
        }
        return a * b;
    }

    inline void vector_add(numeric_t fTot[], numeric_t fAdd[], int64_t n) {
        for (int64_t k = 0; k < n; k++) fTot[k] += fAdd[k];
    }

    inline numeric_t vector_sum(numeric_t f1[], int64_t n) {
        numeric_t acc = 0;
        for (int64_t k = 0; k < n; k++) acc += f1[k];
        return acc;
    }

    inline void vector_multiply_by(numeric_t f[], numeric_t fBy, int64_t n, numeric_t fOut[]) {
        for (int64_t k = 0; k < n; k++) fOut[k] = f[k] * fBy;
    }

    inline void vector_add_mult(numeric_t fTot[], numeric_t fAdd[], numeric_t weight, int64_t n) {
        for (int64_t k = 0; k < n; k++) fTot[k] += fAdd[k] * weight;
    }

    template<int row>
    inline void matrix_by_vector4(numeric_t mat[][row], numeric_t vec[], numeric_t out[]) {
        for (int j = 0; j < 4; j++) {
            double acc = 0;                                // double accumulator, BasicOperations.tcc:112-118
            for (int k = 0; k < 4; k++) acc += vec[k] * mat[k][j];
            out[j] = (numeric_t) acc;
        }
    }

    // exp() at the four -fastexp levels (BasicOperations.tcc:121-216): 0 libm double, 1 libm float,
    // 2 / 3 the Cephes-style rational P/Q in double / float with exact 2^m scaling
    inline void fastexp(numeric_t fTot[], int64_t n, int lvl) {
        for (int64_t k = 0; k < n; k++) {
            if (lvl == 0) fTot[k] = (numeric_t) std::exp((double) fTot[k]);
            else if (lvl == 1) fTot[k] = (numeric_t) std::exp((float) fTot[k]);
            else if (lvl == 2) fTot[k] = rational_exp<double, int64_t>((double) fTot[k]);
            else fTot[k] = rational_exp<float, int32_t>((float) fTot[k]);
        }
    }

    // ---- 2. the batched surface: one method per replaced loop -----------------------------------
    // Bound to a device context by the NeighbourJoining constructor (after seqsToProfiles):
    //   operations.attach(nSeqs, nPos, options.nCodes, options.useMatrix, options.fPostTotalTolerance)
    // nScratch: extra profile rows for the temporaries of the ML phase (stack Profiles of MLQuartetOptimize, upProfiles[])
    void attach(int64_t nSeqs, int64_t nPos, int nCodes, bool useMatrix, double fPostTotalTolerance, int device = 0, int64_t nScratch = 0) {
        vft_config cfg;
        std::memset(&cfg, 0, sizeof cfg);
        cfg.nScratch = nScratch;
        cfg.nSeqs = nSeqs; cfg.nPos = nPos; cfg.nCodes = nCodes;
        cfg.precision = (int32_t) (8 * sizeof(Precision));
        cfg.useMatrix = useMatrix ? 1 : 0;
        cfg.reduction = VFT_REDUCE_AVX2;
        cfg.device = device;
        cfg.fPostTotalTolerance = fPostTotalTolerance;
        vft_ctx *raw = nullptr;
        check(vft_ctx_create(&cfg, &raw));
        ctx.reset(raw, [](vft_ctx *c) { vft_ctx_destroy(c); });
    }
    bool attached() const { return (bool) ctx; }

    void uploadTables(const numeric_t *distances, const numeric_t *eigenval, const numeric_t *eigentot,
                      const numeric_t *codeFreq) { check(vft_upload_tables(ctx.get(), distances, eigenval, eigentot, codeFreq)); }
    void uploadLeaves(const uint8_t *codes) { check(vft_upload_leaves(ctx.get(), codes)); }                       // NJ.tcc:382-534
    void outProfileRebuild(const int64_t *ids, int64_t n) { check(vft_outprofile_rebuild(ctx.get(), ids, n)); }    // NJ.tcc:729
    void outProfileUpdate(int64_t old1, int64_t old2, int64_t nw, int64_t nActiveOld) {                            // NJ.tcc:943
        check(vft_outprofile_update(ctx.get(), old1, old2, nw, nActiveOld));
    }
    void averageProfile(int64_t out, int64_t id1, int64_t id2, double bionjWeight, double diameterOut) {            // NJ.tcc:2067
        check(vft_profile_average(ctx.get(), out, id1, id2, bionjWeight, diameterOut));
    }
    void outDistanceBatch(const int64_t *ids, int64_t n, int64_t nActive, double totdiam, numeric_t *out) {         // NJ.tcc:1012
        check(vft_out_distance_batch(ctx.get(), ids, n, nActive, totdiam, out));
    }
    void outDistanceAll(int64_t nActive, double totdiam, numeric_t *out, int64_t maxnode) {                         // NJ.tcc:4451-4464
        check(vft_out_distance_all(ctx.get(), nActive, totdiam, out, maxnode));
    }
    void distPairs(const int64_t *i, const int64_t *j, int64_t n, bool rawProfile, numeric_t *dist, numeric_t *weight) {   // NJ.tcc:1115-1122
        check(vft_dist_pairs(ctx.get(), i, j, n, rawProfile ? VFT_PAIRS_PROFILE_RAW : VFT_PAIRS_JOIN, dist, weight));
    }
    int64_t distOneVsAll(int64_t query, int64_t nActive, int64_t K, int64_t *j, numeric_t *dist, numeric_t *weight,
                         numeric_t *criterion) {                                                                    // NJ.tcc:3571 + :4471
        int64_t n = 0;
        check(vft_dist_one_vs_all(ctx.get(), query, nActive, K, j, dist, weight, criterion, &n));
        return n;
    }
    // the m list merges of a top-hits refresh, NJ.tcc:4477-4515 (own lists with their active ancestors resolved)
    void topHitsMerge(int64_t newnode, int64_t nActive, int64_t m, int64_t nLists, const int64_t *iNode, const int64_t *ownOffset,
                      const int64_t *ownJ, const numeric_t *ownDist, int64_t nAvail, const int64_t *allJ, const numeric_t *allDist,
                      int64_t *outCount, int64_t *outJ, numeric_t *outDist) {
        check(vft_tophits_merge(ctx.get(), newnode, nActive, m, nLists, iNode, ownOffset, ownJ, ownDist, nAvail, allJ, allDist, outCount,
                                outJ, outDist));
    }
    // likelihood phase: TransitionMatrix tables + Rates, pairLogLk (NJ.tcc:1192), posteriorProfile (NJ.tcc:2137)
    void uploadTransmat(const numeric_t *codeFreq, const numeric_t *eigenval, const numeric_t *eigeninv, const numeric_t *eigeninvT,
                        const numeric_t *statinv) {
        check(vft_upload_transmat(ctx.get(), codeFreq, eigenval, eigeninv, eigeninvT, statinv));
    }
    void syncRates(const numeric_t *rates, int64_t nRateCats, const int64_t *ratecat, double MLMinRelBranchLength,
                   double MLMinBranchLength, int fastexpLevel) {
        check(vft_sync_rates(ctx.get(), rates, nRateCats, ratecat, MLMinRelBranchLength, MLMinBranchLength, fastexpLevel));
    }
    void pairLogLkBatch(const int64_t *i, const int64_t *j, const double *length, int64_t n, double *loglk, double *siteLk = nullptr) {
        check(vft_pair_loglk_batch(ctx.get(), i, j, length, n, loglk, siteLk));
    }
    void posteriorProfile(int64_t out, int64_t id1, int64_t id2, double len1, double len2) {
        check(vft_posterior_profile(ctx.get(), out, id1, id2, len1, len2));
    }
    void posteriorProfileBatch(int64_t n, const int64_t *out, const int64_t *id1, const int64_t *id2, const double *len1, const double *len2) {
        check(vft_posterior_profile_batch(ctx.get(), n, out, id1, id2, len1, len2));                               // one tree level, NJ.tcc:3508-3542
    }
    // whole-tree sweeps: recomputeMLProfiles + treeLogLk (NJ.tcc:3508-3542, :5114-5259), setMLRates (NJ.tcc:5429-5488)
    double treeLogLk(int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child, const numeric_t *branchlength,
                     bool recomputeProfiles, const uint8_t *leafCodes = nullptr, double *siteLoglk = nullptr) {
        double lk = 0;
        check(vft_tree_loglk(ctx.get(), root, maxnode, nChild, child, branchlength, recomputeProfiles ? 1 : 0, leafCodes, &lk, siteLoglk));
        return lk;
    }
    void setMLRates(int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child, const numeric_t *branchlength,
                    int64_t nRateCats, double MLMinRelBranchLength, double MLMinBranchLength, int fastexpLevel, const uint8_t *leafCodes,
                    numeric_t *rates, int64_t *ratecat) {
        check(vft_set_ml_rates(ctx.get(), root, maxnode, nChild, child, branchlength, nRateCats, MLMinRelBranchLength, MLMinBranchLength,
                               fastexpLevel, leafCodes, rates, ratecat, nullptr));
    }
    // branch-length optimisation and ML NNI quartets as lock-step batches (csrc/ml_opt.cpp): MLPairOptimize (NJ.tcc:1790),
    // MLQuartetNNI (NJ.tcc:4885) over MLQuartetOptimize / onedimenmin / brent, the per-node body of
    // optimizeAllBranchLengths (NJ.tcc:5044-5058) and the whole sweep (NJ.tcc:5006-5112).  Needs a context created with
    // scratch rows (attach(..., nScratch)).
    void mlPairOptimizeBatch(const vft_ml_options &opt, int64_t n, const int64_t *idA, const int64_t *idB, double *length, double *loglk) {
        check(vft_ml_pair_optimize_batch(ctx.get(), &opt, n, idA, idB, length, loglk, nullptr));
    }
    void mlQuartetNNIBatch(const vft_ml_options &opt, int64_t n, const int64_t *ids, numeric_t *len, double *criteria, int32_t *choice,
                           int32_t *starTest, int64_t firstScratchRow) {
        check(vft_ml_quartet_nni_batch(ctx.get(), &opt, n, ids, len, criteria, choice, starTest, firstScratchRow, nullptr));
    }
    void mlStarOptimizeBatch(const vft_ml_options &opt, int64_t n, const int64_t *ids, numeric_t *len, int64_t firstScratchRow) {
        check(vft_ml_star_optimize_batch(ctx.get(), &opt, n, ids, len, firstScratchRow, nullptr));
    }
    // chooseNNI (NJ.tcc:4836-4852) and SHSupport (NJ.tcc:1126-1165) for sets of independent quartets / splits
    void chooseNNIBatch(int64_t n, const int64_t *ids, double pseudoWeight, bool logdist, double *criteria, int32_t *choice) {
        check(vft_choose_nni_batch(ctx.get(), n, ids, pseudoWeight, logdist ? 1 : 0, criteria, choice));
    }
    // the per-split body of testSplitsML (NJ.tcc:6884-6952): likelihoods + per-site likelihoods of the three topologies
    void mlSplitTestBatch(const vft_ml_options &opt, int64_t n, const int64_t *ids, const numeric_t *len, double *loglk, double *siteLk,
                          int32_t *choice, int32_t *badSplit, int64_t firstScratchRow) {
        check(vft_ml_split_test_batch(ctx.get(), &opt, n, ids, len, loglk, siteLk, choice, badSplit, firstScratchRow, nullptr));
    }
    // testSplitsML (NJ.tcc:6800-7000) over the whole tree: support[maxnode], returns SplitCount.nBadSplits
    int64_t mlTestSplits(const vft_ml_options &opt, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child,
                         const numeric_t *branchlength, int64_t nBootstrap, const int64_t *col, numeric_t *support) {
        int64_t nBad = 0;
        check(vft_ml_test_splits(ctx.get(), &opt, root, maxnode, nChild, child, branchlength, nBootstrap, col, support, &nBad, nullptr));
        return nBad;
    }
    void shSupportBatch(int64_t n, int64_t nBootstrap, const int64_t *col, const double *loglk, const double *siteLk, double *support) {
        check(vft_sh_support_batch(ctx.get(), n, nBootstrap, col, loglk, siteLk, support));
    }
    void mlOptimizeBranchLengths(const vft_ml_options &opt, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child,
                                 numeric_t *branchlength, bool referenceOrder) {
        check(vft_ml_optimize_branch_lengths(ctx.get(), &opt, root, maxnode, nChild, child, branchlength,
                                             referenceOrder ? VFT_ML_SCHEDULE_REFERENCE : VFT_ML_SCHEDULE_LEVELS, nullptr));
    }

private:
    std::shared_ptr<vft_ctx> ctx;        // copies of the policy object share one device context

    static void check(int rc) {
        if (rc != VFT_OK) throw std::invalid_argument(std::string("B200 backend: ") + vft_last_error());
    }

    template<typename F, typename I>
    static numeric_t rational_exp(F x) {
        // range reduction e^x = 2^m * e^r with ln2 split in two parts, then e^r = 1 + 2r P(r^2)/(Q(r^2) - r P(r^2))
        F px = std::floor((F) 1.4426950408889634073599 * x + (F) 0.5);
        I m = (I) px;
        x -= px * (F) 6.93145751953125E-1;
        x -= px * (F) 1.42860682030941723212E-6;
        const F xx = x * x;
        px = (F) 1.26177193074810590878E-4;
        px *= xx; px += (F) 3.02994407707441961300E-2;
        px *= xx; px += (F) 9.99999999999999999910E-1;
        px *= x;
        F qx = (F) 3.00198505138664455042E-6;
        qx *= xx; qx += (F) 2.52448340349684104192E-3;
        qx *= xx; qx += (F) 2.27265548208155028766E-1;
        qx *= xx; qx += (F) 2.00000000000000000009E0;
        x = px / (qx - px);
        x = (F) (1.0 + 2.0 * x);
        F scale;
        if (sizeof(F) == 8) { int64_t bits = ((int64_t) m + 1023) << 52; std::memcpy(&scale, &bits, 8); }
        else { int32_t bits = ((int32_t) m + 127) << 23; std::memcpy(&scale, &bits, 4); }
        if (sizeof(F) == 8) return (numeric_t) ((numeric_t) x * scale);
        return (numeric_t) x * (numeric_t) scale;
    }
};

}  // namespace veryfasttree

#endif
