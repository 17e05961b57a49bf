// TEST INFRASTRUCTURE ONLY -- the CPU double of the device-resident join loop.
//
// The product runs veryfasttree_b200/csrc/nj_loop_logic.h as one CUDA thread block over state in HBM
// (vft_cuda.cu).  This file compiles THE SAME logic source for the host with a one-thread execution environment
// (tid 0 of 1, barriers are no-ops) over the CPU restatement of the kernels (vft_oracle.c through the public C-ABI), so
// that the block-parallel re-expression of the reference's join loop can be checked against the golden trees on a box
// without a GPU.  Linked into oracle/libvftoracle.so only.
#include "../veryfasttree_b200/csrc/nj_loop.h"
#include "../veryfasttree_b200/csrc/nj_loop_logic.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

namespace {

template<typename P>
struct CpuEnv {
    vft_ctx *ctx;
    njl::State<P> *st;
    int rc = VFT_OK;
    int tid() const { return 0; }
    int nt() const { return 1; }
    void sync() {}
    int atomicAddI(int32_t *p, int v) { const int o = *p; *p += v; return o; }
    int atomicExchI(int32_t *p, int v) { const int o = *p; *p = v; return o; }
    void atomicAddL(int64_t *p, int64_t v) { *p += v; }
    bool syncOr(int pred) { return pred != 0; }
    int64_t clock() const { return 0; }
    int32_t blockMin(uint64_t *, int32_t *, uint64_t k, int32_t idx, uint64_t *keyOut) { if (keyOut) *keyOut = k; return idx; }
    int32_t blockSum(int32_t *, int32_t v) { return v; }
    void evalOut(const int32_t *ids, int n, int32_t nActive) {
        std::vector<int64_t> id((size_t) n);
        std::vector<P> out((size_t) n);
        for (int k = 0; k < n; k++) id[(size_t) k] = ids[k];
        const int r = vft_out_distance_batch(ctx, id.data(), n, nActive, st->sc->totdiam, out.data());
        if (r != VFT_OK) rc = r;
        for (int k = 0; k < n; k++) { st->freshVal[ids[k]] = out[(size_t) k]; st->freshEpoch[ids[k]] = st->sc->epoch; }
        st->sc->outprofileOps += n;
    }
    void evalPairs(const int32_t *pairs, int n, P *outD) {
        std::vector<int64_t> a((size_t) n), b((size_t) n);
        std::vector<P> w((size_t) n);
        for (int k = 0; k < n; k++) { a[(size_t) k] = pairs[2 * k]; b[(size_t) k] = pairs[2 * k + 1]; }
        const int r = vft_dist_pairs(ctx, a.data(), b.data(), n, VFT_PAIRS_JOIN, outD, w.data());
        if (r != VFT_OK) rc = r;
        for (int k = 0; k < n; k++) { if (a[(size_t) k] < st->sc->nSeqs && b[(size_t) k] < st->sc->nSeqs) st->sc->seqOps++; else st->sc->profileOps++; }
    }
};

template<typename P>
struct Loop {
    vft_ctx *ctx;
    njl::Scalars sc{};
    njl::State<P> st{};
    std::vector<int32_t> parent, up, child, nOutAct, freshEpoch, wantStamp, hitJ, hitCount, age, visJ, topvisible, reqOut, reqA, reqB, uJ, uSlot, lSlot;
    std::vector<P> branchlength, diameter, outDist, freshVal, hitDist, visDist, pairD, pairW, tmpP;
    std::vector<int64_t> joins;
    std::vector<unsigned char> scratch;
    int64_t nSteps = 0;

    void init(vft_ctx *c, const vft_config &cfg, const vft_nj_options &opt, int64_t m, int64_t nTV) {
        ctx = c;
        const int64_t N = cfg.nSeqs, M = 2 * N;
        int cap = 32;
        while (cap < 2 * m) cap <<= 1;
        sc.nSeqs = N; sc.maxnodes = M; sc.m = (int32_t) m; sc.nTV = (int32_t) nTV; sc.cap = cap;
        sc.tophitAgeLimit = (int32_t) std::max<int64_t>(1, (int64_t) (0.5 + std::log((double) m) / std::log(2.0)));
        sc.nRefreshMin = (int32_t) (int64_t) (0.5 + m * opt.tophitsRefresh);
        sc.nResetOutProfile = opt.nResetOutProfile; sc.staleOutLimit = opt.staleOutLimit; sc.fResetOutProfile = opt.fResetOutProfile;
        sc.Lbytes = cfg.nPos; sc.profBytes = cfg.nPos * ((int64_t) cfg.nCodes * (cfg.precision / 8) + cfg.precision / 8 + 1);
        parent.assign(M, -1); up.assign(M, 0); child.assign(3 * M, -1); nOutAct.assign(M, 0); freshEpoch.assign(M, -1); wantStamp.assign(M, -1);
        hitJ.assign((size_t) M * m, -1); hitCount.assign(M, 0); age.assign(M, 0); visJ.assign(M, -1); topvisible.assign(nTV, -1);
        branchlength.assign(M, 0); diameter.assign(M, 0); outDist.assign(M, 0); freshVal.assign(M, 0); hitDist.assign((size_t) M * m, 0); visDist.assign(M, 0);
        st.capOut = (int32_t) (4 * cap + 2 * nTV + 8); st.capPair = (int32_t) (cap + 2 * m + 8);
        reqOut.assign(st.capOut, -1); reqA.assign(st.capPair, -1); reqB.assign(st.capPair, -1); pairD.assign(st.capPair, 0); pairW.assign(st.capPair, 0);
        uJ.assign(cap, -1); uSlot.assign(cap, -1); lSlot.assign(2 * m, -1); tmpP.assign(cap, 0);
        joins.assign((size_t) 2 * N, -1);
        st.sc = &sc;
        st.parent = parent.data(); st.up = up.data(); st.child = child.data(); st.branchlength = branchlength.data(); st.diameter = diameter.data();
        st.outDist = outDist.data(); st.nOutAct = nOutAct.data(); st.freshVal = freshVal.data(); st.freshEpoch = freshEpoch.data(); st.wantStamp = wantStamp.data();
        st.hitJ = hitJ.data(); st.hitDist = hitDist.data(); st.hitCount = hitCount.data(); st.age = age.data(); st.visJ = visJ.data(); st.visDist = visDist.data();
        st.topvisible = topvisible.data(); st.reqOut = reqOut.data(); st.reqA = reqA.data(); st.reqB = reqB.data(); st.pairD = pairD.data(); st.pairW = pairW.data();
        st.uJ = uJ.data(); st.uSlot = uSlot.data(); st.lSlot = lSlot.data(); st.tmpP = tmpP.data(); st.joins = joins.data();
        scratch.assign(njl::scratch_bytes<P>(cap, 1) + 64, 0);
        sc.status = njl::ST_RUNNING; sc.resume = njl::RS_SEARCH; sc.epoch = 1; sc.stamp = 1; sc.hintEpoch = -1; sc.hintJoinSlot = -1;
    }

    void upload(const vftx_loop_image &g, int32_t resume) {
        const size_t M = (size_t) sc.maxnodes, m = (size_t) sc.m;
        std::copy(g.parent, g.parent + M, parent.begin()); std::copy(g.up, g.up + M, up.begin()); std::copy(g.child, g.child + 3 * M, child.begin());
        std::copy((const P *) g.branchlength, (const P *) g.branchlength + M, branchlength.begin());
        std::copy((const P *) g.diameter, (const P *) g.diameter + M, diameter.begin());
        std::copy((const P *) g.outDist, (const P *) g.outDist + M, outDist.begin());
        std::copy(g.nOutAct, g.nOutAct + M, nOutAct.begin());
        std::copy(g.hitJ, g.hitJ + M * m, hitJ.begin()); std::copy((const P *) g.hitDist, (const P *) g.hitDist + M * m, hitDist.begin());
        std::copy(g.hitCount, g.hitCount + M, hitCount.begin()); std::copy(g.age, g.age + M, age.begin());
        std::copy(g.visJ, g.visJ + M, visJ.begin()); std::copy((const P *) g.visDist, (const P *) g.visDist + M, visDist.begin());
        std::copy(g.topvisible, g.topvisible + sc.nTV, topvisible.begin());
        sc.maxnode = (int32_t) g.maxnode; sc.nActive = (int32_t) g.nActive; sc.topvisibleAge = (int32_t) g.topvisibleAge;
        sc.nActiveOutProfileReset = (int32_t) g.nActiveOutProfileReset; sc.totdiam = g.totdiam;
        sc.status = njl::ST_RUNNING; sc.resume = resume; sc.epoch++; sc.hintEpoch = -1; sc.hintJoinSlot = -1; sc.jdValid = 0;
        sc.nOutReq = 0; sc.nPairReq = 0;
    }

    void download(vftx_loop_image &g) {
        const size_t M = (size_t) sc.maxnodes, m = (size_t) sc.m;
        std::copy(parent.begin(), parent.end(), g.parent); std::copy(up.begin(), up.end(), g.up); std::copy(child.begin(), child.end(), g.child);
        std::copy(branchlength.begin(), branchlength.end(), (P *) g.branchlength); std::copy(diameter.begin(), diameter.end(), (P *) g.diameter);
        std::copy(outDist.begin(), outDist.end(), (P *) g.outDist); std::copy(nOutAct.begin(), nOutAct.end(), g.nOutAct);
        std::copy(hitJ.begin(), hitJ.end(), g.hitJ); std::copy(hitDist.begin(), hitDist.end(), (P *) g.hitDist);
        std::copy(hitCount.begin(), hitCount.end(), g.hitCount); std::copy(age.begin(), age.end(), g.age);
        std::copy(visJ.begin(), visJ.end(), g.visJ); std::copy(visDist.begin(), visDist.end(), (P *) g.visDist);
        std::copy(topvisible.begin(), topvisible.end(), g.topvisible);
        g.maxnode = sc.maxnode; g.nActive = sc.nActive; g.topvisibleAge = sc.topvisibleAge; g.nActiveOutProfileReset = sc.nActiveOutProfileReset;
        g.totdiam = sc.totdiam;
        (void) M; (void) m;
    }

    int evalRequests() {
        const int nOut = sc.nOutReq, nPair = sc.nPairReq;
        if (nOut + nPair == 0) return VFT_OK;
        std::vector<int64_t> ids((size_t) nOut), a((size_t) nPair), b((size_t) nPair);
        std::vector<P> od((size_t) nOut);
        for (int k = 0; k < nOut; k++) ids[(size_t) k] = reqOut[(size_t) k];
        for (int k = 0; k < nPair; k++) { a[(size_t) k] = reqA[(size_t) k]; b[(size_t) k] = reqB[(size_t) k]; }
        const int rc = vft_eval_batch(ctx, ids.data(), nOut, sc.nActive, sc.totdiam, od.data(), a.data(), b.data(), nPair, VFT_PAIRS_JOIN, pairD.data(), pairW.data());
        if (rc != VFT_OK) return rc;
        for (int k = 0; k < nOut; k++) { freshVal[(size_t) reqOut[(size_t) k]] = od[(size_t) k]; freshEpoch[(size_t) reqOut[(size_t) k]] = sc.epoch; }
        sc.outprofileOps += nOut; sc.nPairHit += nPair;
        for (int k = 0; k < nPair; k++) { if (a[(size_t) k] < sc.nSeqs && b[(size_t) k] < sc.nSeqs) sc.seqOps++; else sc.profileOps++; }
        return VFT_OK;
    }

    int run(vftx_loop_status *out) {
        CpuEnv<P> env{ctx, &st};
        njl::Scratch<P> sm;
        njl::scratch_carve<P>(sm, scratch.data(), sc.cap, 1);
        njl::Logic<P, CpuEnv<P>> logic(st, sm, env);
        int rc = VFT_OK;
        while (rc == VFT_OK && sc.status == njl::ST_RUNNING) {
            logic.step();
            nSteps++;
            if (env.rc != VFT_OK) { rc = env.rc; break; }
            if (sc.jdValid) {                                // the join's profile arithmetic (k_average on the device)
                if (sc.jdUpdate) rc = vft_profile_average_update(ctx, sc.jdNew, sc.jdI, sc.jdJ, -1.0, sc.jdDiameter, sc.jdNActiveOld);
                else rc = vft_profile_average(ctx, sc.jdNew, sc.jdI, sc.jdJ, -1.0, sc.jdDiameter);
                sc.jdValid = 0;
                if (rc != VFT_OK) break;
            }
            if (sc.status == njl::ST_NEED_REBUILD) {         // NJ.tcc:3012-3033
                double td = 0;
                for (int64_t i = 0; i < sc.maxnode; i++) if (parent[(size_t) i] < 0) td += (double) diameter[(size_t) i];
                sc.totdiam = td;
                rc = vft_outprofile_rebuild(ctx, nullptr, sc.nActive);
                sc.nActiveOutProfileReset = sc.nActive;
                sc.status = njl::ST_RUNNING;
                if (rc != VFT_OK) break;
            }
            if (sc.status == njl::ST_RUNNING) rc = evalRequests();
        }
        if (out) {
            out->status = sc.status; out->resume = sc.resume; out->visfixPending = sc.visfixPending; out->newnode = sc.jdNew;
            out->nActive = sc.nActive; out->maxnode = sc.maxnode; out->nJoins = sc.nJoins; out->nRefresh = sc.nRefresh;
            out->nVisibleUpdate = sc.nVisibleUpdate; out->nHillBetter = sc.nHillBetter; out->nReset = sc.nReset; out->nInlineOut = sc.nInlineOut;
            out->nInlinePair = sc.nInlinePair; out->nPairHit = sc.nPairHit; out->nRebuild = sc.nRebuild; out->nSteps = nSteps;
        }
        return rc;
    }
};

}  // namespace

struct vftx_loop {
    int precision;
    Loop<float> *f = nullptr;
    Loop<double> *d = nullptr;
};

extern "C" int vftx_loop_create(vft_ctx *ctx, const vft_nj_options *opt, int64_t m, int64_t nTV, vftx_loop **out) {
    vft_config cfg;
    if (!ctx || !opt || !out || m < 4 || nTV < 1 || vft_get_config(ctx, &cfg, nullptr) != VFT_OK) return VFT_EINVAL;
    vftx_loop *lp = new vftx_loop();
    lp->precision = cfg.precision;
    if (cfg.precision == 32) { lp->f = new Loop<float>(); lp->f->init(ctx, cfg, *opt, m, nTV); }
    else { lp->d = new Loop<double>(); lp->d->init(ctx, cfg, *opt, m, nTV); }
    *out = lp;
    return VFT_OK;
}
extern "C" int vftx_loop_upload(vftx_loop *lp, const vftx_loop_image *img, int32_t resume) {
    if (!lp || !img) return VFT_EINVAL;
    if (lp->f) lp->f->upload(*img, resume); else lp->d->upload(*img, resume);
    return VFT_OK;
}
extern "C" int vftx_loop_run(vftx_loop *lp, vftx_loop_status *st) {
    if (!lp) return VFT_EINVAL;
    return lp->f ? lp->f->run(st) : lp->d->run(st);
}
extern "C" int vftx_loop_download(vftx_loop *lp, vftx_loop_image *img) {
    if (!lp || !img) return VFT_EINVAL;
    if (lp->f) lp->f->download(*img); else lp->d->download(*img);
    return VFT_OK;
}
extern "C" int64_t vftx_loop_joins(vftx_loop *lp, int64_t *out, int64_t maxJoins) {
    if (!lp || !out) return 0;
    const std::vector<int64_t> &j = lp->f ? lp->f->joins : lp->d->joins;
    const int64_t n = std::min<int64_t>(maxJoins, (int64_t) j.size() / 2);
    std::copy(j.begin(), j.begin() + 2 * n, out);
    return n;
}
extern "C" int vftx_loop_destroy(vftx_loop *lp) {
    if (!lp) return VFT_OK;
    delete lp->f; delete lp->d; delete lp;
    return VFT_OK;
}
