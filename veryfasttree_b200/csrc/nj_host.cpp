// nj_host.cpp -- host-side caller of the B200 hot path: the NJ + TopHits phase.
//
// This is the code "either side of the kernels" (SURVEY.md §8): the reference's
// NeighbourJoining ctor tail (NeighbourJoining.tcc:237-260) and fastNJ() with the top-hits
// heuristic (NJ.tcc:2796-3155, :3746-4833), written from scratch as a *batch producer* for the
// C-ABI in include/vft_b200.h.  The reference evaluates one pair at a time from inside OpenMP
// loops and refreshes out-distances lazily from inside setCriterion; here every host loop is
// preceded by a pre-scan that collects the pair distances and out-distances it can possibly
// need, fetches them in ONE device call, and then replays the reference's decisions from the
// fetched values.  The prefetch is only a hint: a value that was not prefetched is fetched on
// demand (counted in nOutSingleFetch / nPairSingleFetch), so the decisions -- and therefore the
// join order, the top-hit lists and the tree -- are exactly those of the reference at
// `-threads 1`, whatever the hints are.
//
// Orderings: the reference sorts with non-strict comparators (NJ.tcc:7285-7311) through
// boost's spinsort; oracle/psort_probe.cpp shows the outcome is always "key ascending, ties in
// REVERSE input order".  rsort() below reproduces exactly that.
//
// Not supported (rejected with VFT_EINVAL, documented in DESIGN.md): -fastest / 2nd-level top
// hits, topological constraints, -slow.  -bionj is supported (BIONJ weights, NJ.tcc:2921-2966).
#include "../../include/vft_b200.h"
#include "nj_loop.h"
extern "C" void vftx_set_error(const char *msg);      // the library's vft_last_error text (vft_cuda.cu / oracle)

#include <algorithm>
#include <omp.h>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

#ifdef VFT_HOSTPROF
static double g_hp[32]; static const char *g_hpName[32]; static double *g_hpCalls = nullptr; static long g_hpN[32];
struct HP { int k; std::chrono::steady_clock::time_point t0; double c0; HP(int k, const char *nm) : k(k), t0(std::chrono::steady_clock::now()), c0(g_hpCalls ? *g_hpCalls : 0) { g_hpName[k] = nm; }
    ~HP() { g_hp[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - ((g_hpCalls ? *g_hpCalls : 0) - c0); g_hpN[k]++; } };
#define HPROF(k, nm) HP hp_##k(k, nm)
#else
#define HPROF(k, nm)
#endif

struct Children { int nChild = 0; int64_t child[3] = {-1, -1, -1}; };

// "key ascending, ties in reverse input order" == psort() with the reference's comparators
template<class T, class Less>
void rsort(std::vector<T> &v, Less strictLess) {
    std::reverse(v.begin(), v.end());
    std::stable_sort(v.begin(), v.end(), strictLess);
}

struct DeviceError { int code; };

// orderable integer keys: ascending unsigned order == ascending floating order
inline uint64_t orderKey(float x) { if (x == 0) x = 0; uint32_t u; std::memcpy(&u, &x, 4); return (u & 0x80000000u) ? (uint32_t) ~u : (u | 0x80000000u); }
inline uint64_t orderKey(double x) { if (x == 0) x = 0; uint64_t u; std::memcpy(&u, &x, 8); return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull); }

// psort order for an array of records with an integer key: key ascending, ties by input position
// DESCENDING.  The usual case -- float criteria, or (i,j) keys with one i: the keys span less than 2^32 -- is a stable
// LSD radix sort (8-bit digits, digits shared by every key skipped) of the records taken in REVERSED input order, which
// leaves equal keys in reverse input order; lists here are a few hundred entries, where this is ~4x cheaper than a
// comparison sort.  Scratch is per thread.
template<class T, class KeyFn>
void rsortByKey(std::vector<T> &v, KeyFn keyOf) {
    const size_t n = v.size();
    if (n < 2) return;
    static thread_local std::vector<std::pair<uint64_t, uint32_t>> kv;
    static thread_local std::vector<uint64_t> k1, k2;
    static thread_local std::vector<T> tmp;
    k1.resize(n);
    uint64_t lo = ~0ull, hi = 0;
    for (size_t k = 0; k < n; k++) {
        const uint64_t key = keyOf(v[k]);
        k1[k] = key;
        lo = std::min(lo, key); hi = std::max(hi, key);
    }
    tmp.resize(n);
    if (hi - lo < (1ull << 32) && n < (1ull << 31)) {
        // word r = (key - lo) << 32 | source position, for the sources in reversed order
        k2.resize(n);
        for (size_t r = 0; r < n; r++) { const size_t src = n - 1 - r; k2[r] = ((k1[src] - lo) << 32) | (uint32_t) src; }
        if (n <= 32) {
            // tiny lists: insertion sort on the key half only (stable)
            for (size_t a = 1; a < n; a++) {
                const uint64_t x = k2[a];
                size_t c = a;
                while (c > 0 && (k2[c - 1] >> 32) > (x >> 32)) { k2[c] = k2[c - 1]; c--; }
                k2[c] = x;
            }
            for (size_t k = 0; k < n; k++) tmp[k] = v[(uint32_t) k2[k]];
            v.swap(tmp);
            return;
        }
        uint64_t *src = k2.data(), *dst = k1.data();
        const uint64_t span = hi - lo;
        for (int shift = 32; shift < 64 && (span >> (shift - 32)) != 0; shift += 8) {
            uint32_t count[256] = {0};
            for (size_t k = 0; k < n; k++) count[(src[k] >> shift) & 255]++;
            if (count[(src[0] >> shift) & 255] == n) continue;          // every key has this digit
            uint32_t sum = 0;
            for (int d = 0; d < 256; d++) { const uint32_t c = count[d]; count[d] = sum; sum += c; }
            for (size_t k = 0; k < n; k++) dst[count[(src[k] >> shift) & 255]++] = src[k];
            std::swap(src, dst);
        }
        for (size_t k = 0; k < n; k++) tmp[k] = v[(uint32_t) src[k]];
    } else {
        kv.resize(n);
        for (size_t k = 0; k < n; k++) kv[k] = {k1[k], (uint32_t) (n - 1 - k)};      // reversed position: plain pair order
        std::sort(kv.begin(), kv.end());
        for (size_t k = 0; k < n; k++) tmp[k] = v[n - 1 - kv[k].second];
    }
    v.swap(tmp);
}

template<typename P>
class NJ {
public:
    // node ids fit 31 bits (maxnodes = 2N < 2^31, checked in vft_nj_build): 32-bit ids halve the
    // memory of the top-hit lists (the reference's are int64_t, NJ.h:192-217) and the copy traffic
    typedef int32_t id_t;
    struct Hit { id_t j; P dist; };                                // NJ.h:214-217
    struct Besthit { id_t i, j; P weight, dist, criterion; };      // NJ.h:192-200
    struct TopHitsList { std::vector<Hit> hits; int64_t hitSource = -1; int64_t age = 0; };

    NJ(vft_ctx *ctx, const vft_config &cfg, const vft_nj_options &opt, vft_nj_result *res, const uint8_t *codes)
            : ctx(ctx), cfg(cfg), opt(opt), res(res), nSeqs(cfg.nSeqs), nPos(cfg.nPos) {
        maxnodes = 2 * nSeqs;
        maxnode = nSeqs;
        parent.assign(maxnodes, -1);
        child.resize(maxnodes);
        branchlength.assign(maxnodes, 0);
        diameter.assign(maxnodes, 0);
        varDiameter.assign(maxnodes, 0);
        outDistances.assign(maxnodes, 0);
        nOutDistActive.assign(maxnodes, (int32_t) std::min<int64_t>(nSeqs * 10, 2000000000));      // "never computed": stale at any nActive
        freshVal.assign(maxnodes, 0);
        freshEpoch.assign(maxnodes, -1);
        wantEpoch.assign(maxnodes, -1);
        hintedEpoch.assign(maxnodes, -1);
        up.resize(maxnodes);
        for (int64_t i = 0; i < maxnodes; i++) up[i] = (int32_t) i;
        hostThreads = opt.hostThreads > 0 ? opt.hostThreads : std::max(1, std::min(16, omp_get_num_procs()));
        // EXPERIMENTAL, off unless VFT_SPECULATION=1: bit-identical trees and a 98 % hit rate, but measured SLOWER on the
        // B200 (C2 1.68 s vs 1.26 s; 100k x 1287 aa 30.6 s vs 26.5 s): each speculated join costs three more launches of
        // host API time, and the hints for the search after next still need their own synchronous call.  It pays only
        // once those go out asynchronously too.
        specEnabled = opt.prefetch && !opt.bionj && std::getenv("VFT_SPECULATION") != nullptr && std::getenv("VFT_SPECULATION")[0] == '1';
        selfdistH.assign(maxnodes, 0); selfweightH.assign(maxnodes, 0); selfKnown.assign(maxnodes, 0);
        specSeenP.assign(maxnodes, 0); specSeenO.assign(maxnodes, 0);
        // nGaps(i) = nPos - selfweight[i] (NJ.tcc:249-252, :3762): gap/unknown columns of leaf i
        leafGaps.assign(nSeqs, 0);
        for (int64_t i = 0; i < nSeqs; i++) {
            int64_t g = 0;
            for (int64_t p = 0; p < nPos; p++) g += codes[i * nPos + p] >= cfg.nCodes;
            leafGaps[i] = g;
            selfweightH[i] = (P) (nPos - g); selfdistH[i] = 0; selfKnown[i] = 1;      // NJ.tcc:249-252
        }
    }

    // ---- NeighbourJoining ctor tail, NJ.tcc:237-260 ------------------------------------------
    void init() {
        check(timed([&] { return vft_outprofile_rebuild(ctx, nullptr, nSeqs); }));
        totdiam = 0.0;
        std::vector<P> od(maxnodes);
        check(timed([&] { return vft_out_distance_all(ctx, nSeqs, totdiam, od.data(), maxnodes); }));
        for (int64_t i = 0; i < nSeqs; i++) { outDistances[i] = od[i]; nOutDistActive[i] = nSeqs; }
        newEpoch(nSeqs);
    }

    void fastNJ();                       // NJ.tcc:2796-3155

    int64_t root = -1;
    int64_t maxnode;
    std::vector<int64_t> parent;
    std::vector<Children> child;
    std::vector<P> branchlength;

private:
    vft_ctx *ctx;
    vft_config cfg;
    vft_nj_options opt;
    vft_nj_result *res;
    int64_t nSeqs, nPos, maxnodes;
    std::vector<P> diameter, varDiameter, outDistances;
    std::vector<int32_t> nOutDistActive;     // (32-bit like the node ids: the per-join scans are bound by cache misses on these arrays)
    double totdiam = 0;

    // top-hits state, NJ.h:225-248
    int64_t m = 0;
    std::vector<TopHitsList> topHitsLists;
    std::vector<Hit> visible;
    std::vector<int64_t> topvisible;
    int64_t topvisibleAge = 0;

    void check(int rc) { res->nDeviceCalls++; if (rc != VFT_OK) throw DeviceError{rc}; }
    // wall time spent inside ABI calls (device + its synchronisation), for the host/device split
    // host-side time per section of the phase (wall clock minus time inside ABI calls)
    struct Section {
        NJ *nj; int k; std::chrono::steady_clock::time_point t0; double calls0;
        Section(NJ *nj, int k) : nj(nj), k(k), t0(std::chrono::steady_clock::now()), calls0(nj->res->secondsInCalls) {}
        ~Section() {
            double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            nj->res->secondsHost[k] += wall - (nj->res->secondsInCalls - calls0);
        }
    };
    template<class F> int timed(F f) {
        auto t0 = std::chrono::steady_clock::now();
        int rc = f();
        res->secondsInCalls += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    }

    // ---- fresh out-distance service ------------------------------------------------------------
    // A fresh value is a pure function of (node, out-profile, nActive, totdiam); all three are
    // fixed between two joins = one "epoch".
    int64_t epoch = 0, epochActive = 0;
    std::vector<P> freshVal;
    std::vector<int32_t> freshEpoch, wantEpoch, hintedEpoch;
    std::vector<int64_t> leafGaps;
    std::vector<int64_t> wantIds;
    std::vector<P> wantVals;

    int64_t hintGen = 0;             // generation of the pair cache: a list is hinted at most once per generation
    void clearPairs() { pairCache.clear(); hintGen++; }
    void newEpoch(int64_t nActive) { epoch++; epochActive = nActive; clearPairs(); }

    bool stale(int64_t i, int64_t nActive) const {       // trigger of NJ.tcc:1092-1098
        int64_t nDiffAllow = opt.tophitsMult > 0 ? (int64_t) (nActive * opt.staleOutLimit) : 0;
        return nOutDistActive[i] - nActive > nDiffAllow;
    }

    // hint: these nodes may get setOutDistance(.., nActive) soon
    void wantOut(int64_t i, int64_t nActive, bool evenIfNotStale = false) {
        if (!opt.prefetch || i < 0 || parent[i] >= 0) return;
        if (nOutDistActive[i] == nActive || freshEpoch[i] == epoch || wantEpoch[i] == epoch) return;
        if (!evenIfNotStale && !stale(i, nActive)) return;
        wantEpoch[i] = epoch;             // queued; becomes valid (freshEpoch) in flushOut
        wantIds.push_back(i);
    }

    // ONE device call for everything queued so far: out-distance requests and pair requests
    void flush(int64_t nActive) {
        if (wantIds.empty() && reqI.empty()) return;
        HPROF(12, "flush(host part)");
        wantVals.resize(wantIds.size());
        reqD.resize(reqI.size()); reqW.resize(reqI.size());
        check(timed([&] { return vft_eval_batch(ctx, wantIds.data(), (int64_t) wantIds.size(), nActive, totdiam, wantVals.data(),
                             reqI.data(), reqJ.data(), (int64_t) reqI.size(), VFT_PAIRS_JOIN, reqD.data(), reqW.data()); }));
        for (size_t k = 0; k < wantIds.size(); k++) { freshVal[wantIds[k]] = wantVals[k]; freshEpoch[wantIds[k]] = epoch; }
        for (size_t k = 0; k < reqI.size(); k++)
            if (reqCached[k]) pairCache.put(pkey(reqI[k], reqJ[k]), DW{reqD[k], reqW[k]});
        wantIds.clear();
        reqI.clear(); reqJ.clear(); reqCached.clear();
    }
    void flushOut(int64_t nActive) { flush(nActive); }
    void flushPairs() { flush(epochActive); }

    // setOutDistance, NJ.tcc:1012-1053 (the arithmetic lives behind vft_out_distance_batch)
    void setOutDistance(int64_t i, int64_t nActive) {
        if (nOutDistActive[i] == nActive) return;
        if (freshEpoch[i] == epoch && epochActive == nActive) {
            res->nOutPrefetchHit++;
        } else {
            P v;
            check(timed([&] { return vft_out_distance_batch(ctx, &i, 1, nActive, totdiam, &v); }));
            freshVal[i] = v;
            if (epochActive == nActive) freshEpoch[i] = epoch;
            res->nOutSingleFetch++;
        }
        outDistances[i] = freshVal[i];
        nOutDistActive[i] = nActive;
    }

    // ---- pair distance service (distance half of setDistCriterion, NJ.tcc:1115-1122) -----------
    struct DW { P dist, weight; };
    // (i,j) -> distance of the current out-profile epoch: open addressing, cleared in O(1) by a generation stamp
    struct PairCache {
        struct Slot { uint64_t key; int64_t gen; DW val; };
        std::vector<Slot> tab = std::vector<Slot>(1 << 12, Slot{0, -1, DW{0, 0}});
        int64_t gen = 0;
        size_t used = 0;
        static size_t hash(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; return (size_t) k; }
        void clear() { gen++; used = 0; }
        DW *find(uint64_t key) {
            const size_t mask = tab.size() - 1;
            for (size_t h = hash(key) & mask;; h = (h + 1) & mask) {
                Slot &s = tab[h];
                if (s.gen != gen) return nullptr;
                if (s.key == key) return &s.val;
            }
        }
        void put(uint64_t key, DW v) {
            if (2 * (used + 1) > tab.size()) grow();
            const size_t mask = tab.size() - 1;
            for (size_t h = hash(key) & mask;; h = (h + 1) & mask) {
                Slot &s = tab[h];
                if (s.gen != gen) { s.key = key; s.gen = gen; s.val = v; used++; return; }
                if (s.key == key) { s.val = v; return; }
            }
        }
        void grow() {
            std::vector<Slot> old;
            old.swap(tab);
            tab.assign(old.size() * 2, Slot{0, -1, DW{0, 0}});
            used = 0;
            for (const Slot &s : old) if (s.gen == gen) put(s.key, s.val);
        }
    } pairCache;
    std::vector<int64_t> reqI, reqJ;
    std::vector<P> reqD, reqW;        // results of the last flush, by request slot

    uint64_t pkey(int64_t i, int64_t j) const { return (uint64_t) i * (uint64_t) maxnodes + (uint64_t) j; }

    std::vector<char> reqCached;

    // hint for a pair looked up later through pairDist() (small, sporadic lists)
    void wantPair(int64_t i, int64_t j) {
        if (!opt.prefetch) return;
        if (pairCache.find(pkey(i, j)) != nullptr) return;
        pairCache.put(pkey(i, j), DW{0, -1});       // weight -1 marks "requested"
        reqI.push_back(i); reqJ.push_back(j); reqCached.push_back(1);
    }

    // positional request for the big lists: the result is read back from slot s of reqD/reqW
    // right after the next flush (no hashing); -1 when prefetching is off
    int64_t slotPair(int64_t i, int64_t j) {
        if (!opt.prefetch) return -1;
        reqI.push_back(i); reqJ.push_back(j); reqCached.push_back(0);
        return (int64_t) reqI.size() - 1;
    }

    DW pairDist(int64_t i, int64_t j) {
        const DW *it = pairCache.find(pkey(i, j));
        if (it != nullptr && it->weight >= 0) { res->nPairPrefetchHit++; return *it; }
        DW r;
        check(timed([&] { return vft_dist_pairs(ctx, &i, &j, 1, VFT_PAIRS_JOIN, &r.dist, &r.weight); }));
        res->nPairSingleFetch++;
        return r;
    }

    // callable from host threads: no shared counters.  An exception must not leave an OpenMP region, so a failure is
    // recorded and thrown by the caller after the region (threadError)
    int threadError = VFT_OK;
    DW pairDistNoCache(int64_t i, int64_t j) {
        DW r{0, 0};
        int rc;
#pragma omp critical(vft_device_call)
        rc = vft_dist_pairs(ctx, &i, &j, 1, VFT_PAIRS_JOIN, &r.dist, &r.weight);
        if (rc != VFT_OK) {
#pragma omp atomic write
            threadError = rc;
        }
        return r;
    }

    // ---- criterion ------------------------------------------------------------------------------
    // setCriterion, NJ.tcc:1085-1113
    void setCriterion(int64_t nActive, Besthit &join) {
        if (join.i < 0 || join.j < 0 || parent[join.i] >= 0 || parent[join.j] >= 0) return;
        if (stale(join.i, nActive)) setOutDistance(join.i, nActive);
        if (stale(join.j, nActive)) setOutDistance(join.j, nActive);
        double outI = outDistances[join.i];
        if (nOutDistActive[join.i] != nActive) outI *= (nActive - 1) / (double) (nOutDistActive[join.i] - 1);
        double outJ = outDistances[join.j];
        if (nOutDistActive[join.j] != nActive) outJ *= (nActive - 1) / (double) (nOutDistActive[join.j] - 1);
        join.criterion = (P) (join.dist - (outI + outJ) / (double) (nActive - 2));
    }

    void hintCriterion(int64_t nActive, int64_t i, int64_t j) {
        if (i < 0 || j < 0 || parent[i] >= 0 || parent[j] >= 0) return;
        wantOut(i, nActive); wantOut(j, nActive);
    }

    // setDistCriterion, NJ.tcc:1115-1124
    void setDistCriterion(int64_t nActive, Besthit &hit) {
        DW r = pairDist(hit.i, hit.j);
        hit.dist = r.dist; hit.weight = r.weight;
        setCriterion(nActive, hit);
    }

    // activeAncestor, NJ.tcc:536-544: the root of the subtree that contains i.  Same answer as
    // walking parent[], but over a union-find shortcut array with path halving (the walk can be
    // hundreds of levels deep on ladder-like trees); halving is skipped inside host-thread regions.
    std::vector<int32_t> up;
    int64_t activeAncestor(int64_t i) {
        if (i < 0) return i;
        if (omp_in_parallel()) { while (up[i] != i) i = up[i]; return i; }
        while (up[i] != i) { const int64_t g = up[up[i]]; up[i] = g; i = g; }
        return i;
    }

    static void hitToBestHit(int64_t i, const Hit &hit, Besthit &out) {   // NJ.tcc:4627-4633
        out.i = i; out.j = hit.j; out.dist = hit.dist; out.criterion = (P) 1e20; out.weight = -1;
    }

    void hitsToBestHits(const std::vector<Hit> &hits, int64_t iNode, Besthit *out) {   // NJ.tcc:4615-4625
        for (size_t k = 0; k < hits.size(); k++) hitToBestHit(iNode, hits[k], out[k]);
    }

    // getVisible, NJ.tcc:546-557
    bool getVisible(int64_t nActive, int64_t iNode, Besthit &v) {
        if (iNode < 0 || parent[iNode] >= 0) return false;
        const Hit &h = visible[iNode];
        if (h.j < 0 || parent[h.j] >= 0) return false;
        hitToBestHit(iNode, h, v);
        setCriterion(nActive, v);
        return true;
    }

    // what getBestFromTopHits(iNode) will ask for (NJ.tcc:4267-4298)
    void hintList(int64_t nActive, int64_t iNode) {
        HPROF(13, "hintList");
        if (iNode < 0 || parent[iNode] >= 0) return;
        if (hintedEpoch[iNode] == hintGen) return;       // already queued since the pair cache was last cleared
        hintedEpoch[iNode] = hintGen;
        wantOut(iNode, nActive, /*evenIfNotStale*/true);
        for (const Hit &h : topHitsLists[iNode].hits) {
            int64_t j = activeAncestor(h.j);
            if (j < 0 || j == iNode) continue;
            if (j != h.j) wantPair(iNode, j);
            wantOut(j, nActive);
        }
    }

    void hintVisible(int64_t nActive, int64_t iNode) {
        if (iNode < 0 || parent[iNode] >= 0) return;
        hintCriterion(nActive, iNode, visible[iNode].j);
    }

    // updateBestHit, NJ.tcc:1626-1648
    bool updateBestHit(int64_t nActive, Besthit &hit, bool bUpdateDist) {
        int64_t i = activeAncestor(hit.i), j = activeAncestor(hit.j);
        if (i < 0 || j < 0 || i == j) {
            hit.i = -1; hit.j = -1; hit.weight = 0; hit.dist = (P) 1e20; hit.criterion = (P) 1e20;
            return false;
        }
        if (i != hit.i || j != hit.j) {
            hit.i = i; hit.j = j;
            if (bUpdateDist) setDistCriterion(nActive, hit);
            else { hit.dist = (P) -1e20; hit.criterion = (P) 1e20; }
        }
        return true;
    }

    static void sortByCriterion(std::vector<Besthit> &v) {     // psort(.., CompareHitsByCriterion), NJ.tcc:7301-7306
        rsortByKey(v, [](const Besthit &a) { return orderKey(a.criterion); });
    }

    void sortSaveBestHits(int64_t iNode, std::vector<Besthit> &besthits, int64_t nIn, int64_t nOut, bool sort);
    void transferBestHits(int64_t nActive, int64_t iNode, const std::vector<Besthit> &oldhits, int64_t nOldHits,
                          Besthit *newhits, bool updateDistances);
    void uniqueBestHitsPrepare(int64_t nActive, std::vector<Besthit> &combined, std::vector<Besthit> &out,
                               std::vector<int64_t> &slots);
    void uniqueBestHitsFinish(int64_t nActive, std::vector<Besthit> &out, const std::vector<int64_t> &slots);
    void uniqueCore(int64_t nActive, std::vector<Besthit> &combined, std::vector<Besthit> &out);
    int hostThreads = 1;
    void setAllLeafTopHits();
    void resetTopVisible(int64_t nActive);
    void updateTopVisible(int64_t nActive, int64_t iIn, const Hit &hit);
    void updateVisible(int64_t nActive, std::vector<Besthit> &tophitsNode);
    void topHitNJSearch(int64_t nActive, Besthit &join);
    void speculateSearch(int64_t nActive);
    void getBestFromTopHits(int64_t iNode, int64_t nActive, Besthit &bestjoin);
    void topHitJoin(int64_t newnode, int64_t nActive);
    void refreshListsOnDevice(int64_t newnode, int64_t nActive, const std::vector<Besthit> &allhits);
    void refreshLists(int64_t newnode, int64_t nActive);
    void repairVisible(int64_t nActive);
    bool runDeviceLoop();
    std::vector<Besthit> thjCombined, thjUnique;          // scratch of topHitJoin
    std::vector<int64_t> thjSlots;
    // scratch of resetTopVisible
    std::vector<int64_t> rtvTouched; int64_t rtvStamp = 0;
    std::vector<std::vector<id_t>> rtvCandT, rtvWantT;
    std::vector<id_t> rtvCand;
    std::vector<Besthit> rtvVis;
    std::vector<std::pair<uint64_t, uint32_t>> rtvKv;
    std::vector<int64_t> rfNodes, rfOffset, rfOwnJ, rfAllJ, rfCount, rfOutJ;      // scratch of refreshListsOnDevice
    std::vector<P> rfOwnDist, rfAllDist, rfOutDist;
    // ---- speculative next join (vft_spec_join_*) ---------------------------------------------------------------
    // While the host finishes join t (list bookkeeping, the next search), the device already computes join t+1 as
    // guessed by speculateSearch -- the new profile, the updated out-profile (shadow copy) and the raw distances the
    // next topHitJoin will ask for.  If the search then picks that very pair (~95 %), the values are finished here in
    // the reference's arithmetic (diameter correction NJ.tcc:1120, setOutDistance NJ.tcc:1046-1052) and injected
    // into the pair / out-distance services, exactly as if they had been fetched; otherwise they are dropped.  Like
    // every prefetch in this file it is only a hint: the decisions never depend on it.
    struct Spec {
        bool valid = false;
        int64_t g = -1, gj = -1, out = -1, nActiveNext = 0, refreshes = 0;
        std::vector<int64_t> pairJ, outIds;
    } spec;
    bool specEnabled = false;
    int64_t guessG = -1, guessGj = -1;                    // the last guess of speculateSearch
    bool searchHinted = false;                            // topHitJoin already ran speculateSearch for the coming search
    int64_t nActiveOutProfileReset = 0;
    std::vector<P> selfdistH, selfweightH;                // host mirror of the device's self distances (needed to finish raw out-distances)
    std::vector<char> selfKnown;
    std::vector<int64_t> specSeenP, specSeenO; int64_t specStamp = 0;
    std::vector<P> specPD, specPW, specOD, specOW;
    void specLaunch(int64_t nActiveNext);
    bool specTake(int64_t i, int64_t j, int64_t newnode, int64_t nActive);
    void specInject(int64_t newnode, int64_t nActive);
    int64_t oneVsAll(int64_t query, int64_t nActive, int64_t K, std::vector<Besthit> &out);
    void fastNJSearch(int64_t nActive, std::vector<Besthit> &besthits, Besthit &join);
    void setBestHitFull(int64_t node, int64_t nActive, Besthit &bestjoin, std::vector<Besthit> *allhits);
};

// setBestHit + psort + cut, through the device (NJ.tcc:3571-3639, :3930, :4471).  Returns the
// number of *active* entries; entries [n, K) are the inactive sentinels in psort order.
template<typename P>
int64_t NJ<P>::oneVsAll(int64_t query, int64_t nActive, int64_t K, std::vector<Besthit> &out) {
    HPROF(10, "oneVsAll(host part)");
    std::vector<int64_t> js(K);
    std::vector<P> d(K), w(K), c(K);
    int64_t n = 0;
    check(timed([&] { return vft_dist_one_vs_all(ctx, query, nActive, K, js.data(), d.data(), w.data(), c.data(), &n); }));
    out.resize(K);
    for (int64_t k = 0; k < n; k++) out[k] = Besthit{(id_t) query, (id_t) js[k], w[k], d[k], c[k]};
    // sentinels of inactive nodes (NJ.tcc:3613-3617): criterion 1e20 ties, later index first
    int64_t k = n;
    for (int64_t j = maxnode - 1; j >= 0 && k < K; j--)
        if (parent[j] >= 0) out[k++] = Besthit{-1, (id_t) j, 0, (P) 1e20, (P) 1e20};
    out.resize(k);
    return n;
}

// sortSaveBestHits, NJ.tcc:4535-4578
template<typename P>
void NJ<P>::sortSaveBestHits(int64_t iNode, std::vector<Besthit> &besthits, int64_t nIn, int64_t nOut, bool sort) {
    HPROF(8, "sortSaveBestHits");
    if (sort) sortByCriterion(besthits);
    if (nIn > (int64_t) besthits.size()) nIn = (int64_t) besthits.size();
    int64_t nSave = 0, jLast = -1;
    for (int64_t k = 0; k < nIn && nSave < nOut; k++) {
        if (besthits[k].i < 0) continue;
        int64_t j = besthits[k].j;
        if (j != iNode && j != jLast && j >= 0) { nSave++; jLast = j; }
    }
    TopHitsList &l = topHitsLists[iNode];
    hintedEpoch[iNode] = -1;                         // the list changes: hints issued for the old one do not cover it
    l.hits.resize(nSave);
    int64_t iSave = 0;
    jLast = -1;
    for (int64_t k = 0; k < nIn && iSave < nSave; k++) {
        int64_t j = besthits[k].j;
        if (j != iNode && j != jLast && j >= 0) {
            l.hits[iSave].j = j; l.hits[iSave].dist = besthits[k].dist;
            iSave++; jLast = j;
        }
    }
}

// transferBestHits, NJ.tcc:4580-4613.  Distances come from the pair service (prefetched by caller).
template<typename P>
void NJ<P>::transferBestHits(int64_t nActive, int64_t iNode, const std::vector<Besthit> &oldhits, int64_t nOldHits,
                             Besthit *newhits, bool updateDistances) {
    for (int64_t k = 0; k < nOldHits; k++) {
        const Besthit &oldhit = oldhits[k];
        Besthit &nh = newhits[k];
        nh.i = iNode;
        nh.j = activeAncestor(oldhit.j);
        nh.dist = oldhit.dist; nh.weight = oldhit.weight; nh.criterion = oldhit.criterion;
        if (nh.j < 0 || nh.j == iNode) {
            nh.weight = 0; nh.dist = (P) -1e20; nh.criterion = (P) 1e20;
        } else if (nh.i != oldhit.i || nh.j != oldhit.j) {
            if (updateDistances) setDistCriterion(nActive, nh);
            else { nh.dist = (P) -1e20; nh.criterion = (P) 1e20; }
        } else {
            if (updateDistances) setCriterion(nActive, nh);
            else nh.criterion = (P) 1e20;
        }
    }
}

// uniqueBestHits, NJ.tcc:4786-4833, split around the device call:
//   prepare = ancestor walk + psort by (i,j) + dedupe; finish = distances + criteria
template<typename P>
void NJ<P>::uniqueBestHitsPrepare(int64_t nActive, std::vector<Besthit> &combined, std::vector<Besthit> &out,
                                  std::vector<int64_t> &slots) {
    HPROF(6, "uniqueBestHitsPrepare");
    uniqueCore(nActive, combined, out);
    slots.assign(out.size(), -1);
    for (size_t k = 0; k < out.size(); k++) {
        const Besthit &h = out[k];
        if (h.dist < 0.0) {
            const DW *have = pairCache.find(pkey(h.i, h.j));                 // a speculated join left it there (specInject)
            slots[k] = (have != nullptr && have->weight >= 0) ? -2 : slotPair(h.i, h.j);
        }
        hintCriterion(nActive, h.i, h.j);
    }
}

// the part of uniqueBestHits that touches no shared state: ancestor walk, psort by (i,j), dedupe
template<typename P>
void NJ<P>::uniqueCore(int64_t nActive, std::vector<Besthit> &combined, std::vector<Besthit> &out) {
    for (auto &h : combined) updateBestHit(nActive, h, false);
    {   // psort by (i,j), NJ.tcc:4797, :7309-7311; ids fit 31 bits, -1 sorts first
        // every valid entry of one list usually has the same i (the list's node): the (i,j) order is then the j order
        int64_t firstI = -1;
        bool uniformI = true;
        for (const Besthit &h : combined)
            if (h.i >= 0 && h.j >= 0) { if (firstI < 0) firstI = h.i; else if (h.i != firstI) { uniformI = false; break; } }
        if (uniformI) rsortByKey(combined, [](const Besthit &a) { return (a.i < 0 || a.j < 0) ? (uint64_t) 0 : (uint64_t) a.j + 1; });
        else rsortByKey(combined, [](const Besthit &a) { return ((uint64_t) (uint32_t) ((int64_t) a.i + 1) << 32) | (uint32_t) ((int64_t) a.j + 1); });
    }
    out.clear();
    out.reserve(combined.size());
    int64_t iSavedLast = -1;
    for (int64_t k = 0; k < (int64_t) combined.size(); k++) {
        const Besthit &hit = combined[k];
        if (hit.i < 0 || hit.j < 0) continue;
        if (iSavedLast >= 0) {
            const Besthit &saved = combined[iSavedLast];
            if (saved.i == hit.i && saved.j == hit.j) continue;
        }
        out.push_back(hit);
        iSavedLast = k;
    }
}

template<typename P>
void NJ<P>::uniqueBestHitsFinish(int64_t nActive, std::vector<Besthit> &out, const std::vector<int64_t> &slots) {
    HPROF(7, "uniqueBestHitsFinish");
    for (size_t k = 0; k < out.size(); k++) {
        Besthit &h = out[k];
        if (h.dist < 0.0) {                                      // :4826-4827
            if (slots[k] >= 0) { h.dist = reqD[slots[k]]; h.weight = reqW[slots[k]]; res->nPairPrefetchHit++; setCriterion(nActive, h); }
            else setDistCriterion(nActive, h);                           // -2: from the pair cache; -1: fetched on demand
        } else setCriterion(nActive, h);
    }
}

// setAllLeafTopHits, NJ.tcc:3746-4124 (the `-threads 1` branch, :3884-4015, without 2nd-level lists)
template<typename P>
void NJ<P>::setAllLeafTopHits() {
    double close = opt.tophitsClose;
    if (close < 0) {
        double logN = std::log((double) nSeqs) / std::log(2.0);
        close = logN / (logN + 2.0);
    }
    const std::vector<int64_t> &nGaps = leafGaps;                        // :3759-3763
    std::vector<int64_t> seeds(nSeqs);
    for (int64_t i = 0; i < nSeqs; i++) seeds[i] = i;
    // psort(seeds, CompareSeeds), NJ.tcc:3770, :7285-7299
    rsort(seeds, [&](int64_t a, int64_t b) {
        if (nGaps[a] != nGaps[b]) return nGaps[a] < nGaps[b];
        return outDistances[a] < outDistances[b];
    });

    std::vector<char> visited(nSeqs, 0);
    std::vector<Besthit> besthitsSeed, besthitsNeighbor(2 * m);
    std::vector<int64_t> closeNodes;
    for (int64_t iSeed = 0; iSeed < nSeqs; iSeed++) {
        int64_t seed = seeds[iSeed];
        if (visited[seed]) continue;
        visited[seed] = 1;
        res->nSeeds++;
        oneVsAll(seed, nSeqs, 2 * m, besthitsSeed);                      // setBestHit + sort, :3927-3930
        sortSaveBestHits(seed, besthitsSeed, (int64_t) besthitsSeed.size(), m, /*sort*/false);

        double neardist = besthitsSeed[2 * m - 1].dist * close;          // :3934
        double nearweight = 0;
        for (int64_t k = 0; k < 2 * m; k++) nearweight += besthitsSeed[k].weight;
        nearweight = nearweight / (2.0 * m);
        nearweight *= (1.0 - 2.0 * neardist / 3.0);
        double nearcover = 1.0 - neardist / 2.0;

        // which of the top m are close neighbours (:3953-3966); each j occurs once, so the set is
        // known before any transfer is done and all their 2m-candidate lists go in ONE device call
        closeNodes.clear();
        for (int64_t iClose = 0; iClose < m; iClose++) {
            const Besthit &closehit = besthitsSeed[iClose];
            int64_t closeNode = closehit.j;
            if (visited[closeNode]) continue;
            bool isClose = closehit.dist <= neardist
                           && (closehit.weight >= nearweight || closehit.weight >= (nPos - nGaps[closeNode]) * nearcover);
            bool identical = closehit.dist < 1e-6
                             && std::fabs(closehit.weight - (nPos - nGaps[seed])) < 1e-5
                             && std::fabs(closehit.weight - (nPos - nGaps[closeNode])) < 1e-5;
            if (isClose || identical) { closeNodes.push_back(closeNode); visited[closeNode] = 1; }
        }
        if (opt.prefetch) {
            // all 2m-candidate lists of this seed's close neighbours in ONE device call, results by slot
            for (int64_t closeNode : closeNodes)
                for (int64_t k = 0; k < 2 * m; k++) {
                    int64_t j = besthitsSeed[k].j;
                    if (j >= 0 && j != closeNode) slotPair(closeNode, j);
                }
            flush(nSeqs);
            // slot offsets per close node, then the lists are completed by the host threads
            std::vector<size_t> firstSlot(closeNodes.size() + 1, 0);
            for (size_t ci = 0; ci < closeNodes.size(); ci++) {
                size_t cnt = 0;
                for (int64_t k = 0; k < 2 * m; k++) { int64_t j = besthitsSeed[k].j; cnt += (j >= 0 && j != closeNodes[ci]); }
                firstSlot[ci + 1] = firstSlot[ci] + cnt;
            }
            res->nCloseUsed += (int64_t) closeNodes.size();
            res->nPairPrefetchHit += (int64_t) firstSlot.back();
#pragma omp parallel for schedule(dynamic, 4) num_threads(hostThreads)
            for (int64_t ci = 0; ci < (int64_t) closeNodes.size(); ci++) {
                const int64_t closeNode = closeNodes[ci];
                std::vector<Besthit> nb(2 * m);
                size_t r = firstSlot[ci];
                for (int64_t k = 0; k < 2 * m; k++) {                    // transferBestHits, :4585-4612
                    const Besthit &oldhit = besthitsSeed[k];
                    Besthit &nh = nb[k];
                    nh.i = closeNode; nh.j = oldhit.j;                   // every leaf is its own active ancestor
                    if (nh.j < 0 || nh.j == closeNode) { nh.weight = 0; nh.dist = (P) -1e20; nh.criterion = (P) 1e20; }
                    else { nh.dist = reqD[r]; nh.weight = reqW[r]; r++; setCriterion(nSeqs, nh); }
                }
                sortSaveBestHits(closeNode, nb, 2 * m, m, true);                                      // :3991
            }
        } else {
            for (int64_t closeNode : closeNodes) {
                res->nCloseUsed++;
                besthitsNeighbor.resize(2 * m);
                transferBestHits(nSeqs, closeNode, besthitsSeed, 2 * m, besthitsNeighbor.data(), true);   // :3988
                sortSaveBestHits(closeNode, besthitsNeighbor, 2 * m, m, true);                            // :3991
            }
        }
        clearPairs();
    }

    for (int64_t i = 0; i < nSeqs; i++) visible[i] = topHitsLists[i].hits[0];       // :4037-4044

    // checking phase, NJ.tcc:4052-4119 (criteria only; every out-distance is fresh here)
    int64_t nCheck = (int64_t) (0.5 + 2.0 * std::sqrt((double) m));
    for (int64_t iNode = 0; iNode < nSeqs; iNode++) {
        TopHitsList &lNode = topHitsLists[iNode];
        for (int64_t iHit = 0; iHit < nCheck && iHit < (int64_t) lNode.hits.size(); iHit++) {
            Besthit bh;
            hitToBestHit(iNode, lNode.hits[iHit], bh);
            setCriterion(nSeqs, bh);
            TopHitsList &lTarget = topHitsLists[bh.j];
            Besthit bhCheck;
            hitToBestHit(bh.j, lTarget.hits[nCheck - 1], bhCheck);
            setCriterion(nSeqs, bhCheck);
            if (bhCheck.criterion < bh.criterion) continue;
            bool bFound = false;
            for (size_t k = 0; k < lTarget.hits.size() && !bFound; k++)
                if (lTarget.hits[k].j == iNode) bFound = true;
            if (bFound) continue;
            int64_t iWorst = -1;
            double dWorstCriterion = -1e20;
            for (int64_t k = 0; k < (int64_t) lTarget.hits.size(); k++) {
                Besthit bh2;
                hitToBestHit(bh.j, lTarget.hits[k], bh2);
                setCriterion(nSeqs, bh2);
                if (bh2.criterion > dWorstCriterion) { iWorst = k; dWorstCriterion = bh2.criterion; }
            }
            if (dWorstCriterion > bh.criterion) {
                lTarget.hits[iWorst].j = iNode;
                lTarget.hits[iWorst].dist = bh.dist;
                Besthit v;
                getVisible(nSeqs, bh.j, v);
                if (bh.criterion < v.criterion) visible[bh.j] = lTarget.hits[iWorst];
            }
        }
    }
}

// resetTopVisible, NJ.tcc:4728-4784.  O(nActive) work, called every m/2 joins and after every refresh: the scans
// run on the host threads.  The reference evaluates getVisible() for every live node in ascending order; the
// only side effect of that loop is the lazy out-distance refresh of the nodes it touches (setCriterion,
// :1092-1098), which does not depend on the order -- so the touched set is marked first, its stale members are
// fetched in one device call and committed, and the criteria are then computed from read-only state.
template<typename P>
void NJ<P>::resetTopVisible(int64_t nActive) {
    HPROF(1, "resetTopVisible");
    searchHinted = false;                              // a new top-visible set: the next search hints for itself
    const int nT = maxnode >= 4096 ? hostThreads : 1;
    if ((int64_t) rtvTouched.size() < maxnodes) rtvTouched.assign(maxnodes, 0);
    const int64_t stamp = ++rtvStamp;
    std::vector<std::vector<id_t>> &candT = rtvCandT, &wantT = rtvWantT;
    candT.resize(nT); wantT.resize(nT);
    std::vector<id_t> &cand = rtvCand;
    std::vector<Besthit> &vis = rtvVis;
    std::vector<std::pair<uint64_t, uint32_t>> &kv = rtvKv;
    bool needSequential = false;
    size_t nVisible = 0, nAll = 0;
    const uint64_t zeroKey = orderKey((P) 0);
    // Device calls (and the exceptions they can raise) stay on the calling thread, outside the parallel regions:
    // the context is bound to that thread's CUDA device.
#pragma omp parallel num_threads(nT)
    {
        const int t = omp_get_thread_num(), nth = omp_get_num_threads();
        const int64_t lo = maxnode * t / nth, hi = maxnode * (t + 1) / nth;
        // pass 1: which nodes have a live visible hit (getVisible's tests, :546-553); mark what the loop touches
        std::vector<id_t> &mine = candT[t];
        mine.clear();
        for (int64_t i = lo; i < hi; i++) {
            if (parent[i] >= 0) continue;
            const Hit &h = visible[i];
            if (h.j < 0 || parent[h.j] >= 0) continue;
            mine.push_back((id_t) i);
            __atomic_store_n(&rtvTouched[i], stamp, __ATOMIC_RELAXED);
            __atomic_store_n(&rtvTouched[h.j], stamp, __ATOMIC_RELAXED);
        }
#pragma omp barrier
        // the touched nodes whose out-distance is stale and not in hand yet (owner computes: wantOut's state is per node)
        std::vector<id_t> &w = wantT[t];
        w.clear();
        if (opt.prefetch)
            for (int64_t i = lo; i < hi; i++)
                if (rtvTouched[i] == stamp && parent[i] < 0 && nOutDistActive[i] != nActive && freshEpoch[i] != epoch
                    && wantEpoch[i] != epoch && stale(i, nActive)) {
                    wantEpoch[i] = epoch;
                    w.push_back((id_t) i);
                }
    }
    cand.clear();
    for (int k = 0; k < (int) candT.size(); k++) cand.insert(cand.end(), candT[k].begin(), candT[k].end());   // ascending node order
    for (int k = 0; k < (int) wantT.size(); k++) for (id_t i : wantT[k]) wantIds.push_back(i);
    flush(nActive);
    nVisible = cand.size();
    nAll = (size_t) std::max<int64_t>(nActive, (int64_t) nVisible);
    vis.resize(nVisible);
    kv.resize(nAll);
#pragma omp parallel num_threads(nT)
    {
        const int t = omp_get_thread_num(), nth = omp_get_num_threads();
        const int64_t lo = maxnode * t / nth, hi = maxnode * (t + 1) / nth;
        // commit what setCriterion would refresh (:1092-1098)
        bool seq = false;
        for (int64_t i = lo; i < hi; i++)
            if (rtvTouched[i] == stamp && parent[i] < 0 && stale(i, nActive)) {
                if (freshEpoch[i] == epoch && epochActive == nActive) { outDistances[i] = freshVal[i]; nOutDistActive[i] = nActive; }
                else seq = true;
            }
        if (seq) {
#pragma omp atomic write
            needSequential = true;
        }
    }
    if (needSequential)                             // prefetching off (or a value not in hand): one at a time, as the reference would
        for (int64_t i = 0; i < maxnode; i++)
            if (rtvTouched[i] == stamp && parent[i] < 0 && stale(i, nActive)) setOutDistance(i, nActive);
#pragma omp parallel num_threads(nT)
    {
        // pass 2: criteria, in ascending node order = the order of visibleSorted[] in the reference (read-only now)
#pragma omp for schedule(static)
        for (int64_t k = 0; k < (int64_t) nVisible; k++) getVisible(nActive, cand[k], vis[k]);
        // The reference allocates nActive slots (:4729), fills nVisible of them and psorts ALL of them
        // (:4744); the unfilled tail is value-initialised (i=j=0, criterion 0) and takes part in the
        // sort, and only the first nVisible sorted slots are read (:4761).  Reproduced literally, but
        // only as much of the order as is consumed is materialised: (key, reversed position) pairs.
#pragma omp for schedule(static)
        for (int64_t k = 0; k < (int64_t) nAll; k++)
            kv[k] = {(size_t) k < nVisible ? orderKey(vis[k].criterion) : zeroKey, (uint32_t) (nAll - 1 - k)};
    }
    size_t sorted = std::min(nAll, 4 * topvisible.size() + 64);
    if (sorted < nAll) std::nth_element(kv.begin(), kv.begin() + sorted, kv.end());
    std::sort(kv.begin(), kv.begin() + sorted);
    static thread_local std::vector<int64_t> inTopVisible;
    static thread_local std::vector<int64_t> touched;
    if ((int64_t) inTopVisible.size() < maxnodes) inTopVisible.assign(maxnodes, -1);
    touched.clear();
    size_t iSave = 0;
    for (size_t k = 0; k < nVisible && iSave < topvisible.size(); k++) {
        if (k >= sorted) { std::sort(kv.begin() + sorted, kv.end()); sorted = nAll; }
        const size_t src = nAll - 1 - kv[k].second;
        const int64_t vi = src < nVisible ? vis[src].i : 0, vj = src < nVisible ? vis[src].j : 0;
        if (inTopVisible[vi] != vj) {
            topvisible[iSave++] = vi;
            inTopVisible[vi] = vj; touched.push_back(vi);
            inTopVisible[vj] = vi; touched.push_back(vj);
        }
    }
    for (int64_t t : touched) inTopVisible[t] = -1;
    while (iSave < topvisible.size()) topvisible[iSave++] = -1;
    topvisibleAge = 0;
}

// updateTopVisible, NJ.tcc:4661-4711
template<typename P>
void NJ<P>::updateTopVisible(int64_t nActive, int64_t iIn, const Hit &hit) {
    HPROF(2, "updateTopVisible");
    bool bIn = false;
    for (size_t k = 0; k < topvisible.size() && !bIn; k++) {
        int64_t iNode = topvisible[k];
        if (iNode == iIn) bIn = true;
        else if (iNode < 0 || parent[iNode] >= 0) { bIn = true; topvisible[k] = iIn; }
    }
    if (bIn) return;
    if (opt.prefetch) {
        for (int64_t iNode : topvisible) hintVisible(nActive, iNode);
        hintCriterion(nActive, iIn, hit.j);
        flushOut(nActive);
    }
    int64_t iPosWorst = -1;
    double dCriterionWorst = -1e20;
    for (size_t k = 0; k < topvisible.size() && !bIn; k++) {
        int64_t iNode = topvisible[k];
        Besthit v;
        if (!getVisible(nActive, iNode, v)) { topvisible[k] = iIn; bIn = true; }
        else if (v.i == hit.j && v.j == iIn) bIn = true;
        else if (v.criterion >= dCriterionWorst) { iPosWorst = (int64_t) k; dCriterionWorst = v.criterion; }
    }
    if (!bIn && iPosWorst >= 0) {
        Besthit v;
        hitToBestHit(iIn, hit, v);
        setCriterion(nActive, v);
        if (v.criterion < dCriterionWorst) topvisible[iPosWorst] = iIn;
    }
}

// updateVisible, NJ.tcc:4635-4658
template<typename P>
void NJ<P>::updateVisible(int64_t nActive, std::vector<Besthit> &tophitsNode) {
    HPROF(3, "updateVisible(incl updateTopVisible)");
    if (opt.prefetch) {
        for (const Besthit &hit : tophitsNode) if (hit.i >= 0) hintVisible(nActive, hit.j);
        flushOut(nActive);
    }
    for (Besthit &hit : tophitsNode) {
        if (hit.i < 0) continue;
        Besthit v;
        bool bSuccess = getVisible(nActive, hit.j, v);
        if (!bSuccess || hit.criterion < v.criterion) {
            if (bSuccess) res->nVisibleUpdate++;
            Hit &vis = visible[hit.j];
            vis.j = hit.i;
            vis.dist = hit.dist;
            updateTopVisible(nActive, hit.j, vis);
        }
    }
}

// getBestFromTopHits, NJ.tcc:4267-4298
template<typename P>
void NJ<P>::getBestFromTopHits(int64_t iNode, int64_t nActive, Besthit &bestjoin) {
    HPROF(4, "getBestFromTopHits");
    TopHitsList &l = topHitsLists[iNode];
    if (opt.prefetch) { hintList(nActive, iNode); flush(nActive); }
    setOutDistance(iNode, nActive);                                      // :4276
    bestjoin.i = -1; bestjoin.j = -1; bestjoin.weight = 0;
    bestjoin.dist = (P) 1e20; bestjoin.criterion = (P) 1e20;
    for (size_t k = 0; k < l.hits.size(); k++) {
        Besthit bh;
        hitToBestHit(iNode, l.hits[k], bh);
        if (updateBestHit(nActive, bh, true)) {
            setCriterion(nActive, bh);
            if (bh.criterion < bestjoin.criterion) bestjoin = bh;
        }
    }
}

// Hints for topHitNJSearch (queued, not flushed): the out-distances its scan over the top-visible set can refresh,
// and -- speculation -- what the hill-climb over the two top-hit lists of the likely join will ask for.  The join
// is guessed from the values already held (stale out-distances rescaled, no refresh).  Only a hint: the exact
// search decides, and fetches whatever the guess did not cover.  Called a second time from topHitJoin, whose one
// device call then usually covers the NEXT search as well (the next join almost never involves the node just
// created), so that a join costs one synchronous round trip instead of two.
template<typename P>
void NJ<P>::speculateSearch(int64_t nActive) {
    HPROF(15, "speculateSearch");
    // one guess is enough: measured on 8 000 taxa, guessing the best three instead of the best one saves 1.5 % of the
    // device calls and costs two more list scans per join.  One pass over the top-visible set: the out-distance hints
    // of hintVisible and the approximate criterion of the guess read the same entries.
    int64_t g1 = -1;
    double c1 = 1e300;
    const double invN2 = 1.0 / (double) (nActive - 2);          // (a hint: need not round like the reference's division)
    for (int64_t iNode : topvisible) {
        if (iNode < 0 || parent[iNode] >= 0) continue;
        const Hit &h = visible[iNode];
        if (h.j < 0 || parent[h.j] >= 0) continue;
        wantOut(iNode, nActive); wantOut(h.j, nActive);          // = hintVisible(nActive, iNode)
        double outI = outDistances[iNode], outJ = outDistances[h.j];
        if (nOutDistActive[iNode] != nActive) outI *= (nActive - 1) / (double) (nOutDistActive[iNode] - 1);
        if (nOutDistActive[h.j] != nActive) outJ *= (nActive - 1) / (double) (nOutDistActive[h.j] - 1);
        const double c = h.dist - (outI + outJ) * invN2;
        if (c < c1) { c1 = c; g1 = iNode; }
    }
    guessG = g1; guessGj = g1 >= 0 ? (int64_t) visible[g1].j : -1;
    for (int64_t g : {g1}) {
        if (g < 0) continue;
        const int64_t gj = visible[g].j;
        hintList(nActive, g); hintList(nActive, gj);
        wantPair(g, gj);
    }
}

// the visible-set repair of topHitNJSearch, NJ.tcc:4171-4201
template<typename P>
void NJ<P>::repairVisible(int64_t nActive) {
    for (int64_t iNode = 0; iNode < maxnode; iNode++) {
        if (parent[iNode] >= 0) continue;
        Hit &v = visible[iNode];
        int64_t newj = activeAncestor(v.j);
        if (newj >= 0 && newj != v.j) {
            if (newj == iNode) {
                newj = 0;
                while (parent[newj] >= 0 || newj == iNode) newj++;
            }
            Besthit bh = {(id_t) iNode, (id_t) newj, (P) -1e20, (P) -1e20, (P) -1e20};
            setDistCriterion(nActive, bh);
            v.j = newj;
            v.dist = bh.dist;
        }
    }
}

// topHitNJSearch, NJ.tcc:4137-4264
template<typename P>
void NJ<P>::topHitNJSearch(int64_t nActive, Besthit &join) {
    HPROF(5, "topHitNJSearch(total)");
    if (opt.prefetch) {
        // topHitJoin has usually hinted this very search already (same epoch, same top-visible set); what its list
        // bookkeeping changed since is fetched on demand by the scan below (counted in nOutSingleFetch)
        if (!searchHinted) speculateSearch(nActive);
        searchHinted = false;
        flush(nActive);
    }
    int64_t nCandidate = 0, iNodeBestCandidate = -1;
    double dBestCriterion = 1e20;
    HPROF(16, "topHitNJSearch(scan of the top-visible set)");
    for (size_t k = 0; k < topvisible.size(); k++) {
        int64_t iNode = topvisible[k];
        Besthit v;
        if (getVisible(nActive, iNode, v)) {
            nCandidate++;
            if (iNodeBestCandidate < 0 || v.criterion < dBestCriterion) {
                iNodeBestCandidate = iNode;
                dBestCriterion = v.criterion;
            }
        }
    }
    topvisibleAge++;
    if (2 * topvisibleAge > m
        || (3 * nCandidate < (int64_t) topvisible.size() && 3 * nCandidate < nActive)) {
        if (topvisibleAge <= 2) repairVisible(nActive);                  // :4171-4201
        resetTopVisible(nActive);
        topHitNJSearch(nActive, join);
        return;
    }
    getVisible(nActive, iNodeBestCandidate, join);                       // :4212

    // hill-climbing, NJ.tcc:4222-4263 (single thread: bests[] has one entry)
    bool changed;
    do {
        changed = false;
        Besthit best;
        getBestFromTopHits(join.i, nActive, best);
        if (best.j != join.j && best.criterion < join.criterion) { changed = true; join = best; }
        getBestFromTopHits(join.j, nActive, best);
        if (best.j != join.i && best.criterion < join.criterion) { changed = true; join = best; }
        if (changed) res->nHillBetter++;
    } while (changed);
}

// topHitJoin, NJ.tcc:4306-4533 (no 2nd-level lists: hitSource is always -1)
template<typename P>
void NJ<P>::topHitJoin(int64_t newnode, int64_t nActive) {
    HPROF(11, "topHitJoin(total)");
    TopHitsList &lNew = topHitsLists[newnode];
    TopHitsList *lChild[2] = {&topHitsLists[child[newnode].child[0]], &topHitsLists[child[newnode].child[1]]};
    size_t n0 = lChild[0]->hits.size();
    std::vector<Besthit> &combinedList = thjCombined, &uniqueList = thjUnique;      // scratch reused across joins
    std::vector<int64_t> &uniqueSlots = thjSlots;
    combinedList.resize(n0 + lChild[1]->hits.size());
    hitsToBestHits(lChild[0]->hits, child[newnode].child[0], combinedList.data());
    hitsToBestHits(lChild[1]->hits, child[newnode].child[1], combinedList.data() + n0);
    uniqueBestHitsPrepare(nActive, combinedList, uniqueList, uniqueSlots);
    if (opt.prefetch) {      // what updateTopVisible / updateVisible below can touch (a superset)
        for (const Besthit &h : uniqueList) hintVisible(nActive, h.j);
        speculateSearch(nActive); searchHinted = true;   // (covers the top-visible set)                    // ... and what the NEXT join search will most likely ask for
    }
    flush(nActive);
    if (specEnabled) specLaunch(nActive);           // the device starts on the NEXT join while the host finishes this one
    uniqueBestHitsFinish(nActive, uniqueList, uniqueSlots);
    int64_t nUnique = (int64_t) uniqueList.size();
    lChild[0]->hits.clear(); lChild[0]->hits.shrink_to_fit();
    lChild[1]->hits.clear(); lChild[1]->hits.shrink_to_fit();

    lNew.age = (lChild[0]->age + lChild[1]->age + 1) / 2 + 1;            // :4342
    int64_t tophitAgeLimit = std::max((int64_t) 1, (int64_t) (0.5 + std::log((double) m) / std::log(2.0)));
    bool bUseUnique = nUnique == nActive - 1
                      || (lNew.age <= tophitAgeLimit && nUnique >= (int64_t) (0.5 + m * opt.tophitsRefresh));
    if (bUseUnique) {
        int64_t nSave = std::min(nUnique, m);
        sortSaveBestHits(newnode, uniqueList, nUnique, nSave, true);     // :4432
        visible[newnode] = lNew.hits[0];
        updateTopVisible(nActive, newnode, visible[newnode]);
        uniqueList.resize(nSave);
        updateVisible(nActive, uniqueList);
        return;
    }

    refreshLists(newnode, nActive);
}

// the refresh branch of topHitJoin, NJ.tcc:4439-4517
template<typename P>
void NJ<P>::refreshLists(int64_t newnode, int64_t nActive) {
    TopHitsList &lNew = topHitsLists[newnode];
    Section secRefresh(this, 5);
    res->nRefreshTopHits++;
    if (spec.valid) { spec.valid = false; check(vft_spec_join_discard(ctx)); }      // a refresh rewrites the lists the guess was built from
    lNew.age = 0;
    {   // every out-distance up to date, :4451-4464
        std::vector<P> od(maxnodes);
        check(timed([&] { return vft_out_distance_all(ctx, nActive, totdiam, od.data(), maxnodes); }));
        for (int64_t i = 0; i < maxnode; i++)
            if (parent[i] < 0) { outDistances[i] = od[i]; nOutDistActive[i] = nActive; }
    }
    std::vector<Besthit> allhits;
    const int64_t nRealHits = oneVsAll(newnode, nActive, 2 * m, allhits);    // :4470-4471 (top 2m is all that is read)
    sortSaveBestHits(newnode, allhits, (int64_t) allhits.size(), m, false);   // :4472
    if (opt.prefetch && nRealHits >= 2 * m) {
        // The m list merges of :4477-4515 in ONE device pass (vft_tophits_merge): the host only resolves the
        // active ancestors of the stored hits (it owns the tree) and packs the lists.
        refreshListsOnDevice(newnode, nActive, allhits);
        clearPairs();
        resetTopVisible(nActive);                                        // :4517
        return;
    }

    // expand the lists of the top m hits, :4477-4515.  The m iterations are independent (they
    // read allhits and parent[], write only their own list), so their distance requests are
    // gathered first and evaluated in one device call.
    struct Work { int64_t iNode; std::vector<Besthit> unique; std::vector<int64_t> slots; };
    std::vector<Work> work;
    for (int64_t iHit = 0; iHit < m && iHit < (int64_t) allhits.size(); iHit++) {
        if (allhits[iHit].i < 0) continue;
        int64_t iNode = allhits[iHit].j;
        if (parent[iNode] >= 0) continue;
        work.push_back(Work{iNode, {}, {}});
    }
    const int64_t nAvail = std::min<int64_t>(2 * m, (int64_t) allhits.size());
    // host threads over the lists, as the reference does (NJ.tcc:4477): every out-distance is fresh
    // here, so setCriterion is a pure function and the iterations share nothing they write
#pragma omp parallel for schedule(dynamic, 2) num_threads(hostThreads)
    for (int64_t wi = 0; wi < (int64_t) work.size(); wi++) {
        Work &wk = work[wi];
        const int64_t iNode = wk.iNode;
        TopHitsList &l = topHitsLists[iNode];
        int64_t nHitsOld = (int64_t) l.hits.size();
        l.age = 0;
        std::vector<Besthit> bothList(nHitsOld + nAvail);
        hitsToBestHits(l.hits, iNode, bothList.data());
        for (int64_t k = 0; k < nHitsOld; k++) setCriterion(nActive, bothList[k]);
        transferBestHits(nActive, iNode, allhits, nAvail, bothList.data() + nHitsOld, false);
        uniqueCore(nActive, bothList, wk.unique);
    }
    for (Work &wk : work) {
        wk.slots.assign(wk.unique.size(), -1);
        for (size_t k = 0; k < wk.unique.size(); k++)
            if (wk.unique[k].dist < 0.0) wk.slots[k] = slotPair(wk.unique[k].i, wk.unique[k].j);
    }
    flush(nActive);
    int64_t hits = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(hostThreads) reduction(+ : hits)
    for (int64_t wi = 0; wi < (int64_t) work.size(); wi++) {
        Work &wk = work[wi];
        for (size_t k = 0; k < wk.unique.size(); k++) {            // uniqueBestHits tail, :4823-4831
            Besthit &h = wk.unique[k];
            if (h.dist < 0.0) {
                if (wk.slots[k] >= 0) { h.dist = reqD[wk.slots[k]]; h.weight = reqW[wk.slots[k]]; hits++; }
                else { DW r = pairDistNoCache(h.i, h.j); h.dist = r.dist; h.weight = r.weight; }
            }
            setCriterion(nActive, h);
        }
        sortSaveBestHits(wk.iNode, wk.unique, (int64_t) wk.unique.size(), m, true);   // :4512
        visible[wk.iNode] = topHitsLists[wk.iNode].hits[0];
    }
    if (threadError != VFT_OK) throw DeviceError{threadError};
    res->nPairPrefetchHit += hits;
    clearPairs();
    resetTopVisible(nActive);                                            // :4517
}

// The list-merging loop of a refresh (NJ.tcc:4477-4515) through vft_tophits_merge.  Every entry of allhits is a
// real, active hit here (the caller checked), so transferBestHits has no ancestor to resolve on that side.
template<typename P>
void NJ<P>::refreshListsOnDevice(int64_t newnode, int64_t nActive, const std::vector<Besthit> &allhits) {
    HPROF(9, "refreshListsOnDevice");
    // (plain locals, captured by reference into the host-thread regions below)
    std::vector<int64_t> &iNodes = rfNodes, &ownOffset = rfOffset, &ownJ = rfOwnJ, &allJ = rfAllJ, &outCount = rfCount, &outJ = rfOutJ;
    std::vector<P> &ownDist = rfOwnDist, &allDist = rfAllDist, &outDist = rfOutDist;
    iNodes.clear();
    for (int64_t iHit = 0; iHit < m; iHit++) {                            // :4477-4482
        if (allhits[iHit].i < 0) continue;
        const int64_t iNode = allhits[iHit].j;
        if (parent[iNode] >= 0) continue;
        iNodes.push_back(iNode);
    }
    const int64_t nLists = (int64_t) iNodes.size(), nAvail = 2 * m;
    ownOffset.assign(nLists + 1, 0);
    for (int64_t l = 0; l < nLists; l++) ownOffset[l + 1] = ownOffset[l] + (int64_t) topHitsLists[iNodes[l]].hits.size();
    ownJ.resize(ownOffset[nLists]); ownDist.resize(ownOffset[nLists]);
    allJ.resize(nAvail); allDist.resize(nAvail);
    for (int64_t k = 0; k < nAvail; k++) { allJ[k] = allhits[k].j; allDist[k] = allhits[k].dist; }
#pragma omp parallel for schedule(static) num_threads(hostThreads)
    for (int64_t l = 0; l < nLists; l++) {
        const std::vector<Hit> &hits = topHitsLists[iNodes[l]].hits;
        int64_t o = ownOffset[l];
        for (const Hit &h : hits) {                                      // updateBestHit(.., false), :1626-1648
            const int64_t j = activeAncestor(h.j);
            ownJ[o] = j;
            ownDist[o] = j == h.j ? h.dist : (P) -1e20;
            o++;
        }
    }
    outCount.resize(nLists); outJ.resize(nLists * m); outDist.resize(nLists * m);
    check(timed([&] { return vft_tophits_merge(ctx, newnode, nActive, m, nLists, iNodes.data(), ownOffset.data(), ownJ.data(),
                                               ownDist.data(), nAvail, allJ.data(), allDist.data(), outCount.data(), outJ.data(),
                                               outDist.data()); }));
#pragma omp parallel for schedule(static) num_threads(hostThreads)
    for (int64_t l = 0; l < nLists; l++) {
        TopHitsList &lst = topHitsLists[iNodes[l]];
        lst.age = 0;
        hintedEpoch[iNodes[l]] = -1;
        lst.hits.resize(outCount[l]);
        for (int64_t k = 0; k < outCount[l]; k++) { lst.hits[k].j = (id_t) outJ[l * m + k]; lst.hits[k].dist = outDist[l * m + k]; }
        visible[iNodes[l]] = lst.hits[0];                                // :4513
    }
}

// setBestHit with the full allhits array (visible-set mode only, N tiny): NJ.tcc:3571-3639
template<typename P>
void NJ<P>::setBestHitFull(int64_t node, int64_t nActive, Besthit &bestjoin, std::vector<Besthit> *allhits) {
    std::vector<Besthit> sorted;
    int64_t n = oneVsAll(node, nActive, maxnode, sorted);
    bestjoin = Besthit{(id_t) node, -1, 0, (P) 1e20, (P) 1e20};
    // arg-min with strict '<' scanning j ascending (:3627): lowest j among equal criteria
    for (int64_t k = 0; k < n; k++) {
        const Besthit &h = sorted[k];
        if (h.j == node) continue;
        if (h.criterion < bestjoin.criterion || (h.criterion == bestjoin.criterion && bestjoin.j >= 0 && h.j < bestjoin.j))
            bestjoin = h;
    }
    if (allhits) {
        allhits->assign(maxnode, Besthit{-1, -1, 0, (P) 1e20, (P) 1e20});
        for (int64_t j = 0; j < maxnode; j++) (*allhits)[j].j = j;
        for (int64_t k = 0; k < n; k++) (*allhits)[sorted[k].j] = sorted[k];
    }
}

// fastNJSearch, NJ.tcc:3686-3744 (visible-set mode)
template<typename P>
void NJ<P>::fastNJSearch(int64_t nActive, std::vector<Besthit> &besthits, Besthit &join) {
    join = Besthit{-1, -1, 0, (P) 1e20, (P) 1e20};
    for (int64_t iNode = 0; iNode < maxnode; iNode++) {
        int64_t jNode = besthits[iNode].j;
        if (parent[iNode] < 0 && parent[jNode] < 0) {
            setCriterion(nActive, besthits[iNode]);
            if (besthits[iNode].criterion < join.criterion) join = besthits[iNode];
        }
    }
    int changed;
    do {
        changed = 0;
        setBestHitFull(join.i, nActive, besthits[join.i], nullptr);
        if (besthits[join.i].j != join.j) changed = 1;
        join.j = besthits[join.i].j; join.weight = besthits[join.i].weight;
        join.dist = besthits[join.i].dist; join.criterion = besthits[join.i].criterion;
        setBestHitFull(join.j, nActive, besthits[join.j], nullptr);
        if (besthits[join.j].j != join.i) {
            changed = 1;
            join.i = besthits[join.j].j; join.weight = besthits[join.j].weight;
            join.dist = besthits[join.j].dist; join.criterion = besthits[join.j].criterion;
        }
        if (changed) res->nHillBetter++;
    } while (changed);
}

// Launch the guessed next join (g, gj) -> row maxnode, with the distances topHitJoin will ask for: the new node against
// the active ancestors of both children's top hits (uniqueBestHits, NJ.tcc:4786-4833: every one of them is recomputed,
// the parent changed), and the out-distances setCriterion / updateVisible / updateTopVisible can refresh at the next
// nActive (the new node's own, the candidates', their visible partners', the top-visible set's).  A superset is harmless.
template<typename P>
void NJ<P>::specLaunch(int64_t nActiveNext) {
    spec.valid = false;
    const int64_t g = guessG, gj = guessGj;
    if (g < 0 || gj < 0 || g == gj || parent[g] >= 0 || parent[gj] >= 0 || nActiveNext <= 3 || maxnode >= maxnodes - 1) return;
    {   // the next join must not be one that rebuilds the out-profile from scratch (NJ.tcc:3012-3033)
        const int64_t changed = nActiveOutProfileReset - (nActiveNext - 1);
        if (changed >= opt.nResetOutProfile && changed >= opt.fResetOutProfile * nActiveOutProfileReset) return;
    }
    const int64_t R = maxnode, nAfter = nActiveNext - 1;
    const int64_t stampP = ++specStamp;
    spec.pairJ.clear(); spec.outIds.clear();
    auto wantOutNext = [&](int64_t i) {
        if (i < 0 || parent[i] >= 0 || i == g || i == gj || specSeenO[i] == stampP) return;
        specSeenO[i] = stampP;
        if (!selfKnown[i]) return;
        int64_t nDiffAllow = opt.tophitsMult > 0 ? (int64_t) (nAfter * opt.staleOutLimit) : 0;
        if (nOutDistActive[i] - nAfter > nDiffAllow) spec.outIds.push_back(i);
    };
    spec.outIds.push_back(R);
    for (int64_t c : {g, gj})
        for (const Hit &h : topHitsLists[c].hits) {
            const int64_t j = activeAncestor(h.j);
            if (j < 0 || j == g || j == gj || specSeenP[j] == stampP) continue;
            specSeenP[j] = stampP;
            spec.pairJ.push_back(j);
            wantOutNext(j);
            wantOutNext(visible[j].j);
        }
    for (int64_t iNode : topvisible) {
        if (iNode < 0 || parent[iNode] >= 0) continue;
        wantOutNext(iNode);
        wantOutNext(visible[iNode].j);
    }
    if (spec.pairJ.empty() || spec.pairJ.size() + spec.outIds.size() > 4096) return;
    check(timed([&] { return vft_spec_join_launch(ctx, R, std::min(g, gj), std::max(g, gj), opt.bionj ? 0.5 : -1.0, nActiveNext, spec.pairJ.data(),
                                                  (int64_t) spec.pairJ.size(), spec.outIds.data(), (int64_t) spec.outIds.size()); }));
    spec.valid = true; spec.g = g; spec.gj = gj; spec.out = R; spec.nActiveNext = nActiveNext; spec.refreshes = res->nRefreshTopHits;
}

// The search picked (i, j) at nActive: if that is the speculated pair, commit it on the device and fetch the raw results
template<typename P>
bool NJ<P>::specTake(int64_t i, int64_t j, int64_t newnode, int64_t nActive) {
    if (!((i == spec.g && j == spec.gj) || (i == spec.gj && j == spec.g)) || newnode != spec.out || nActive != spec.nActiveNext) return false;
    specPD.resize(spec.pairJ.size()); specPW.resize(spec.pairJ.size());
    specOD.resize(spec.outIds.size()); specOW.resize(spec.outIds.size());
    P self2[2];
    check(timed([&] { return vft_spec_join_take(ctx, (double) diameter[newnode], specPD.data(), specPW.data(), specOD.data(), specOW.data(), self2); }));
    selfdistH[newnode] = self2[0]; selfweightH[newnode] = self2[1]; selfKnown[newnode] = 1;
    spec.valid = false;
    return true;
}

// After newEpoch(nActive): hand the speculated values to the pair and out-distance services of this epoch
template<typename P>
void NJ<P>::specInject(int64_t newnode, int64_t nActive) {
    for (size_t k = 0; k < spec.pairJ.size(); k++) {
        const int64_t j = spec.pairJ[k];
        P d = specPD[k] - (P) (diameter[newnode] + diameter[j]);                             // NJ.tcc:1120, numeric_t arithmetic
        d = (P) ((double) d + 0.0);                                                          // :1122 (no constraints)
        pairCache.put(pkey(newnode, j), DW{d, specPW[k]});
    }
    const P pN = (P) nActive, pN1 = (P) (nActive - 1);
    for (size_t k = 0; k < spec.outIds.size(); k++) {
        const int64_t i = spec.outIds[k];
        const P ddist = specOD[k], dweight = specOW[k];
        // setOutDistance's algebra, NJ.tcc:1046-1052, operand types as in the reference
        const P t1 = ddist * dweight;
        const P t2 = t1 * pN;
        const P t3 = selfweightH[i] * selfdistH[i];
        const P t4 = t2 - t3;
        const double top = (double) (P) (pN1 * t4);
        const P b1 = dweight * pN;
        const double bottom = (double) (P) (b1 - selfweightH[i]);
        const double pd = top / bottom;
        const P dn = diameter[i] * pN1;
        freshVal[i] = (P) (bottom > 0.01 ? pd - (double) dn - (totdiam - (double) diameter[i]) : 3.0);
        freshEpoch[i] = epoch;
    }
}

// The join loop on the device (nj_loop.h): the state built by the leaf phase goes to the loop backend, which runs joins
// until three nodes are left.  What the backend does not do itself comes back as a status: the image is downloaded, the
// host code above performs that one operation (the rebuild of the top-visible set, a top-hits refresh) exactly as the
// host-driven loop would, and the image goes back.  Returns false when the loop is not applicable (the caller then
// runs the host-driven loop).
template<typename P>
bool NJ<P>::runDeviceLoop() {
    if (m <= 0 || opt.bionj || !opt.deviceLoop) return false;
    const int64_t M = maxnodes, nTV = (int64_t) topvisible.size();
    vftx_loop *lp = nullptr;
    if (vftx_loop_create(ctx, &opt, m, nTV, &lp) != VFT_OK) return false;
    std::vector<int32_t> iParent(M), iUp(M), iChild(3 * M), iNOut(M), iHitJ((size_t) M * m), iHitCount(M), iAge(M), iVisJ(M), iTop(nTV);
    std::vector<P> iBl(M), iDiam(M), iOut(M), iHitD((size_t) M * m), iVisD(M);
    vftx_loop_image img{};
    img.nSeqs = nSeqs; img.maxnodes = M; img.m = m; img.nTV = nTV;
    img.parent = iParent.data(); img.up = iUp.data(); img.child = iChild.data(); img.branchlength = iBl.data(); img.diameter = iDiam.data();
    img.outDist = iOut.data(); img.nOutAct = iNOut.data(); img.hitJ = iHitJ.data(); img.hitDist = iHitD.data(); img.hitCount = iHitCount.data();
    img.age = iAge.data(); img.visJ = iVisJ.data(); img.visDist = iVisD.data(); img.topvisible = iTop.data();
    int64_t nActive = nSeqs;
    auto exportImage = [&] {
        for (int64_t i = 0; i < M; i++) {
            iParent[i] = (int32_t) parent[i]; iUp[i] = up[i];
            for (int k = 0; k < 3; k++) iChild[3 * i + k] = (int32_t) child[i].child[k];
            iBl[i] = branchlength[i]; iDiam[i] = diameter[i]; iOut[i] = outDistances[i]; iNOut[i] = nOutDistActive[i];
            const std::vector<Hit> &h = topHitsLists[i].hits;
            iHitCount[i] = (int32_t) h.size(); iAge[i] = (int32_t) topHitsLists[i].age;
            for (size_t k = 0; k < h.size(); k++) { iHitJ[(size_t) i * m + k] = h[k].j; iHitD[(size_t) i * m + k] = h[k].dist; }
            iVisJ[i] = visible[i].j; iVisD[i] = visible[i].dist;
        }
        for (int64_t k = 0; k < nTV; k++) iTop[k] = (int32_t) topvisible[k];
        img.maxnode = maxnode; img.nActive = nActive; img.topvisibleAge = topvisibleAge; img.nActiveOutProfileReset = nActiveOutProfileReset;
        img.totdiam = totdiam;
    };
    auto importImage = [&] {
        maxnode = img.maxnode; nActive = img.nActive; topvisibleAge = img.topvisibleAge; nActiveOutProfileReset = img.nActiveOutProfileReset;
        totdiam = img.totdiam;
        for (int64_t i = 0; i < M; i++) {
            parent[i] = iParent[i]; up[i] = iUp[i];
            int nc = 0;
            for (int k = 0; k < 3; k++) { child[i].child[k] = iChild[3 * i + k]; nc += iChild[3 * i + k] >= 0; }
            child[i].nChild = nc;
            branchlength[i] = iBl[i]; diameter[i] = iDiam[i]; outDistances[i] = iOut[i]; nOutDistActive[i] = iNOut[i];
            std::vector<Hit> &h = topHitsLists[i].hits;
            h.resize((size_t) iHitCount[i]); topHitsLists[i].age = iAge[i];
            for (size_t k = 0; k < h.size(); k++) { h[k].j = iHitJ[(size_t) i * m + k]; h[k].dist = iHitD[(size_t) i * m + k]; }
            hintedEpoch[i] = -1;
            visible[i].j = iVisJ[i]; visible[i].dist = iVisD[i];
        }
        for (int64_t k = 0; k < nTV; k++) topvisible[k] = iTop[k];
        newEpoch(nActive);
        searchHinted = false;
    };
    int rc = VFT_OK;
    int32_t resume = 0;
    vftx_loop_status stt{};
    for (;;) {
        exportImage();
        rc = vftx_loop_upload(lp, &img, resume);
        if (rc != VFT_OK) break;
        rc = timed([&] { return vftx_loop_run(lp, &stt); });
        res->nDeviceCalls++;
        if (rc != VFT_OK) break;
        rc = vftx_loop_download(lp, &img);
        if (rc != VFT_OK) break;
        importImage();
        if (stt.status == 1 /* DONE */) break;
        if (stt.status == 2 /* NEED_RESET */) {
            Section sec(this, 2);
            if (stt.visfixPending) repairVisible(nActive);
            resetTopVisible(nActive);
        } else if (stt.status == 3 /* NEED_REFRESH */) {
            Section sec(this, 4);
            refreshLists(stt.newnode, nActive);
        } else { rc = VFT_EINVAL; break; }
        resume = 0;
    }
    if (rc == VFT_OK) {
        res->nRefreshTopHits = stt.nRefresh;      // every refresh the loop asked for, wherever it was carried out
        res->nVisibleUpdate += stt.nVisibleUpdate; res->nHillBetter += stt.nHillBetter;
        res->nOutSingleFetch += stt.nInlineOut; res->nPairSingleFetch += stt.nInlinePair; res->nPairPrefetchHit += stt.nPairHit;
        if (res->joins) vftx_loop_joins(lp, res->joins, nSeqs - 3);
    }
    vftx_loop_destroy(lp);
    if (rc != VFT_OK) throw DeviceError{rc};
    return true;
}

// fastNJ, NJ.tcc:2796-3155
template<typename P>
void NJ<P>::fastNJ() {
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    if (nSeqs < 3) {                                                     // :2798-2815
        root = maxnode++;
        child[root].nChild = (int) nSeqs;
        for (int64_t i = 0; i < nSeqs; i++) { parent[i] = root; child[root].child[i] = i; }
        if (nSeqs == 2) {
            DW r = pairDist(0, 1);
            branchlength[0] = (P) (r.dist / 2.0);
            branchlength[1] = (P) (r.dist / 2.0);
        }
        return;
    }
    std::vector<Besthit> visibleSet, besthitNew;
    m = 0;
    if (opt.tophitsMult > 0) {
        m = (int64_t) (0.5 + opt.tophitsMult * std::sqrt((double) nSeqs));
        if (m < 4 || 2 * m >= nSeqs) m = 0;
    }
    res->m = m;
    if (m > 0) {
        topHitsLists.resize(maxnodes);
        visible.assign(maxnodes, Hit{-1, (P) 1e20});
        topvisible.assign((size_t) (0.5 + opt.topvisibleMult * m), -1);
        { Section sec(this, 0); setAllLeafTopHits(); }
        if (res->leafTopHits) {
            for (int64_t i = 0; i < nSeqs; i++)
                for (int64_t k = 0; k < m; k++)
                    res->leafTopHits[i * m + k] = k < (int64_t) topHitsLists[i].hits.size() ? topHitsLists[i].hits[k].j : -1;
        }
        { Section sec(this, 1); resetTopVisible(nSeqs); }
    } else {
        visibleSet.resize(maxnodes, Besthit{-1, -1, 0, (P) 1e20, (P) 1e20});
        for (int64_t i = 0; i < nSeqs; i++) setBestHitFull(i, nSeqs, visibleSet[i], nullptr);
    }
    auto t1 = clk::now();
    res->secondsLeafTopHits = std::chrono::duration<double>(t1 - t0).count();

    nActiveOutProfileReset = nSeqs;
    const bool onDevice = runDeviceLoop();
    for (int64_t nActive = nSeqs; !onDevice && nActive > 3; nActive--) {
        Besthit join;
        {
            Section sec(this, 2);
            if (m > 0) topHitNJSearch(nActive, join);
            else fastNJSearch(nActive, visibleSet, join);
        }
        Section secJoin(this, 3);

        // :2897-2900 -- out-distances of the pair up to date, distance and weight recomputed
        if (opt.prefetch) {
            wantOut(join.i, nActive, true); wantOut(join.j, nActive, true);
            wantPair(join.i, join.j);
            flushOut(nActive); flushPairs();
        }
        setOutDistance(join.i, nActive);
        setOutDistance(join.j, nActive);
        setDistCriterion(nActive, join);

        int64_t newnode = maxnode++;
        parent[join.i] = newnode;
        parent[join.j] = newnode;
        up[join.i] = newnode; up[join.j] = newnode;
        child[newnode].nChild = 2;
        child[newnode].child[0] = join.i < join.j ? join.i : join.j;
        child[newnode].child[1] = join.i > join.j ? join.i : join.j;
        if (res->joins) {
            res->joins[2 * (nSeqs - nActive)] = child[newnode].child[0];
            res->joins[2 * (nSeqs - nActive) + 1] = child[newnode].child[1];
        }

        // :2911-2916.  P + P + P is P arithmetic, then widened
        double rawIJ = (P) ((P) (join.dist + diameter[join.i]) + diameter[join.j]);
        double distIJ = join.dist;
        double deltaDist = (P) (outDistances[join.i] - outDistances[join.j]) / (double) (nActive - 2);
        branchlength[join.i] = (P) ((distIJ + deltaDist) / 2);
        branchlength[join.j] = (P) ((distIJ - deltaDist) / 2);

        double bionjWeight = 0.5;
        double varIJ = rawIJ - varDiameter[join.i] - varDiameter[join.j];
        if (opt.bionj && join.weight > 0.01 && varIJ > 0.001) {
            // BIONJ weighting, NJ.tcc:2921-2966 (Gascuel 1997, eq. 9, with the variances read off the out-profile):
            // bare profileDist of both children against the CURRENT out-profile, then scalar algebra in the
            // reference's operand types (numeric_t products, double quotients)
            const int64_t bi[2] = {join.i, join.j}, bo[2] = {-1, -1};
            P od[2], ow[2];
            check(timed([&] { return vft_dist_pairs(ctx, bi, bo, 2, VFT_PAIRS_PROFILE_RAW, od, ow); }));
            P sdw[2], sww[2];
            for (int k = 0; k < 2; k++) {
                if (!selfKnown[bi[k]]) {
                    double sd = 0, sw = 0;
                    check(timed([&] { return vft_get_self(ctx, bi[k], &sd, &sw); }));
                    selfdistH[bi[k]] = (P) sd; selfweightH[bi[k]] = (P) sw; selfKnown[bi[k]] = 1;
                }
                sdw[k] = selfdistH[bi[k]]; sww[k] = selfweightH[bi[k]];
            }
            const P pN = (P) nActive;
            const double varIWeight = (P) ((P) ((P) (pN * ow[0]) - sww[0]) - join.weight);
            const double varJWeight = (P) ((P) ((P) (pN * ow[1]) - sww[1]) - join.weight);
            const double varITop = (double) (P) ((P) ((P) (od[0] * ow[0]) * pN) - (P) (sdw[0] * sww[0])) - rawIJ * (double) join.weight;
            const double varJTop = (double) (P) ((P) ((P) (od[1] * ow[1]) * pN) - (P) (sdw[1] * sww[1])) - rawIJ * (double) join.weight;
            const double deltaProfileVarOut = (double) (nActive - 2) * (varJTop / varJWeight - varITop / varIWeight);
            const double deltaVarDiam = (P) ((P) (nActive - 2) * (P) (varDiameter[join.i] - varDiameter[join.j]));
            if (varJWeight > 0.01 && varIWeight > 0.01)
                bionjWeight = 0.5 + (deltaProfileVarOut + deltaVarDiam) / ((double) (2 * (nActive - 2)) * varIJ);
            if (bionjWeight < 0) bionjWeight = 0;
            if (bionjWeight > 1) bionjWeight = 1;
        }
        // :3003-3007: double * (P+P) ...
        diameter[newnode] = (P) (bionjWeight * (P) (branchlength[join.i] + diameter[join.i])
                                 + (1 - bionjWeight) * (P) (branchlength[join.j] + diameter[join.j]));
        varDiameter[newnode] = (P) (bionjWeight * varDiameter[join.i] + (1 - bionjWeight) * varDiameter[join.j]
                                    + bionjWeight * (1 - bionjWeight) * varIJ);
        // averageProfile (:3008, + the self-distance of :3040-3043) and the out-profile (:3012-3037)
        int64_t changedActiveOutProfile = nActiveOutProfileReset - (nActive - 1);
        const bool rebuildOut = changedActiveOutProfile >= opt.nResetOutProfile
                                && changedActiveOutProfile >= opt.fResetOutProfile * nActiveOutProfileReset;
        bool taken = false;
        if (spec.valid) {
            if (!rebuildOut) taken = specTake(join.i, join.j, newnode, nActive);
            if (!taken) { spec.valid = false; check(vft_spec_join_discard(ctx)); res->nSpecMiss++; }
        }
        if (taken) {
            totdiam += (P) ((P) (diameter[newnode] - diameter[join.i]) - diameter[join.j]);
            res->nSpecHit++;
        } else if (rebuildOut) {
            check(timed([&] { return vft_profile_average(ctx, newnode, join.i, join.j, opt.bionj ? bionjWeight : -1.0,
                                                         (double) diameter[newnode]); }));
            totdiam = 0;
            for (int64_t i = 0; i < maxnode; i++) if (parent[i] < 0) totdiam += diameter[i];
            check(timed([&] { return vft_outprofile_rebuild(ctx, nullptr, nActive - 1); }));
            nActiveOutProfileReset = nActive - 1;
        } else {
            check(timed([&] { return vft_profile_average_update(ctx, newnode, join.i, join.j, opt.bionj ? bionjWeight : -1.0,
                                                                (double) diameter[newnode], nActive); }));
            totdiam += (P) ((P) (diameter[newnode] - diameter[join.i]) - diameter[join.j]);
        }
        newEpoch(nActive - 1);
        if (taken) specInject(newnode, nActive - 1);
        else if (specEnabled) {                     // the self distance of a node joined the ordinary way: fetched (rare)
            double sd = 0, sw = 0;
            check(timed([&] { return vft_get_self(ctx, newnode, &sd, &sw); }));
            selfdistH[newnode] = (P) sd; selfweightH[newnode] = (P) sw; selfKnown[newnode] = 1;
        }

        if (m > 0) {
            Section sec(this, 4);
            topHitJoin(newnode, nActive - 1);
        } else {
            std::vector<P> od(maxnodes);
            check(timed([&] { return vft_out_distance_all(ctx, nActive - 1, totdiam, od.data(), maxnodes); }));
            for (int64_t i = 0; i < maxnode; i++)
                if (parent[i] < 0) { outDistances[i] = od[i]; nOutDistActive[i] = nActive - 1; }
            setBestHitFull(newnode, nActive - 1, visibleSet[newnode], &besthitNew);
            for (int64_t iNode = 0; iNode < maxnode; iNode++) {          // :3066-3095
                if (parent[iNode] >= 0 || iNode == newnode) continue;
                int64_t iOldVisible = visibleSet[iNode].j;
                if (parent[iOldVisible] < 0) setCriterion(nActive - 1, visibleSet[iNode]);
                if (parent[iOldVisible] >= 0 || besthitNew[iNode].criterion < visibleSet[iNode].criterion) {
                    if (parent[iOldVisible] < 0) res->nVisibleUpdate++;
                    visibleSet[iNode].j = newnode;
                    visibleSet[iNode].dist = besthitNew[iNode].dist;
                    visibleSet[iNode].criterion = besthitNew[iNode].criterion;
                }
            }
        }
    }

    // root the last three nodes, NJ.tcc:3107-3135
    int64_t top[3], nTop = 0;
    for (int64_t i = 0; i < maxnode; i++) if (parent[i] < 0) top[nTop++] = i;
    root = maxnode++;
    child[root].nChild = 3;
    for (nTop = 0; nTop < 3; nTop++) { parent[top[nTop]] = root; child[root].child[nTop] = top[nTop]; }
    // bare profileDist of the three pairs, then "dist - diameter - diameter" in P (:3125-3132)
    int64_t pi[3] = {top[0], top[0], top[1]}, pj[3] = {top[1], top[2], top[2]};
    P pd[3], pw[3];
    check(timed([&] { return vft_dist_pairs(ctx, pi, pj, 3, VFT_PAIRS_PROFILE_RAW, pd, pw); }));
    double d01 = (P) ((P) (pd[0] - diameter[top[0]]) - diameter[top[1]]);
    double d02 = (P) ((P) (pd[1] - diameter[top[0]]) - diameter[top[2]]);
    double d12 = (P) ((P) (pd[2] - diameter[top[1]]) - diameter[top[2]]);
    branchlength[top[0]] = (P) ((d01 + d02 - d12) / 2);
    branchlength[top[1]] = (P) ((d01 + d12 - d02) / 2);
    branchlength[top[2]] = (P) ((d02 + d12 - d01) / 2);
    res->secondsJoins = std::chrono::duration<double>(clk::now() - t1).count();
}

template<typename P>
int run(vft_ctx *ctx, const vft_config &cfg, const vft_nj_options &opt, const uint8_t *codes, vft_nj_result *res) {
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
#ifdef VFT_HOSTPROF
    g_hpCalls = &res->secondsInCalls; for (int k = 0; k < 32; k++) { g_hp[k] = 0; g_hpN[k] = 0; }
#endif
    NJ<P> nj(ctx, cfg, opt, res, codes);
    try {
        vft_timer_start(ctx);
        nj.init();
        nj.fastNJ();
        vft_timer_stop(ctx, &res->deviceMsResident);
    } catch (const DeviceError &e) {
        return e.code;
    }
    int64_t M = 2 * cfg.nSeqs;
    for (int64_t i = 0; i < M; i++) {
        res->parent[i] = nj.parent[i];
        res->nChild[i] = nj.child[i].nChild;
        for (int k = 0; k < 3; k++) res->child[3 * i + k] = nj.child[i].child[k];
        ((P *) res->branchlength)[i] = nj.branchlength[i];
    }
    res->root = nj.root;
    res->maxnode = nj.maxnode;
    res->secondsTotal = std::chrono::duration<double>(clk::now() - t0).count();
#ifdef VFT_HOSTPROF
    for (int k = 0; k < 32; k++) if (g_hpN[k]) std::fprintf(stderr, "[hostprof] %-40s %9ld calls %8.1f ms  %7.2f us/call\n", g_hpName[k], g_hpN[k], 1e3 * g_hp[k], 1e6 * g_hp[k] / g_hpN[k]);
#endif
    return VFT_OK;
}

}  // namespace

extern "C" void vft_nj_default_options(vft_nj_options *o) {
    std::memset(o, 0, sizeof *o);
    o->tophitsMult = 1.0;        // Options.h:20
    o->tophitsClose = -1.0;      // Options.h:22
    o->topvisibleMult = 1.5;     // Options.h:25
    o->tophitsRefresh = 0.8;     // Options.h:28
    o->staleOutLimit = 0.01;     // Options.h:36
    o->fResetOutProfile = 0.02;  // Options.h:38
    o->nResetOutProfile = 200;   // Options.h:40
    o->bionj = 0;
    o->prefetch = 1;
    o->hostThreads = 0;
    o->deviceLoop = 0;
    if (const char *e = std::getenv("VFT_DEVICE_LOOP")) o->deviceLoop = std::atoi(e);
}

extern "C" int vft_nj_build(const vft_config *cfg, const vft_nj_options *opt_in, const uint8_t *codes,
                            const void *const tables[4], vft_nj_result *res) {
    auto bad = [](const char *msg) { vftx_set_error(msg); return VFT_EINVAL; };
    if (!cfg || !codes || !res || !res->parent || !res->nChild || !res->child || !res->branchlength) return bad("vft_nj_build: null argument");
    if (cfg->nSeqs >= (int64_t) 1 << 30) return bad("vft_nj_build: too many sequences");
    vft_nj_options opt;
    if (opt_in) opt = *opt_in; else vft_nj_default_options(&opt);
    if (cfg->useMatrix && !tables) return bad("vft_nj_build: useMatrix needs the four tables");
    // limits of the kernels this phase uses (include/vft_b200.h), checked before anything is allocated or computed
    if (((cfg->nPos + 31) / 32 * 32) * 16 > 200 * 1024) return bad("vft_nj_build: alignments longer than 12 800 columns are not supported (term buffer of the average kernel)");
    {
        const double mm = opt.tophitsMult > 0 ? 0.5 + opt.tophitsMult * std::sqrt((double) cfg->nSeqs) : 0;
        if (3 * mm > 4096) return bad("vft_nj_build: top-hit lists longer than 1 365 entries are not supported (m = tophitsMult * sqrt(nSeqs); ~1.8 million taxa at the default)");
    }
    // zero the output counters, keep the caller's buffers
    res->root = -1; res->maxnode = 0; res->m = 0;
    res->nSeeds = res->nCloseUsed = res->nRefreshTopHits = res->nVisibleUpdate = res->nHillBetter = 0;
    res->nOutPrefetchHit = res->nOutSingleFetch = res->nPairPrefetchHit = res->nPairSingleFetch = res->nDeviceCalls = 0;
    res->nSpecHit = res->nSpecMiss = 0;
    res->secondsLeafTopHits = res->secondsJoins = res->secondsTotal = 0;
    res->deviceMsResident = res->secondsEndToEnd = res->secondsInCalls = 0;
    for (double &x : res->secondsHost) x = 0;
    auto e2e0 = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - e2e0).count(); };
    vft_ctx *ctx = nullptr;
    int rc = vft_ctx_create(cfg, &ctx);
    if (rc != VFT_OK) return rc;
    const double tCreate = since();
    if (cfg->useMatrix) rc = vft_upload_tables(ctx, tables[0], tables[1], tables[2], tables[3]);
    if (rc == VFT_OK) rc = vft_upload_leaves(ctx, codes);
    const double tUpload = since();
    if (rc == VFT_OK) rc = cfg->precision == 32 ? run<float>(ctx, *cfg, opt, codes, res) : run<double>(ctx, *cfg, opt, codes, res);
    const double tRun = since();
    if (rc == VFT_OK) vft_get_counters(ctx, &res->counters);
    vft_ctx_destroy(ctx);
    res->secondsEndToEnd = since();
    if (std::getenv("VFT_TIMING"))
        std::fprintf(stderr, "[vft_nj_build] create %.1f ms, upload %.1f ms, run %.1f ms, destroy %.1f ms\n", 1e3 * tCreate,
                     1e3 * (tUpload - tCreate), 1e3 * (tRun - tUpload), 1e3 * (res->secondsEndToEnd - tRun));
    return rc;
}
