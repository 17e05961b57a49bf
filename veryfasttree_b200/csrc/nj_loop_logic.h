// nj_loop_logic.h -- the join loop of fastNJ with the top-hits heuristic (NeighbourJoining.tcc:2857-3100, :4137-4533)
// as DEVICE-RESIDENT state + block-parallel (SPMD) logic: the host is out of the loop.
//
// Round 1 kept the reference's bookkeeping (top-hit lists, visible / top-visible sets, lazily refreshed
// out-distances, the hill-climb) on the host and paid one synchronous device round trip plus ~40 us of host list
// work per join.  Here that state lives in device memory and ONE thread block replays the reference's decisions:
// every loop of the reference over a list (the top-visible scan, getBestFromTopHits, uniqueBestHits,
// sortSaveBestHits, updateVisible, updateTopVisible) is a data-parallel step of that block, the sequence of steps is
// the reference's.  The distances those steps need are requested one step ahead (a hint list evaluated by the whole
// grid, vft_cuda.cu: k_nj_eval); a value that was not hinted is computed by the block itself, so -- exactly as in
// nj_host.cpp -- the decisions never depend on the hints and the tree is the reference's at `-threads 1`.
//
// The same source compiles for the CPU double (oracle/nj_loop_cpu.cpp) with a one-thread execution environment,
// so the logic is tested without a GPU against the golden trees; only the environment differs:
//   X::tid(), X::nt(), X::sync()                  block geometry and barrier
//   X::atomicAddI / atomicExchI                   block/global atomics
//   X::evalOut(ids, n) / evalPairs(a, b, n, d, w) distances the block computes itself (misses of the hint lists)
//
// Parallelisation rules (why the block-parallel steps give the reference's results):
//   * a criterion (setCriterion, NJ.tcc:1085-1113) uses, for each node, the fresh out-distance if the node is stale and
//     the stored one (rescaled) otherwise; refreshing a stale node does not change any criterion of the same epoch,
//     only the state later epochs see.  So a step first commits the refreshes of exactly the nodes the reference's loop
//     would touch, then evaluates all criteria from the committed state;
//   * "first strictly smaller wins" scans = minimum of (criterion, position);
//   * psort order (key ascending, ties in REVERSE input order; oracle/psort_probe.cpp) = rank by (key, -position).
#pragma once
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define NJL_D __device__ __forceinline__
#define NJL_DN __device__ __noinline__
#else
#define NJL_D inline
#define NJL_DN inline
#endif

namespace njl {

enum { ST_RUNNING = 0, ST_DONE = 1, ST_NEED_RESET = 2, ST_NEED_REFRESH = 3, ST_NEED_REBUILD = 4, ST_NEED_HOST = 5, ST_ERROR = 9 };
enum { RS_SEARCH = 0, RS_THJ_FINISH = 1 };

struct Scalars {
    // configuration (written once)
    int64_t nSeqs, maxnodes;
    int32_t m, nTV, cap, tophitAgeLimit, nRefreshMin, nResetOutProfile;
    double staleOutLimit, fResetOutProfile;
    int64_t Lbytes, profBytes;            // algorithmic bytes of a leaf / an internal profile
    // dynamic state of the loop
    int32_t status, resume;
    int32_t nActive;                      // active nodes at the current search
    int32_t maxnode, epoch, stamp, topvisibleAge, nActiveOutProfileReset;
    double totdiam;
    // the join being carried out: consumed by the averageProfile kernel and by thj_finish
    int32_t jdValid, jdNew, jdI, jdJ, jdNActiveOld, jdUpdate, jdSelfPending;
    double jdDiameter;
    // request list for the grid-wide evaluation
    int32_t nOutReq, nPairReq;
    int32_t nUnique;                      // candidates of the pending topHitJoin
    int32_t hintNode[2], hintEpoch, hintJoinSlot, hintJoinI, hintJoinJ;
    int32_t visfixPending;
    int32_t rtvVisible;                   // candidates of the last top-visible rebuild
    // counters
    int64_t nJoins, nRefresh, nVisibleUpdate, nHillBetter, nReset, nInlineOut, nInlinePair, nOutHit, nPairHit, nRebuild;
    int64_t seqOps, profileOps, outprofileOps, algoBytes;
    int64_t tPhase[16], tLast;            // cycles per phase of the step (diagnostic: VFT_LOOP_TIMING)
};

template<typename P>
struct State {
    Scalars *sc;
    int32_t *parent, *up, *child;         // [M], [M], [M*3]
    P *branchlength, *diameter;           // [M]
    P *outDist; int32_t *nOutAct;         // [M] lazily refreshed out-distances (NJ.h:278-292)
    P *freshVal; int32_t *freshEpoch, *wantStamp;   // [M] fresh out-distances of the current epoch, as evaluated
    int32_t *hitJ; P *hitDist; int32_t *hitCount, *age;   // top-hit lists [M*m], [M]
    int32_t *visJ; P *visDist;            // [M] visible set
    int32_t *topvisible;                  // [nTV]
    int32_t *reqOut;                      // [capOut] out-distance requests (node ids)
    int32_t *reqA, *reqB; P *pairD, *pairW;   // [capPair] pair requests and their results
    int32_t *uJ, *uSlot;                  // [cap] candidates of the pending topHitJoin and their request slots
    int32_t *lSlot;                       // [2*m] request slots of the two hinted lists
    P *tmpP;                              // [cap] scratch
    int64_t *joins;                       // [(nSeqs-3)*2] trace (may be null)
    int32_t capOut, capPair;
};

// block scratch (shared memory on the device)
template<typename P>
struct Scratch {
    int32_t *cJ, *cAux, *perm;            // [cap] candidate j, auxiliary int, permutation
    P *cDist, *cCrit;                     // [cap]
    uint64_t *key;                        // [cap]
    int32_t *list;                        // [4*cap] node lists for ensure()
    int32_t *miss;                        // [4*cap]
    uint64_t *redK; int32_t *redI;        // [nt] reductions
    int32_t *ctl;                         // [32] block-uniform control words
};
template<typename P>
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t scratch_bytes(int cap, int nt) {
    return (size_t) cap * (3 * 4 + 2 * sizeof(P) + 8) + (size_t) 8 * cap * 4 + (size_t) nt * 12 + 32 * 4 + 64;
}
template<typename P>
NJL_D void scratch_carve(Scratch<P> &s, unsigned char *base, int cap, int nt) {
    unsigned char *p = base;
    s.key = (uint64_t *) p; p += (size_t) cap * 8;
    s.redK = (uint64_t *) p; p += (size_t) nt * 8;
    s.cDist = (P *) p; p += (size_t) cap * sizeof(P);
    s.cCrit = (P *) p; p += (size_t) cap * sizeof(P);
    s.cJ = (int32_t *) p; p += (size_t) cap * 4;
    s.cAux = (int32_t *) p; p += (size_t) cap * 4;
    s.perm = (int32_t *) p; p += (size_t) cap * 4;
    s.list = (int32_t *) p; p += (size_t) 4 * cap * 4;
    s.miss = (int32_t *) p; p += (size_t) 4 * cap * 4;
    s.redI = (int32_t *) p; p += (size_t) nt * 4;
    s.ctl = (int32_t *) p;
}

// ---- separately rounded arithmetic (the device build also uses -fmad=false; g++ builds with -ffp-contract=off) -------
#ifdef __CUDACC__
NJL_D float  q_add(float a, float b) { return __fadd_rn(a, b); }
NJL_D double q_add(double a, double b) { return __dadd_rn(a, b); }
NJL_D float  q_sub(float a, float b) { return __fsub_rn(a, b); }
NJL_D double q_sub(double a, double b) { return __dsub_rn(a, b); }
NJL_D double q_mul(double a, double b) { return __dmul_rn(a, b); }
#else
NJL_D float  q_add(float a, float b) { return a + b; }
NJL_D double q_add(double a, double b) { return a + b; }
NJL_D float  q_sub(float a, float b) { return a - b; }
NJL_D double q_sub(double a, double b) { return a - b; }
NJL_D double q_mul(double a, double b) { return a * b; }
#endif

// orderable keys: ascending unsigned order == ascending floating order (-0 == +0, as the reference's `<` sees them)
NJL_D uint64_t okey(float x) { if (x == 0) x = 0; uint32_t u; memcpy(&u, &x, 4); return (u & 0x80000000u) ? (uint32_t) ~u : (u | 0x80000000u); }
NJL_D uint64_t okey(double x) { if (x == 0) x = 0; uint64_t u; memcpy(&u, &x, 8); return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull); }

// the same rules as free functions, for the grid-wide kernels of the rebuild / refresh paths
NJL_D bool stale_fn(int32_t nOutAct, int32_t nActive, double staleOutLimit) {
    const int64_t allow = (int64_t) ((double) (int64_t) nActive * staleOutLimit);
    return (int64_t) nOutAct - (int64_t) nActive > allow;
}
NJL_D int32_t ancestor_fn(int32_t *up, int32_t i) {
    if (i < 0) return i;
    for (;;) {
        const int32_t u = up[i];
        if (u == i) return i;
        const int32_t g = up[u];
        up[i] = g;
        i = g;
    }
}
// setCriterion (NJ.tcc:1085-1113) WITHOUT committing its refreshes: a stale node enters with its fresh value
template<typename P>
NJL_D P crit_effective(const State<P> &st, int32_t i, int32_t j, P dist, int32_t nActive) {
    const double lim = st.sc->staleOutLimit;
    int32_t ni = st.nOutAct[i], nj = st.nOutAct[j];
    double outI, outJ;
    if (stale_fn(ni, nActive, lim)) { outI = (double) st.freshVal[i]; ni = nActive; } else outI = (double) st.outDist[i];
    if (stale_fn(nj, nActive, lim)) { outJ = (double) st.freshVal[j]; nj = nActive; } else outJ = (double) st.outDist[j];
    if (ni != nActive) outI = q_mul(outI, (double) (int64_t) (nActive - 1) / (double) ((int64_t) ni - 1));
    if (nj != nActive) outJ = q_mul(outJ, (double) (int64_t) (nActive - 1) / (double) ((int64_t) nj - 1));
    return (P) q_sub((double) dist, q_add(outI, outJ) / (double) (int64_t) (nActive - 2));
}

template<typename P, class X>
struct Logic {
    State<P> st;
    Scratch<P> sm;
    X &x;
    Scalars &sc;
    NJL_D Logic(const State<P> &s, const Scratch<P> &scr, X &env) : st(s), sm(scr), x(env), sc(*s.sc) {}

    // ---- tree ----------------------------------------------------------------------------------------------------
    // activeAncestor (NJ.tcc:536-544) over a pointer-jumping shortcut array; concurrent halving is safe (every value
    // written is an ancestor of the slot's node)
    NJL_D int32_t ancestor(int32_t i) const {
        if (i < 0) return i;
        for (;;) {
            const int32_t u = st.up[i];
            if (u == i) return i;
            const int32_t g = st.up[u];
            st.up[i] = g;
            i = g;
        }
    }
    NJL_D bool active(int32_t i) const { return i >= 0 && st.parent[i] < 0; }

    // ---- out-distances -------------------------------------------------------------------------------------------
    NJL_D bool stale(int32_t i, int32_t nActive) const {                       // trigger of NJ.tcc:1092-1098
        const int64_t allow = (int64_t) ((double) (int64_t) nActive * sc.staleOutLimit);
        return (int64_t) st.nOutAct[i] - (int64_t) nActive > allow;
    }
    // setCriterion (NJ.tcc:1099-1107) from committed state
    NJL_D P crit(int32_t i, int32_t j, P dist, int32_t nActive) const {
        double outI = (double) st.outDist[i];
        const int32_t ni = st.nOutAct[i];
        if (ni != nActive) outI = q_mul(outI, (double) (int64_t) (nActive - 1) / (double) ((int64_t) ni - 1));
        double outJ = (double) st.outDist[j];
        const int32_t nj = st.nOutAct[j];
        if (nj != nActive) outJ = q_mul(outJ, (double) (int64_t) (nActive - 1) / (double) ((int64_t) nj - 1));
        return (P) q_sub((double) dist, q_add(outI, outJ) / (double) (int64_t) (nActive - 2));
    }

    // Refresh exactly the nodes of sm.list[0..n) the reference would refresh here: the stale ones (setCriterion), or
    // -- `always` -- every one not at this nActive (setOutDistance, NJ.tcc:1012-1053).  Values come from the hint
    // evaluation of this epoch when present, else the block computes them.  Block-wide; n is block-uniform.
    NJL_DN void ensureCommit(int n, int32_t nActive, bool always) {
        const int32_t epoch = sc.epoch;
        // common case, one pass: every value needed is in hand (hinted) and is committed at once
        int missing = 0;
        for (int e = x.tid(); e < n; e += x.nt()) {
            const int32_t i = sm.list[e];
            if (i < 0) continue;
            const bool need = always ? st.nOutAct[i] != nActive : stale(i, nActive);
            if (!need) continue;
            if (st.freshEpoch[i] == epoch) { st.outDist[i] = st.freshVal[i]; st.nOutAct[i] = nActive; }      // concurrent writers store the same values
            else missing = 1;
        }
        if (!x.syncOr(missing)) return;
        // some values were not hinted: the block computes them itself
        if (x.tid() == 0) { sm.ctl[0] = 0; sc.stamp++; }
        x.sync();
        const int32_t stamp = sc.stamp;
        for (int e = x.tid(); e < n; e += x.nt()) {
            const int32_t i = sm.list[e];
            if (i < 0) continue;
            const bool need = always ? st.nOutAct[i] != nActive : stale(i, nActive);
            if (need && st.freshEpoch[i] != epoch && x.atomicExchI(&st.wantStamp[i], stamp) != stamp)
                sm.miss[x.atomicAddI(&sm.ctl[0], 1)] = i;
        }
        x.sync();
        const int nMiss = sm.ctl[0];
        x.evalOut(sm.miss, nMiss, nActive);                   // writes freshVal / freshEpoch
        if (x.tid() == 0) sc.nInlineOut += nMiss;
        x.sync();
        for (int e = x.tid(); e < n; e += x.nt()) {
            const int32_t i = sm.list[e];
            if (i < 0) continue;
            const bool need = always ? st.nOutAct[i] != nActive : stale(i, nActive);
            if (need) { st.outDist[i] = st.freshVal[i]; st.nOutAct[i] = nActive; }
        }
        x.sync();
    }

    // ---- block reductions ----------------------------------------------------------------------------------------
    // minimum of (key, idx) over the threads' candidates; idx < 0 = no candidate.  Returns the winning idx (or -1) to
    // every thread; *keyOut its key.
    NJL_D int32_t blockMin(uint64_t k, int32_t idx, uint64_t *keyOut = nullptr) { return x.blockMin(sm.redK, sm.redI, k, idx, keyOut); }
    NJL_D int32_t blockSum(int32_t v) { return x.blockSum(sm.redI, v); }
    // psort order of sm.key[0..n): perm[rank] = element; rank = #smaller keys + #equal keys at LATER positions
    NJL_DN void rankSort(int n) {
        for (int e = x.tid(); e < n; e += x.nt()) {
            const uint64_t k = sm.key[e];
            int r = 0;
            for (int q = 0; q < n; q++) { const uint64_t kq = sm.key[q]; r += (kq < k || (kq == k && q > e)) ? 1 : 0; }
            sm.perm[r] = e;
        }
        x.sync();
    }

    // ---- pair distances ------------------------------------------------------------------------------------------
    // distance half of setDistCriterion (NJ.tcc:1115-1122) for sm.cJ[e] (e in the marked set cAux[e] < -1: needs a distance):
    // from the hinted slot when there is one (cAux[e] = slot >= 0 set by the caller), else computed by the block
    NJL_DN void resolvePairs(int32_t iNode, int n) {
        // entries without a usable slot (cAux[e] == -2) are computed by the block: entry numbers in miss[0..), the (a,b) pairs
        // behind them (miss + cap), so that sm.list survives for the caller's ensureCommit
        int32_t *ent = sm.miss, *pairs = sm.miss + sc.cap;
        if (x.tid() == 0) sm.ctl[1] = 0;
        x.sync();
        for (int e = x.tid(); e < n; e += x.nt()) {
            const int32_t a = sm.cAux[e];
            if (a == -2) ent[x.atomicAddI(&sm.ctl[1], 1)] = e;
            else if (a >= 0) sm.cDist[e] = st.pairD[a];
        }
        x.sync();
        const int nMiss = sm.ctl[1];
        if (nMiss > 0) {
            for (int q = x.tid(); q < nMiss; q += x.nt()) { pairs[2 * q] = iNode; pairs[2 * q + 1] = sm.cJ[ent[q]]; }
            x.sync();
            x.evalPairs(pairs, nMiss, sm.cCrit);                  // distance of pair q into cCrit[q] (scratch)
            x.sync();
            for (int q = x.tid(); q < nMiss; q += x.nt()) sm.cDist[ent[q]] = sm.cCrit[q];
            if (x.tid() == 0) sc.nInlinePair += nMiss;
        }
        x.sync();
    }

    // ---- getBestFromTopHits (NJ.tcc:4267-4298) ---------------------------------------------------------------------
    struct Best { int32_t j; P dist, crit; };
    NJL_DN Best getBest(int32_t iNode, int32_t nActive) {
        const int n = st.hitCount[iNode], m = sc.m;
        const int32_t *hj = st.hitJ + (size_t) iNode * m;
        const P *hd = st.hitDist + (size_t) iNode * m;
        int hinted = -1;
        if (sc.hintEpoch == sc.epoch) { if (sc.hintNode[0] == iNode) hinted = 0; else if (sc.hintNode[1] == iNode) hinted = 1; }
        // setOutDistance(iNode) (:4276), unconditional
        if (x.tid() == 0) sm.list[0] = iNode;
        x.sync();
        ensureCommit(1, nActive, true);
        for (int k = x.tid(); k < n; k += x.nt()) {
            const int32_t j0 = hj[k], j = ancestor(j0);
            int32_t aux = -1;                                    // -1: stored distance; -3: dropped
            if (j < 0 || j == iNode) aux = -3;
            else if (j != j0) {
                aux = -2;
                if (hinted >= 0) { const int32_t s = st.lSlot[hinted * m + k]; if (s >= 0 && st.reqA[s] == iNode && st.reqB[s] == j) aux = s; }
            }
            sm.cJ[k] = j; sm.cAux[k] = aux; sm.cDist[k] = hd[k];
            sm.list[k] = aux == -3 ? -1 : j;
        }
        x.sync();
        resolvePairs(iNode, n);
        ensureCommit(n, nActive, false);
        uint64_t bk = ~0ull; int32_t bi = -1;
        for (int k = x.tid(); k < n; k += x.nt()) {
            if (sm.cAux[k] == -3) continue;
            const P c = crit(iNode, sm.cJ[k], sm.cDist[k], nActive);
            sm.cCrit[k] = c;
            if (c < (P) 1e20) { const uint64_t kk = okey(c); if (bi < 0 || kk < bk) { bk = kk; bi = k; } }   // first strictly smaller wins
        }
        const int32_t w = blockMin(bk, bi);
        Best b; b.j = -1; b.dist = (P) 1e20; b.crit = (P) 1e20;
        if (w >= 0) { b.j = sm.cJ[w]; b.dist = sm.cDist[w]; b.crit = sm.cCrit[w]; }
        x.sync();
        return b;
    }

    // ---- topHitNJSearch (NJ.tcc:4137-4264); returns false when the top-visible set must be rebuilt first -----------
    NJL_DN bool searchDecide(int32_t &ji, int32_t &jj, P &jdist) {
        const int32_t nActive = sc.nActive, nTV = sc.nTV;
        // scan of the top-visible set: getVisible (NJ.tcc:546-557) for every entry
        for (int k = x.tid(); k < nTV; k += x.nt()) {
            const int32_t i = st.topvisible[k];
            int32_t a = -1, b = -1;
            if (active(i)) { const int32_t vj = st.visJ[i]; if (active(vj)) { a = i; b = vj; } }
            sm.list[2 * k] = a; sm.list[2 * k + 1] = b;
        }
        x.sync();
        ensureCommit(2 * nTV, nActive, false);
        uint64_t bk = ~0ull; int32_t bi = -1, cnt = 0;
        for (int k = x.tid(); k < nTV; k += x.nt()) {
            const int32_t i = sm.list[2 * k], vj = sm.list[2 * k + 1];
            if (i < 0) continue;
            cnt++;
            const uint64_t kk = okey(crit(i, vj, st.visDist[i], nActive));
            if (bi < 0 || kk < bk) { bk = kk; bi = k; }
        }
        const int32_t wk = blockMin(bk, bi);
        const int32_t nCandidate = blockSum(cnt);
        if (x.tid() == 0) sc.topvisibleAge++;
        x.sync();
        if (2 * (int64_t) sc.topvisibleAge > sc.m || (3 * (int64_t) nCandidate < nTV && 3 * (int64_t) nCandidate < nActive)) {
            if (x.tid() == 0) { sc.visfixPending = sc.topvisibleAge <= 2 ? 1 : 0; }      // NJ.tcc:4171-4201
            x.sync();
            return false;
        }
        ji = st.topvisible[wk]; jj = st.visJ[ji]; jdist = st.visDist[ji];
        P jcrit = crit(ji, jj, jdist, nActive);
        tmark(6);
        // hill-climbing, NJ.tcc:4222-4263
        bool changed;
        do {
            changed = false;
            Best b = getBest(ji, nActive);
            if (b.j != jj && b.crit < jcrit) { changed = true; jj = b.j; jdist = b.dist; jcrit = b.crit; }
            b = getBest(jj, nActive);
            if (b.j != ji && b.crit < jcrit) { changed = true; const int32_t oi = jj; ji = oi; jj = b.j; jdist = b.dist; jcrit = b.crit; }
            if (changed && x.tid() == 0) sc.nHillBetter++;
        } while (changed);
        tmark(7);
        return true;
    }

    // ---- the join itself (NJ.tcc:2897-3043 without BIONJ): everything but the profile arithmetic ----------------------
    NJL_DN void joinBookkeeping(int32_t ji, int32_t jj) {
        const int32_t nActive = sc.nActive;
        // :2897-2900 -- out-distances of the pair up to date, distance recomputed
        if (x.tid() == 0) { sm.list[0] = ji; sm.list[1] = jj; }
        x.sync();
        ensureCommit(2, nActive, true);
        if (x.tid() == 0) {
            sm.cJ[0] = jj; sm.cAux[0] = -2;
            if (sc.hintEpoch == sc.epoch && sc.hintJoinSlot >= 0 && ((sc.hintJoinI == ji && sc.hintJoinJ == jj) || (sc.hintJoinI == jj && sc.hintJoinJ == ji))) {
                sm.cAux[0] = sc.hintJoinSlot;
            }
        }
        x.sync();
        resolvePairs(ji, 1);
        if (x.tid() == 0) {
            const P dist = sm.cDist[0];
            const int32_t newnode = sc.maxnode++;
            st.parent[ji] = newnode; st.parent[jj] = newnode;
            st.up[ji] = newnode; st.up[jj] = newnode;
            const int32_t c0 = ji < jj ? ji : jj, c1 = ji < jj ? jj : ji;
            st.child[3 * (size_t) newnode] = c0; st.child[3 * (size_t) newnode + 1] = c1; st.child[3 * (size_t) newnode + 2] = -1;
            if (st.joins) { const int64_t t = sc.nSeqs - nActive; st.joins[2 * t] = c0; st.joins[2 * t + 1] = c1; }
            // :2911-2916
            const double distIJ = (double) dist;
            const double deltaDist = (double) q_sub(st.outDist[ji], st.outDist[jj]) / (double) (int64_t) (nActive - 2);
            st.branchlength[ji] = (P) (q_add(distIJ, deltaDist) / 2);
            st.branchlength[jj] = (P) (q_sub(distIJ, deltaDist) / 2);
            // :3003-3007 with bionjWeight = 0.5
            const double a = (double) q_add(st.branchlength[ji], st.diameter[ji]), b = (double) q_add(st.branchlength[jj], st.diameter[jj]);
            const P dn = (P) q_add(q_mul(0.5, a), q_mul(0.5, b));
            st.diameter[newnode] = dn;
            const int64_t changedActive = (int64_t) sc.nActiveOutProfileReset - ((int64_t) nActive - 1);
            const bool rebuild = changedActive >= sc.nResetOutProfile && (double) changedActive >= sc.fResetOutProfile * (double) sc.nActiveOutProfileReset;
            if (!rebuild) sc.totdiam = q_add(sc.totdiam, (double) q_sub(q_sub(dn, st.diameter[ji]), st.diameter[jj]));
            sc.jdValid = 1; sc.jdSelfPending = 1; sc.jdNew = newnode; sc.jdI = ji; sc.jdJ = jj; sc.jdNActiveOld = nActive; sc.jdUpdate = rebuild ? 0 : 1;
            sc.jdDiameter = (double) dn;
            st.nOutAct[newnode] = 2000000000;                   // never computed: stale at any nActive
            st.outDist[newnode] = 0;
            st.hitCount[newnode] = 0;
            st.visJ[newnode] = -1;
            sc.epoch++;                                          // newEpoch(nActive - 1)
            sc.nActive = nActive - 1;
            sc.nJoins++;
            sc.profileOps++;                                     // averageProfile's self distance
            if (rebuild) { sc.status = ST_NEED_REBUILD; sc.nRebuild++; }
        }
        x.sync();
    }

    // ---- request lists --------------------------------------------------------------------------------------------
    NJL_D void wantOut(int32_t i, int32_t nActive, bool evenIfNotStale) {
        if (!active(i)) return;
        if (st.nOutAct[i] == nActive || st.freshEpoch[i] == sc.epoch) return;
        if (!evenIfNotStale && !stale(i, nActive)) return;
        if (x.atomicExchI(&st.wantStamp[i], sc.stamp) == sc.stamp) return;
        const int q = x.atomicAddI(&sc.nOutReq, 1);
        if (q < st.capOut) st.reqOut[q] = i; else x.atomicAddI(&sc.nOutReq, -1);
    }
    NJL_D int32_t wantPair(int32_t a, int32_t b) {
        const int q = x.atomicAddI(&sc.nPairReq, 1);
        if (q >= st.capPair) { x.atomicAddI(&sc.nPairReq, -1); return -1; }
        st.reqA[q] = a; st.reqB[q] = b;
        return q;
    }
    NJL_DN void beginRequests() {
        if (x.tid() == 0) { sc.nOutReq = 0; sc.nPairReq = 0; sc.stamp++; sc.hintEpoch = -1; sc.hintJoinSlot = -1; sc.hintNode[0] = sc.hintNode[1] = -1; }
        x.sync();
    }

    // what the coming topHitNJSearch will most likely ask for (nj_host.cpp speculateSearch): the out-distances its scan
    // can refresh, the two top-hit lists of the likely join, the join's own distance.  A hint only.
    NJL_DN void hintSearch() {
        const int32_t nActive = sc.nActive, nTV = sc.nTV, m = sc.m;
        uint64_t bk = ~0ull; int32_t bi = -1;
        for (int k = x.tid(); k < nTV; k += x.nt()) {
            const int32_t i = st.topvisible[k];
            if (!active(i)) continue;
            const int32_t vj = st.visJ[i];
            if (!active(vj)) continue;
            wantOut(i, nActive, false); wantOut(vj, nActive, false);
            const uint64_t kk = okey(crit(i, vj, st.visDist[i], nActive));      // from the stale values: a guess
            if (bi < 0 || kk < bk) { bk = kk; bi = k; }
        }
        const int32_t wk = blockMin(bk, bi);
        if (wk < 0) return;
        const int32_t g = st.topvisible[wk], gj = st.visJ[g];
        for (int side = 0; side < 2; side++) {
            const int32_t node = side ? gj : g;
            const int n = st.hitCount[node];
            if (x.tid() == 0) wantOut(node, nActive, true);
            for (int k = x.tid(); k < m; k += x.nt()) {
                int32_t slot = -1;
                if (k < n) {
                    const int32_t j0 = st.hitJ[(size_t) node * m + k], j = ancestor(j0);
                    if (j >= 0 && j != node) {
                        if (j != j0) slot = wantPair(node, j);
                        wantOut(j, nActive, false);
                    }
                }
                st.lSlot[side * m + k] = slot;
            }
        }
        if (x.tid() == 0) {
            sc.hintNode[0] = g; sc.hintNode[1] = gj; sc.hintEpoch = sc.epoch;
            sc.hintJoinSlot = wantPair(g, gj); sc.hintJoinI = g; sc.hintJoinJ = gj;
        }
        x.sync();
    }

    // ---- topHitJoin, first half (NJ.tcc:4306-4340): the children's lists merged into the candidate set of the new node
    NJL_DN void thjPrepare() {
        const int32_t newnode = sc.jdNew, nActive = sc.nActive, m = sc.m;
        const int32_t c0 = st.child[3 * (size_t) newnode], c1 = st.child[3 * (size_t) newnode + 1];
        const int n0 = st.hitCount[c0], n1 = st.hitCount[c1], n = n0 + n1;
        for (int e = x.tid(); e < n; e += x.nt()) {
            const int32_t j0 = e < n0 ? st.hitJ[(size_t) c0 * m + e] : st.hitJ[(size_t) c1 * m + (e - n0)];
            const int32_t j = ancestor(j0);
            const bool ok = j >= 0 && j != newnode;
            sm.cJ[e] = j;
            sm.key[e] = ok ? (uint64_t) j + 1 : 0;               // psort by (i,j): every valid entry has i = newnode (:4797)
        }
        x.sync();
        rankSort(n);
        // first of each run of equal j survives (:4802-4821); the survivors keep their sorted (ascending j) order
        if (x.tid() == 0) sm.ctl[2] = 0;
        x.sync();
        for (int r = x.tid(); r < n; r += x.nt()) {
            const uint64_t k = sm.key[sm.perm[r]];
            sm.cAux[r] = (k != 0 && (r == 0 || sm.key[sm.perm[r - 1]] != k)) ? 1 : 0;
        }
        x.sync();
        // ordered compaction: position = number of survivors before r
        for (int r = x.tid(); r < n; r += x.nt()) {
            if (!sm.cAux[r]) continue;
            int pos = 0;
            for (int q = 0; q < r; q++) pos += sm.cAux[q];
            st.uJ[pos] = sm.cJ[sm.perm[r]];
        }
        int mine = 0;
        for (int r = x.tid(); r < n; r += x.nt()) mine += sm.cAux[r];
        const int32_t nUnique = blockSum(mine);
        if (x.tid() == 0) sc.nUnique = nUnique;
        x.sync();
        // requests: every candidate needs its distance to the new node (:4826: the stored ones belong to the children);
        // the out-distances setCriterion / updateVisible / updateTopVisible may refresh (a superset is harmless)
        for (int u = x.tid(); u < nUnique; u += x.nt()) {
            const int32_t j = st.uJ[u];
            st.uSlot[u] = wantPair(newnode, j);
            wantOut(j, nActive, false);
            const int32_t vj = st.visJ[j];
            if (active(vj)) wantOut(vj, nActive, false);
        }
        if (x.tid() == 0) {
            const int q = x.atomicAddI(&sc.nOutReq, 1);          // the new node itself: never computed
            if (q < st.capOut) { st.reqOut[q] = newnode; st.wantStamp[newnode] = sc.stamp; } else x.atomicAddI(&sc.nOutReq, -1);
        }
        x.sync();
    }

    // ---- updateTopVisible (NJ.tcc:4661-4711) ------------------------------------------------------------------------
    NJL_DN void updateTopVisible(int32_t iIn, int32_t hitJ_, P hitDist_, int32_t nActive) {
        const int32_t nTV = sc.nTV;
        int32_t bi = -1;
        for (int k = x.tid(); k < nTV; k += x.nt()) {
            const int32_t i = st.topvisible[k];
            if ((i == iIn || !active(i)) && bi < 0) bi = k;
        }
        const int32_t first = blockMin(bi >= 0 ? (uint64_t) bi : ~0ull, bi);
        if (first >= 0) {
            if (x.tid() == 0 && st.topvisible[first] != iIn) st.topvisible[first] = iIn;
            x.sync();
            return;
        }
        // second scan: stops at the first entry without a live visible hit, or that is the reciprocal of the new one
        bi = -1;
        for (int k = x.tid(); k < nTV; k += x.nt()) {
            const int32_t i = st.topvisible[k], vj = st.visJ[i];
            const bool ok = active(vj);
            const bool stop = !ok || (i == hitJ_ && vj == iIn);
            sm.cAux[k] = ok ? 1 : 0;
            if (stop && bi < 0) bi = k;
        }
        const int32_t kStop = blockMin(bi >= 0 ? (uint64_t) bi : ~0ull, bi);
        const int32_t kEnd = kStop < 0 ? nTV : (sm.cAux[kStop] ? kStop + 1 : kStop);     // entries whose criterion the reference evaluates
        for (int k = x.tid(); k < nTV; k += x.nt()) {
            int32_t a = -1, b = -1;
            if (k < kEnd) { a = st.topvisible[k]; b = st.visJ[a]; }
            sm.list[2 * k] = a; sm.list[2 * k + 1] = b;
        }
        x.sync();
        ensureCommit(2 * nTV, nActive, false);
        if (kStop >= 0) {
            if (x.tid() == 0 && !sm.cAux[kStop]) st.topvisible[kStop] = iIn;
            x.sync();
            return;
        }
        // the worst entry: `>=` keeps the LAST maximum (:4696)
        uint64_t wk = 0; int32_t wi = -1;
        for (int k = x.tid(); k < nTV; k += x.nt()) {
            const int32_t i = st.topvisible[k];
            const uint64_t kk = okey(crit(i, st.visJ[i], st.visDist[i], nActive));
            if (wi < 0 || kk >= wk) { wk = kk; wi = k; }
        }
        // maximum of (key, position): as a minimum over the complemented key and position
        uint64_t worstKeyC;
        const int32_t wpos = blockMin(~wk, wi >= 0 ? nTV - 1 - wi : -1, &worstKeyC);
        if (wpos < 0) return;
        const int32_t iPosWorst = nTV - 1 - wpos;
        const uint64_t worstKey = ~worstKeyC;
        if (x.tid() == 0) { sm.list[0] = iIn; sm.list[1] = hitJ_; }
        x.sync();
        ensureCommit(2, nActive, false);
        if (x.tid() == 0) {
            // `v.criterion >= dCriterionWorst` starts from -1e20: a real criterion always replaces it
            const P c = crit(iIn, hitJ_, hitDist_, nActive);
            if (okey(c) < worstKey) st.topvisible[iPosWorst] = iIn;
        }
        x.sync();
    }

    // ---- topHitJoin, second half (NJ.tcc:4326-4437); returns false when the lists must be refreshed ------------------
    NJL_DN bool thjFinish() {
        const int32_t newnode = sc.jdNew, nActive = sc.nActive, m = sc.m, nUnique = sc.nUnique;
        const int32_t c0 = st.child[3 * (size_t) newnode], c1 = st.child[3 * (size_t) newnode + 1];
        // uniqueBestHits' tail (:4823-4831): distance + criterion of every candidate
        for (int u = x.tid(); u < nUnique; u += x.nt()) {
            const int32_t j = st.uJ[u], s = st.uSlot[u];
            sm.cJ[u] = j;
            sm.cAux[u] = (s >= 0 && st.reqA[s] == newnode && st.reqB[s] == j) ? s : -2;
            sm.list[u] = j;
        }
        if (x.tid() == 0) sm.list[nUnique] = newnode;
        x.sync();
        resolvePairs(newnode, nUnique);
        tmark(0);
        ensureCommit(nUnique + 1, nActive, false);
        tmark(1);
        for (int u = x.tid(); u < nUnique; u += x.nt()) {
            const P c = crit(newnode, sm.cJ[u], sm.cDist[u], nActive);
            sm.cCrit[u] = c; sm.key[u] = okey(c);
        }
        if (x.tid() == 0) {
            st.age[newnode] = (st.age[c0] + st.age[c1] + 1) / 2 + 1;     // :4342
            st.hitCount[c0] = 0; st.hitCount[c1] = 0;
        }
        x.sync();
        const bool useUnique = nUnique == nActive - 1 || (st.age[newnode] <= sc.tophitAgeLimit && nUnique >= sc.nRefreshMin);
        if (!useUnique) return false;
        // sortSaveBestHits (:4535-4578): psort by criterion, the first nSave
        rankSort(nUnique);
        const int32_t nSave = nUnique < m ? nUnique : m;
        for (int r = x.tid(); r < nSave; r += x.nt()) {
            const int e = sm.perm[r];
            st.hitJ[(size_t) newnode * m + r] = sm.cJ[e];
            st.hitDist[(size_t) newnode * m + r] = sm.cDist[e];
            st.tmpP[r] = sm.cCrit[e];                              // criteria of the saved hits, in list order (for updateVisible)
        }
        if (x.tid() == 0) {
            st.hitCount[newnode] = nSave;
            const int e0 = sm.perm[0];
            st.visJ[newnode] = sm.cJ[e0]; st.visDist[newnode] = sm.cDist[e0];
        }
        x.sync();
        tmark(2);
        updateTopVisible(newnode, st.visJ[newnode], st.visDist[newnode], nActive);
        tmark(3);
        // updateVisible (:4635-4658) over the saved hits, in order.  Which hits replace a visible entry is decided for all of
        // them first (each test reads only its own node's entry), the replacements -- each followed by an updateTopVisible
        // -- are then made one after the other
        const int32_t *hj = st.hitJ + (size_t) newnode * m;
        const P *hd = st.hitDist + (size_t) newnode * m;
        for (int r = x.tid(); r < nSave; r += x.nt()) {
            const int32_t j = hj[r], vj = st.visJ[j];
            const bool ok = active(vj);
            sm.list[2 * r] = ok ? j : -1; sm.list[2 * r + 1] = ok ? vj : -1;
            sm.cAux[r] = ok ? 1 : 0;
        }
        x.sync();
        ensureCommit(2 * nSave, nActive, false);
        for (int r = x.tid(); r < nSave; r += x.nt()) {
            const int32_t j = hj[r];
            bool flag = true;
            if (sm.cAux[r]) flag = st.tmpP[r] < crit(j, st.visJ[j], st.visDist[j], nActive);
            sm.perm[r] = flag ? (sm.cAux[r] ? 2 : 1) : 0;
        }
        x.sync();
        // the hits that replace a visible entry, in list order (ordered compaction of the flags into sm.key)
        int mineF = 0;
        for (int r = x.tid(); r < nSave; r += x.nt()) {
            if (!sm.perm[r]) continue;
            int pos = 0;
            for (int q = 0; q < r; q++) pos += sm.perm[q] ? 1 : 0;
            sm.key[pos] = (uint64_t) r;
            mineF++;
        }
        const int32_t nFlag = blockSum(mineF);
        tmark(4);
        for (int q = 0; q < nFlag; q++) {
            const int r = (int) sm.key[q];                        // (updateTopVisible leaves sm.key and sm.perm alone)
            const int f = sm.perm[r];
            const int32_t j = hj[r];
            const P d = hd[r];
            if (x.tid() == 0) { if (f == 2) sc.nVisibleUpdate++; st.visJ[j] = newnode; st.visDist[j] = d; }
            x.sync();
            updateTopVisible(j, newnode, d, nActive);
        }
        tmark(5);
        return true;
    }

    NJL_D void tmark(int k) {
        if (x.tid() == 0) { const int64_t t = x.clock(); sc.tPhase[k] += t - sc.tLast; sc.tLast = t; }
    }
    // ---- one step of the loop: [finish the pending topHitJoin] -> search -> join -> prepare its topHitJoin ------------
    NJL_DN void step() {
        if (sc.status != ST_RUNNING) return;
        if (x.tid() == 0) { sc.jdSelfPending = 0; sc.tLast = x.clock(); }   // (the self distance was consumed by the evaluation of the previous step's request list)
        if (sc.resume == RS_THJ_FINISH) {
            if (!thjFinish()) {
                if (x.tid() == 0) { sc.status = ST_NEED_REFRESH; sc.nRefresh++; }
                x.sync();
                return;
            }
            if (x.tid() == 0) sc.resume = RS_SEARCH;
            x.sync();
        }
        if (sc.nActive <= 3) {
            if (x.tid() == 0) sc.status = ST_DONE;
            x.sync();
            return;
        }
        int32_t ji = -1, jj = -1; P jd = 0;
        if (!searchDecide(ji, jj, jd)) {
            if (x.tid() == 0) { sc.status = ST_NEED_RESET; sc.nReset++; }
            x.sync();
            return;
        }
        tmark(8);
        joinBookkeeping(ji, jj);
        tmark(9);
        beginRequests();
        thjPrepare();
        tmark(10);
        if (sc.nActive > 3) hintSearch();
        tmark(11);
        if (x.tid() == 0) sc.resume = RS_THJ_FINISH;
        x.sync();
    }
};

}  // namespace njl
