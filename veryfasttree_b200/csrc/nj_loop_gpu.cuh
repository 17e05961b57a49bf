// nj_loop_gpu.cuh -- the device-resident join loop on the GPU (included at the end of vft_cuda.cu).
//
//   k_nj_step     ONE thread block runs nj_loop_logic.h: finishes the pending topHitJoin, searches the next join,
//                 does its bookkeeping and builds the request list of the distances the next step will need
//   k_average     (vft_cuda.cu) the join's profile arithmetic, its arguments read from the loop's scalars
//   k_nj_eval     the whole grid evaluates the request list (out-distances scattered per node, pair distances by slot)
// One join = those three launches, queued ROUNDS at a time with no host synchronisation in between; the host reads the
// loop's scalars once per round and acts only on a status change (out-profile rebuild, top-visible rebuild, refresh).
#pragma once

namespace {

using njl::Scalars;

template<typename P, int A, bool MATRIX>
struct GpuEnv {
    const Store<P> &s;
    njl::State<P> &st;
    unsigned char *evalSmem;                    // [nWarps] x group_smem_bytes(1)
    __device__ __forceinline__ int tid() const { return (int) threadIdx.x; }
    __device__ __forceinline__ int nt() const { return (int) blockDim.x; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ int atomicAddI(int32_t *p, int v) { return atomicAdd(p, v); }
    __device__ __forceinline__ int atomicExchI(int32_t *p, int v) { return atomicExch(p, v); }
    __device__ __forceinline__ void atomicAddL(int64_t *p, int64_t v) { atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long) v); }
    __device__ __forceinline__ bool syncOr(int pred) { return __syncthreads_or(pred) != 0; }
    __device__ __forceinline__ int64_t clock() const { return (int64_t) clock64(); }
    // minimum of (key, idx) over the block, idx < 0 = no candidate: warp shuffles, the warps' results folded by warp 0
    __device__ __forceinline__ int32_t blockMin(uint64_t *redK, int32_t *redI, uint64_t k, int32_t idx, uint64_t *keyOut) {
        const unsigned full = 0xFFFFFFFFu;
        auto fold = [&](uint64_t &k0, int32_t &i0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint64_t kb = __shfl_down_sync(full, k0, o);
                const int32_t ib = __shfl_down_sync(full, i0, o);
                if (ib >= 0 && (i0 < 0 || kb < k0 || (kb == k0 && ib < i0))) { k0 = kb; i0 = ib; }
            }
        };
        fold(k, idx);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
        if (lane == 0) { redK[warp] = k; redI[warp] = idx; }
        __syncthreads();
        if (warp == 0) {
            uint64_t bk = lane < nWarps ? redK[lane] : ~0ull;
            int32_t bi = lane < nWarps ? redI[lane] : -1;
            fold(bk, bi);
            if (lane == 0) { redK[32] = bk; redI[32] = bi; }
        }
        __syncthreads();
        if (keyOut) *keyOut = redK[32];
        return redI[32];
    }
    __device__ __forceinline__ int32_t blockSum(int32_t *redI, int32_t v) {
        const unsigned full = 0xFFFFFFFFu;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(full, v, o);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
        if (lane == 0) redI[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int32_t t = lane < nWarps ? redI[lane] : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(full, t, o);
            if (lane == 0) redI[33] = t;
        }
        __syncthreads();
        return redI[33];
    }
    // one item per warp, the warps of the block side by side
    __device__ __noinline__ void evalItem(int64_t a, int64_t b, bool isOut, int32_t nActive, P &d, P &w) {
        const unsigned full = 0xFFFFFFFFu;
        const int lane = threadIdx.x & 31;
        unsigned char *smw = evalSmem + (threadIdx.x >> 5) * group_smem_bytes<P, A, MATRIX>(1);
        const bool mine = lane == 0;
        const bool isSeq = !isOut && a < s.nSeqs && b < s.nSeqs;
        d = 0; w = 0;
        if (isSeq) {
            if (mine) {
                seq_dist<P, MATRIX>(s, s.codes + a * s.Lp, s.codes + b * s.Lp, d, w);
                d = (P) xadd((double) d, 0.0);                                            // NJ.tcc:1122
            }
            __syncwarp();
            return;
        }
        const unsigned mask = __ballot_sync(full, mine);
        double den, top;
        group_profile_dist<P, A, MATRIX>(s, mine ? a : (int64_t) -1, mine ? (isOut ? (int64_t) -1 : b) : (int64_t) -1, mask, 1, smw, den, top);
        if (mine) {
            P dd, ww;
            finish_dist<P>(den, top, dd, ww);
            if (isOut) d = out_distance_finish<P>(s, a, nActive, st.sc->totdiam, dd, ww);
            else { d = join_correct<P>(s, a, b, dd); w = ww; }
        }
        __syncwarp();
    }
    // SURVEY 8d accounting of one evaluated pair (the query's own bytes are left out: it is shared by the list)
    __device__ __forceinline__ void account(int64_t a, int64_t b) {
        Scalars *sc = st.sc;
        if (a < s.nSeqs && b < s.nSeqs) { atomicAddL(&sc->seqOps, 1); atomicAddL(&sc->algoBytes, sc->Lbytes); }
        else { atomicAddL(&sc->profileOps, 1); atomicAddL(&sc->algoBytes, b < s.nSeqs ? sc->Lbytes : sc->profBytes); }
    }
    __device__ __noinline__ void evalOut(const int32_t *ids, int n, int32_t nActive) {
        const int warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5, lane = threadIdx.x & 31;
        for (int q = warp; q < n; q += nWarps) {
            const int32_t id = ids[q];
            P d, w;
            evalItem(id, -1, true, nActive, d, w);
            if (lane == 0) { st.freshVal[id] = d; st.freshEpoch[id] = st.sc->epoch; atomicAddL(&st.sc->algoBytes, id < s.nSeqs ? st.sc->Lbytes : st.sc->profBytes); }
        }
        if (threadIdx.x == 0) st.sc->outprofileOps += n;
    }
    __device__ __noinline__ void evalPairs(const int32_t *pairs, int n, P *outD) {
        const int warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5, lane = threadIdx.x & 31;
        for (int q = warp; q < n; q += nWarps) {
            P d, w;
            const int64_t a = pairs[2 * q], b = pairs[2 * q + 1];
            evalItem(a, b, false, 0, d, w);
            if (lane == 0) { outD[q] = d; account(a, b); }
        }
    }
};

constexpr int NJ_STEP_T = 256;

template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(NJ_STEP_T, 1)
k_nj_step(Store<P> s, njl::State<P> st, int hintOnly) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    if (st.sc->status != njl::ST_RUNNING) return;
    njl::Scratch<P> sm;
    njl::scratch_carve<P>(sm, smemRaw, st.sc->cap, NJ_STEP_T);
    const size_t off = (njl::scratch_bytes<P>(st.sc->cap, NJ_STEP_T) + 15) & ~(size_t) 15;
    GpuEnv<P, A, MATRIX> env{s, st, smemRaw + off};
    njl::Logic<P, GpuEnv<P, A, MATRIX>> logic(st, sm, env);
    if (hintOnly) { logic.beginRequests(); logic.hintSearch(); }
    else logic.step();
}

// The request list of the step, by the whole grid: items [0, nOutReq) are setOutDistance requests (scattered into the
// loop's fresh-value table), the rest pair requests (results by slot).  Same arithmetic as k_eval; the counts are read
// from device memory (the host does not know them), G items per warp as the list size demands.
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128, sizeof(P) == 4 ? 4 : 2)
k_nj_eval(Store<P> s, njl::State<P> st, int minItems) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Scalars *sc = st.sc;
    if (sc->status != njl::ST_RUNNING) return;
    const int nOut = sc->nOutReq, n = nOut + sc->nPairReq;
    if (n <= minItems) return;                           // the one-CTA-per-item kernel took this list
    const int32_t nActive = sc->nActive;
    const double totdiam = sc->totdiam;
    const int32_t epoch = sc->epoch;
    const int32_t selfNode = sc->jdSelfPending ? sc->jdNew : -1;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    constexpr int R = TileShape<A, MATRIX>::R;
    const int64_t totalWarps = (int64_t) gridDim.x * (blockDim.x >> 5);
    int G = (int) ((n + totalWarps - 1) / totalWarps);
    G = G < 1 ? 1 : (G > R ? R : G);
    unsigned char *smw = smemRaw + (threadIdx.x >> 5) * group_smem_bytes<P, A, MATRIX>(R);
    const int64_t warp0 = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
    int64_t tSeq = 0, tProf = 0, tBytes = 0;
    for (int64_t warp = warp0; warp * G < n; warp += totalWarps) {
        const int64_t item = warp * G + lane;
        const bool inRange = lane < G && item < n;
        const bool isOut = inRange && item < nOut;
        int64_t a = -1, b = -1;
        if (inRange) { if (isOut) a = st.reqOut[item]; else { a = st.reqA[item - nOut]; b = st.reqB[item - nOut]; } }
        const bool valid = inRange && a >= 0;
        const bool isSeq = valid && !isOut && a < s.nSeqs && b < s.nSeqs;
        P d = 0, w = 0;
        if (isSeq) {
            seq_dist<P, MATRIX>(s, s.codes + a * s.Lp, s.codes + b * s.Lp, d, w);
            d = (P) xadd((double) d, 0.0);
        }
        const bool isProf = valid && !isSeq;
        const unsigned mask = __ballot_sync(full, isProf);
        double den, top;
        // the new node's own out-distance needs its self distance (NJ.tcc:3040-3043), which k_average leaves to this kernel
        const unsigned selfMask = __ballot_sync(full, isOut && a == selfNode);
        if (selfMask) {
            group_profile_dist<P, A, MATRIX>(s, a, a, selfMask, G, smw, den, top);
            if (isOut && a == selfNode) { P sd, sw; finish_dist<P>(den, top, sd, sw); s.selfdist[a] = sd; s.selfweight[a] = sw; }
            __syncwarp();
        }
        group_profile_dist<P, A, MATRIX>(s, a, isOut ? (int64_t) -1 : b, mask, G, smw, den, top);
        if (isProf) {
            P dd, ww;
            finish_dist<P>(den, top, dd, ww);
            if (isOut) d = out_distance_finish<P>(s, a, nActive, totdiam, dd, ww);
            else { d = join_correct<P>(s, a, b, dd); w = ww; }
        }
        if (valid) {
            if (isOut) { st.freshVal[a] = d; st.freshEpoch[a] = epoch; }
            else { st.pairD[item - nOut] = d; st.pairW[item - nOut] = w; }
            if (isSeq) { tSeq++; tBytes += sc->Lbytes; } else { if (!isOut) tProf++; tBytes += (isOut ? a : b) < s.nSeqs ? sc->Lbytes : sc->profBytes; }
        }
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) { tSeq += __shfl_down_sync(full, tSeq, o); tProf += __shfl_down_sync(full, tProf, o); tBytes += __shfl_down_sync(full, tBytes, o); }
    if (lane == 0 && tBytes > 0) {
        if (tSeq) atomicAdd(reinterpret_cast<unsigned long long *>(&sc->seqOps), (unsigned long long) tSeq);
        if (tProf) atomicAdd(reinterpret_cast<unsigned long long *>(&sc->profileOps), (unsigned long long) tProf);
        atomicAdd(reinterpret_cast<unsigned long long *>(&sc->algoBytes), (unsigned long long) tBytes);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc->jdValid = 0; sc->outprofileOps += nOut; sc->nPairHit += n - nOut; }
    // (jdSelfPending is cleared by the next k_nj_step: other CTAs of this launch may still read it)
}

// the same list, one CTA per item (long alignments: a few hundred items cannot fill the machine with a warp each)
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(256)
k_nj_eval_wide(Store<P> s, njl::State<P> st, int maxItems) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Scalars *sc = st.sc;
    if (sc->status != njl::ST_RUNNING) return;
    const int nOut = sc->nOutReq, n = nOut + sc->nPairReq;
    if (n > maxItems) return;                            // long lists go to the grouped kernel
    const int32_t selfNode = sc->jdSelfPending ? sc->jdNew : -1;
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc->jdValid = 0; sc->outprofileOps += nOut; sc->nPairHit += n - nOut; }
    for (int item = blockIdx.x; item < n; item += gridDim.x) {
        const bool isOut = item < nOut;
        const int64_t a = isOut ? st.reqOut[item] : st.reqA[item - nOut], b = isOut ? -1 : st.reqB[item - nOut];
        if (a < 0) continue;
        const bool isSeq = !isOut && a < s.nSeqs && b < s.nSeqs;
        double den, top;
        if (isOut && a == selfNode) {                    // the new node: its self distance first (see k_nj_eval)
            group_profile_dist<P, A, MATRIX, true>(s, a, a, 1u, 1, smemRaw, den, top, (int) (threadIdx.x >> 5), (int) (blockDim.x >> 5));
            __syncthreads();
            if (threadIdx.x < 32) {
                cta_ordered_sum<P, A, MATRIX>(s, smemRaw, den, top);
                if (threadIdx.x == 0) { P sd, sw; finish_dist<P>(den, top, sd, sw); s.selfdist[a] = sd; s.selfweight[a] = sw; }
            }
            __syncthreads();
        }
        group_profile_dist<P, A, MATRIX, true>(s, a, isOut ? (int64_t) -1 : b, 1u, 1, smemRaw, den, top, (int) (threadIdx.x >> 5), (int) (blockDim.x >> 5));
        __syncthreads();
        if (threadIdx.x < 32) {
            cta_ordered_sum<P, A, MATRIX>(s, smemRaw, den, top);
            if (threadIdx.x == 0) {
                P dd, ww, d, w = 0;
                finish_dist<P>(den, top, dd, ww);
                if (isOut) d = out_distance_finish<P>(s, a, sc->nActive, sc->totdiam, dd, ww);
                else if (isSeq) { d = (P) xadd((double) dd, 0.0); w = den > 0 ? ww : (P) 0; }
                else { d = join_correct<P>(s, a, b, dd); w = ww; }
                if (isOut) { st.freshVal[a] = d; st.freshEpoch[a] = sc->epoch; }
                else { st.pairD[item - nOut] = d; st.pairW[item - nOut] = w; }
                if (isSeq) { atomicAdd(reinterpret_cast<unsigned long long *>(&sc->seqOps), 1ull); atomicAdd(reinterpret_cast<unsigned long long *>(&sc->algoBytes), (unsigned long long) sc->Lbytes); }
                else {
                    if (!isOut) atomicAdd(reinterpret_cast<unsigned long long *>(&sc->profileOps), 1ull);
                    atomicAdd(reinterpret_cast<unsigned long long *>(&sc->algoBytes), (unsigned long long) ((isOut ? a : b) < s.nSeqs ? sc->Lbytes : sc->profBytes));
                }
            }
        }
        __syncthreads();
    }
}


// ---- the top-visible set rebuilt on the device: resetTopVisible, NJ.tcc:4728-4784 ------------------------------------
// The reference evaluates getVisible() for every live node in ascending order (its only side effect: the lazy refresh of the
// stale out-distances it touches), psorts the criteria and keeps the best entries that are not the reciprocal of one
// already kept.  Here: mark (who is a candidate, which stale nodes need a fresh value) -> k_nj_eval -> keys from the fresh
// values WITHOUT committing them (so that every thread sees the same state) -> commit -> top-K select -> one block applies
// the reference's sequential acceptance rule to the sorted candidates.
template<typename P>
__global__ void k_nj_rtv_mark(njl::State<P> st, int32_t *__restrict__ touchStamp) {
    Scalars *sc = st.sc;
    const int32_t i = (int32_t) (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= sc->maxnode || st.parent[i] >= 0) return;
    const int32_t vj = st.visJ[i];
    if (vj < 0 || st.parent[vj] >= 0) return;
    const int32_t nActive = sc->nActive, stamp = sc->stamp, epoch = sc->epoch;
    for (int t = 0; t < 2; t++) {
        const int32_t x = t ? vj : i;
        touchStamp[x] = stamp;
        if (njl::stale_fn(st.nOutAct[x], nActive, sc->staleOutLimit) && st.freshEpoch[x] != epoch && atomicExch(&st.wantStamp[x], stamp) != stamp) {
            const int q = atomicAdd(&sc->nOutReq, 1);
            st.reqOut[q] = x;                                   // capOut >= maxnodes
        }
    }
}
template<typename P>
__global__ void k_nj_rtv_keys(njl::State<P> st, uint64_t *__restrict__ keys) {
    Scalars *sc = st.sc;
    const int32_t i = (int32_t) (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= sc->maxnode) return;
    uint64_t k = ~0ull;
    if (st.parent[i] < 0) {
        const int32_t vj = st.visJ[i];
        if (vj >= 0 && st.parent[vj] < 0) { k = njl::okey(njl::crit_effective<P>(st, i, vj, st.visDist[i], sc->nActive)); atomicAdd(&sc->rtvVisible, 1); }
    }
    keys[i] = k;
}
template<typename P>
__global__ void k_nj_rtv_commit(njl::State<P> st, const int32_t *__restrict__ touchStamp) {
    Scalars *sc = st.sc;
    const int32_t x = (int32_t) (blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= sc->maxnode || touchStamp[x] != sc->stamp || st.parent[x] >= 0) return;
    if (njl::stale_fn(st.nOutAct[x], sc->nActive, sc->staleOutLimit)) { st.outDist[x] = st.freshVal[x]; st.nOutAct[x] = sc->nActive; }
}
constexpr int RTV_K = 2048, RTV_HS = 8192;
template<typename P>
__global__ void __launch_bounds__(1024)
k_nj_rtv_finish(njl::State<P> st, const Rec<P> *__restrict__ rec, const uint64_t *__restrict__ keys, int K) {
    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *eI = reinterpret_cast<int32_t *>(smem), *eJ = eI + RTV_K, *sI = eJ + RTV_K, *sJ = sI + RTV_K, *geZero = sJ + RTV_K;
    int32_t *hKey = geZero + RTV_K, *inTV = hKey + RTV_HS;
    __shared__ int32_t slot0;
    Scalars *sc = st.sc;
    const int tid = threadIdx.x;
    const uint64_t zeroKey = njl::okey((P) 0);
    for (int h = tid; h < RTV_HS; h += blockDim.x) { hKey[h] = -1; inTV[h] = -1; }
    __syncthreads();
    auto insert = [&](int32_t node) {
        uint32_t h = ((uint32_t) node * 2654435761u) & (RTV_HS - 1);
        for (;;) {
            const int32_t old = atomicCAS(&hKey[h], -1, node);
            if (old == -1 || old == node) return (int32_t) h;
            h = (h + 1) & (RTV_HS - 1);
        }
    };
    for (int k = tid; k < K; k += blockDim.x) {
        const int32_t i = (int32_t) rec[k].j;
        const uint64_t key = keys[i];
        const bool real = key != ~0ull;
        const int32_t vj = real ? st.visJ[i] : -1;
        eI[k] = real ? i : -1; eJ[k] = vj;
        geZero[k] = key >= zeroKey ? 1 : 0;
        if (real) { sI[k] = insert(i); sJ[k] = insert(vj); }
    }
    if (tid == 0) slot0 = insert(0);
    __syncthreads();
    if (tid == 0) {
        const int32_t nVisible = sc->rtvVisible, nTV = sc->nTV;
        const int32_t nAll = sc->nActive > nVisible ? sc->nActive : nVisible;
        int32_t tailLeft = nAll - nVisible, slots = 0, r = 0, iSave = 0;
        bool overflow = false;
        while (slots < nVisible && iSave < nTV) {
            int32_t vi, vj, s1, s2;
            const bool takeTail = tailLeft > 0 && (r >= nVisible || (r < K && geZero[r]));
            if (takeTail) { vi = 0; vj = 0; s1 = s2 = slot0; slots += tailLeft; tailLeft = 0; }      // the other value-initialised slots repeat (0, 0): skipped by the rule below
            else {
                if (r >= K) { overflow = true; break; }
                vi = eI[r]; vj = eJ[r]; s1 = sI[r]; s2 = sJ[r]; r++; slots++;
            }
            if (inTV[s1] != vj) { st.topvisible[iSave++] = vi; inTV[s1] = vj; inTV[s2] = vi; }
        }
        if (overflow) sc->status = njl::ST_NEED_HOST;
        else {
            while (iSave < nTV) st.topvisible[iSave++] = -1;
            sc->topvisibleAge = 0;
            sc->resume = njl::RS_SEARCH;
        }
    }
}

// ---- a top-hits refresh on the device: the refresh branch of topHitJoin, NJ.tcc:4439-4517 ---------------------------------
template<typename P>
__global__ void k_nj_commit_all(Store<P> s, njl::State<P> st) {
    Scalars *sc = st.sc;
    const int32_t i = (int32_t) (blockIdx.x * blockDim.x + threadIdx.x);
    if (i == 0) st.age[sc->jdNew] = 0;
    if (i >= sc->maxnode || !s.active[i]) return;
    st.outDist[i] = s.outDist[i]; st.nOutAct[i] = sc->nActive;
}
// the sorted 2m best hits of the new node -> its own list (sortSaveBestHits, :4472), the lists to merge and their offsets
template<typename P>
__global__ void __launch_bounds__(1024)
k_nj_refresh_self(njl::State<P> st, const Rec<P> *__restrict__ rec, int32_t *__restrict__ mNode, int32_t *__restrict__ mOff,
                  int32_t *__restrict__ allJ, P *__restrict__ allDist) {
    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *flag = reinterpret_cast<int32_t *>(smem);         // [2m]
    Scalars *sc = st.sc;
    const int32_t m = sc->m, newnode = sc->jdNew, n2 = 2 * m;
    for (int k = threadIdx.x; k < n2; k += blockDim.x) {
        const int32_t j = (int32_t) rec[k].j;
        allJ[k] = j; allDist[k] = rec[k].dist;
        flag[k] = j != newnode ? 1 : 0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n2; k += blockDim.x) {
        if (!flag[k]) continue;
        int pos = 0;
        for (int q = 0; q < k; q++) pos += flag[q];
        if (pos < m) { st.hitJ[(size_t) newnode * m + pos] = (int32_t) rec[k].j; st.hitDist[(size_t) newnode * m + pos] = rec[k].dist; }
    }
    if (threadIdx.x == 0) {
        int cnt = 0;
        for (int q = 0; q < n2; q++) cnt += flag[q];
        st.hitCount[newnode] = cnt < m ? cnt : m;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < m; l += blockDim.x) { const int32_t node = (int32_t) rec[l].j; mNode[l] = node; flag[l] = st.hitCount[node]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int l = 0; l < m; l++) { mOff[l] = off; off += flag[l]; }
        mOff[m] = off;
    }
}
template<typename P>
__global__ void k_nj_refresh_pack(njl::State<P> st, const int32_t *__restrict__ mNode, const int32_t *__restrict__ mOff,
                                  int32_t *__restrict__ ownJ, P *__restrict__ ownDist) {
    const int l = blockIdx.x, m = st.sc->m;
    const int32_t node = mNode[l], off = mOff[l], cnt = mOff[l + 1] - off;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        const int32_t j0 = st.hitJ[(size_t) node * m + k], j = njl::ancestor_fn(st.up, j0);      // updateBestHit(.., false), :1626-1648
        ownJ[off + k] = j < 0 ? -1 : j;
        ownDist[off + k] = j == j0 ? st.hitDist[(size_t) node * m + k] : (P) -1e20;
    }
}
template<typename P>
__global__ void k_nj_refresh_unpack(njl::State<P> st, const int32_t *__restrict__ mNode, const int32_t *__restrict__ outCount,
                                    const int32_t *__restrict__ outJ, const P *__restrict__ outDist) {
    const int l = blockIdx.x, m = st.sc->m;
    const int32_t node = mNode[l], cnt = outCount[l];
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        st.hitJ[(size_t) node * m + k] = outJ[(size_t) l * m + k];
        st.hitDist[(size_t) node * m + k] = outDist[(size_t) l * m + k];
    }
    if (threadIdx.x == 0) {
        st.hitCount[node] = cnt; st.age[node] = 0;
        st.visJ[node] = outJ[(size_t) l * m]; st.visDist[node] = outDist[(size_t) l * m];          // :4513
    }
}

struct LoopArrays {       // device allocations of one loop
    void *sc = nullptr, *parent = nullptr, *up = nullptr, *child = nullptr, *branchlength = nullptr, *diameter = nullptr, *outDist = nullptr, *nOutAct = nullptr,
         *freshVal = nullptr, *freshEpoch = nullptr, *wantStamp = nullptr, *hitJ = nullptr, *hitDist = nullptr, *hitCount = nullptr, *age = nullptr, *visJ = nullptr,
         *visDist = nullptr, *topvisible = nullptr, *reqOut = nullptr, *reqA = nullptr, *reqB = nullptr, *pairD = nullptr, *pairW = nullptr, *uJ = nullptr, *uSlot = nullptr,
         *lSlot = nullptr, *tmpP = nullptr, *joins = nullptr, *touchStamp = nullptr, *rec = nullptr, *mNode = nullptr, *mOff = nullptr,
         *allJ = nullptr, *allDist = nullptr, *ownJ = nullptr, *ownDist = nullptr, *outCount = nullptr, *outJ = nullptr, *outDistL = nullptr, *acct = nullptr;
};

}  // namespace

struct vftx_loop {
    vft_ctx *c;
    vft_nj_options opt;
    int64_t m, nTV, cap, capOut, capPair;
    LoopArrays d;
    Scalars *hSc = nullptr;               // pinned mirror of the scalars
    std::vector<int64_t> hJoins;          // joins noted so far (i, j)
    int64_t nNoted = 0;                   // joins already replayed into the context's host mirrors
    int64_t traceBase = -1;               // trace index of the loop's first join (nSeqs - nActive at the first upload)
    int64_t nSteps = 0;
    int64_t seenSeqOps = 0, seenProfileOps = 0, seenOutOps = 0, seenBytes = 0;
    unsigned long long seenAcct[4] = {0, 0, 0, 0};
    int64_t nDevReset = 0, nDevRefresh = 0;
    bool wide = false;
};

template<typename P>
static njl::State<P> loop_state(vftx_loop *lp) {
    njl::State<P> st;
    const LoopArrays &d = lp->d;
    st.sc = (Scalars *) d.sc; st.parent = (int32_t *) d.parent; st.up = (int32_t *) d.up; st.child = (int32_t *) d.child;
    st.branchlength = (P *) d.branchlength; st.diameter = (P *) d.diameter; st.outDist = (P *) d.outDist; st.nOutAct = (int32_t *) d.nOutAct;
    st.freshVal = (P *) d.freshVal; st.freshEpoch = (int32_t *) d.freshEpoch; st.wantStamp = (int32_t *) d.wantStamp;
    st.hitJ = (int32_t *) d.hitJ; st.hitDist = (P *) d.hitDist; st.hitCount = (int32_t *) d.hitCount; st.age = (int32_t *) d.age;
    st.visJ = (int32_t *) d.visJ; st.visDist = (P *) d.visDist; st.topvisible = (int32_t *) d.topvisible;
    st.reqOut = (int32_t *) d.reqOut; st.reqA = (int32_t *) d.reqA; st.reqB = (int32_t *) d.reqB; st.pairD = (P *) d.pairD; st.pairW = (P *) d.pairW;
    st.uJ = (int32_t *) d.uJ; st.uSlot = (int32_t *) d.uSlot; st.lSlot = (int32_t *) d.lSlot; st.tmpP = (P *) d.tmpP; st.joins = (int64_t *) d.joins;
    st.capOut = (int32_t) lp->capOut; st.capPair = (int32_t) lp->capPair;
    return st;
}

extern "C" int vftx_loop_destroy(vftx_loop *lp) {
    if (!lp) return VFT_OK;
    bind_device(lp->c);
    cudaStreamSynchronize(lp->c->stream);
    void **p = reinterpret_cast<void **>(&lp->d);
    for (size_t k = 0; k < sizeof(LoopArrays) / sizeof(void *); k++) mem_free(p[k]);
    mem_free(lp->hSc);
    delete lp;
    return VFT_OK;
}

extern "C" int vftx_loop_create(vft_ctx *c, const vft_nj_options *opt, int64_t m, int64_t nTV, vftx_loop **out) {
    if (!c || !opt || !out || m < 4 || nTV < 1) return fail(VFT_EINVAL, "bad argument");
    bind_device(c);
    vftx_loop *lp = new vftx_loop();
    lp->c = c; lp->opt = *opt; lp->m = m; lp->nTV = nTV;
    int64_t cap = 32;
    while (cap < 2 * m) cap <<= 1;
    lp->cap = cap; lp->capOut = std::max<int64_t>(4 * cap + 2 * nTV + 8, c->M); lp->capPair = cap + 2 * m + 8;
    const size_t M = (size_t) c->M, ps = c->ps;
    LoopArrays &d = lp->d;
    struct { void **p; size_t bytes; int fill; } allocs[] = {
        {&d.sc, sizeof(Scalars), 0}, {&d.parent, M * 4, 0}, {&d.up, M * 4, 0}, {&d.child, 3 * M * 4, 0}, {&d.branchlength, M * ps, 0}, {&d.diameter, M * ps, 0},
        {&d.outDist, M * ps, 0}, {&d.nOutAct, M * 4, 0}, {&d.freshVal, M * ps, 0}, {&d.freshEpoch, M * 4, 0xFF}, {&d.wantStamp, M * 4, 0xFF},
        {&d.hitJ, M * (size_t) m * 4, 0xFF}, {&d.hitDist, M * (size_t) m * ps, 0}, {&d.hitCount, M * 4, 0}, {&d.age, M * 4, 0}, {&d.visJ, M * 4, 0xFF},
        {&d.visDist, M * ps, 0}, {&d.topvisible, (size_t) nTV * 4, 0xFF}, {&d.reqOut, (size_t) lp->capOut * 4, 0xFF}, {&d.reqA, (size_t) lp->capPair * 4, 0xFF},
        {&d.reqB, (size_t) lp->capPair * 4, 0xFF}, {&d.pairD, (size_t) lp->capPair * ps, 0}, {&d.pairW, (size_t) lp->capPair * ps, 0}, {&d.uJ, (size_t) cap * 4, 0xFF},
        {&d.uSlot, (size_t) cap * 4, 0xFF}, {&d.lSlot, (size_t) 2 * m * 4, 0xFF}, {&d.tmpP, (size_t) cap * ps, 0}, {&d.joins, M * 8, 0xFF},
        {&d.touchStamp, M * 4, 0xFF}, {&d.rec, (size_t) SEL_MAXK * sizeof(Rec<double>), 0}, {&d.mNode, (size_t) m * 4, 0}, {&d.mOff, (size_t) (m + 1) * 4, 0},
        {&d.allJ, (size_t) 2 * m * 4, 0}, {&d.allDist, (size_t) 2 * m * ps, 0}, {&d.ownJ, (size_t) m * m * 4, 0}, {&d.ownDist, (size_t) m * m * ps, 0},
        {&d.outCount, (size_t) m * 4, 0}, {&d.outJ, (size_t) m * m * 4, 0}, {&d.outDistL, (size_t) m * m * ps, 0}, {&d.acct, 64, 0}};
    for (auto &a : allocs) {
        if (mem_alloc(a.p, a.bytes, MEM_DEVICE) != cudaSuccess || cudaMemsetAsync(*a.p, a.fill, a.bytes, c->stream) != cudaSuccess) {
            vftx_loop_destroy(lp);
            return fail(VFT_ENOMEM, "device memory for the join loop");
        }
    }
    if (mem_alloc((void **) &lp->hSc, sizeof(Scalars), MEM_PINNED) != cudaSuccess) { vftx_loop_destroy(lp); return fail(VFT_ENOMEM, "pinned memory"); }
    Scalars &sc = *lp->hSc;
    std::memset(&sc, 0, sizeof sc);
    sc.nSeqs = c->N; sc.maxnodes = c->M; sc.m = (int32_t) m; sc.nTV = (int32_t) nTV; sc.cap = (int32_t) cap;
    sc.tophitAgeLimit = (int32_t) std::max<int64_t>(1, (int64_t) (0.5 + std::log((double) m) / std::log(2.0)));
    sc.nRefreshMin = (int32_t) (int64_t) (0.5 + m * opt->tophitsRefresh);
    sc.nResetOutProfile = opt->nResetOutProfile; sc.staleOutLimit = opt->staleOutLimit; sc.fResetOutProfile = opt->fResetOutProfile;
    sc.Lbytes = c->L; sc.profBytes = c->L * ((int64_t) c->A * (int64_t) ps + (int64_t) ps + 1);
    sc.status = njl::ST_RUNNING; sc.resume = njl::RS_SEARCH; sc.epoch = 1; sc.stamp = 1; sc.hintEpoch = -1; sc.hintJoinSlot = -1;
    lp->wide = c->wideOk && c->Lp >= 512;
    // shared memory of the step kernel: the logic's scratch + one term tile per warp for the distances it computes itself
#define SET_STEP_SMEM(P, A_, MX) do { const size_t need = ((njl::scratch_bytes<P>((int) cap, NJ_STEP_T) + 15) & ~(size_t) 15) + (NJ_STEP_T / 32) * group_smem_bytes<P, A_, MX>(1); \
        if (need > 200 * 1024) { vftx_loop_destroy(lp); return fail(VFT_EINVAL, "top-hit lists too long for the step kernel's scratch"); } \
        cudaFuncSetAttribute(k_nj_step<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) need); \
        cudaFuncSetAttribute(k_nj_eval<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (4 * group_smem_bytes<P, A_, MX>(TileShape<A_, MX>::R))); \
        if (lp->wide && wide_smem_bytes<P, A_, MX>(c->Lp) > 48 * 1024) cudaFuncSetAttribute(k_nj_eval_wide<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) wide_smem_bytes<P, A_, MX>(c->Lp)); } while (0)
    VFT_DISPATCH(c, SET_STEP_SMEM);
    *out = lp;
    return VFT_OK;
}

extern "C" int vftx_loop_upload(vftx_loop *lp, const vftx_loop_image *g, int32_t resume) {
    if (!lp || !g) return fail(VFT_EINVAL, "null argument");
    vft_ctx *c = lp->c;
    bind_device(c);
    const size_t M = (size_t) c->M, ps = c->ps, m = (size_t) lp->m;
    LoopArrays &d = lp->d;
    CK(cudaMemcpyAsync(d.parent, g->parent, M * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.up, g->up, M * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.child, g->child, 3 * M * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.branchlength, g->branchlength, M * ps, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.diameter, g->diameter, M * ps, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.outDist, g->outDist, M * ps, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.nOutAct, g->nOutAct, M * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.hitJ, g->hitJ, M * m * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.hitDist, g->hitDist, M * m * ps, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.hitCount, g->hitCount, M * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.age, g->age, M * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.visJ, g->visJ, M * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.visDist, g->visDist, M * ps, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d.topvisible, g->topvisible, (size_t) lp->nTV * 4, cudaMemcpyHostToDevice, c->stream));
    c->cnt.h2dBytes += (int64_t) (M * (7 * 4 + 4 * ps) + M * m * (4 + ps));
    Scalars &sc = *lp->hSc;
    if (lp->traceBase < 0) lp->traceBase = c->N - g->nActive;
    sc.maxnode = (int32_t) g->maxnode; sc.nActive = (int32_t) g->nActive; sc.topvisibleAge = (int32_t) g->topvisibleAge;
    sc.nActiveOutProfileReset = (int32_t) g->nActiveOutProfileReset; sc.totdiam = g->totdiam;
    sc.status = njl::ST_RUNNING; sc.resume = resume; sc.epoch++; sc.hintEpoch = -1; sc.hintJoinSlot = -1; sc.jdValid = 0; sc.jdSelfPending = 0; sc.nOutReq = 0; sc.nPairReq = 0;
    CK(cudaMemcpyAsync(d.sc, &sc, sizeof sc, cudaMemcpyHostToDevice, c->stream));
    CK(sync_stream(c));
    return VFT_OK;
}

extern "C" int vftx_loop_download(vftx_loop *lp, vftx_loop_image *g) {
    if (!lp || !g) return fail(VFT_EINVAL, "null argument");
    vft_ctx *c = lp->c;
    bind_device(c);
    const size_t M = (size_t) c->M, ps = c->ps, m = (size_t) lp->m;
    LoopArrays &d = lp->d;
    CK(cudaMemcpyAsync(g->parent, d.parent, M * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->up, d.up, M * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->child, d.child, 3 * M * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->branchlength, d.branchlength, M * ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->diameter, d.diameter, M * ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->outDist, d.outDist, M * ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->nOutAct, d.nOutAct, M * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->hitJ, d.hitJ, M * m * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->hitDist, d.hitDist, M * m * ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->hitCount, d.hitCount, M * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->age, d.age, M * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->visJ, d.visJ, M * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->visDist, d.visDist, M * ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g->topvisible, d.topvisible, (size_t) lp->nTV * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(lp->hSc, d.sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    CK(sync_stream(c));
    c->cnt.d2hBytes += (int64_t) (M * (7 * 4 + 4 * ps) + M * m * (4 + ps));
    const Scalars &sc = *lp->hSc;
    g->maxnode = sc.maxnode; g->nActive = sc.nActive; g->topvisibleAge = sc.topvisibleAge; g->nActiveOutProfileReset = sc.nActiveOutProfileReset;
    g->totdiam = sc.totdiam;
    return VFT_OK;
}

// the joins the device made since the last call, replayed into the context's host-side mirrors (active flags, maxnode,
// counters) that the ordinary entry points validate and account against
static int loop_note_joins(vftx_loop *lp) {
    vft_ctx *c = lp->c;
    const Scalars &sc = *lp->hSc;
    const int64_t nNow = sc.nJoins;
    if (nNow <= lp->nNoted) return VFT_OK;
    const int64_t first = lp->traceBase + lp->nNoted;                               // trace index of the first new join
    const int64_t cnt = nNow - lp->nNoted;
    lp->hJoins.resize((size_t) (2 * nNow));
    CK(cudaMemcpyAsync(lp->hJoins.data() + 2 * lp->nNoted, (int64_t *) lp->d.joins + 2 * first, (size_t) cnt * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(sync_stream(c));
    for (int64_t k = lp->nNoted; k < nNow; k++) {
        const int64_t newnode = c->maxnode;
        for (int64_t ch : {lp->hJoins[(size_t) (2 * k)], lp->hJoins[(size_t) (2 * k + 1)]})
            if (c->activeHost[ch]) { c->activeHost[ch] = 0; if (ch < c->N) c->nActLeaf--; else c->nActInternal--; }
        c->activeHost[newnode] = 1; c->nActInternal++; c->actDirty = true;
        c->maxnode = newnode + 1;
        c->cnt.profileAvgOps++;
    }
    lp->nNoted = nNow;
    return VFT_OK;
}

extern "C" int64_t vftx_loop_joins(vftx_loop *lp, int64_t *out, int64_t maxJoins) {
    if (!lp || !out) return 0;
    vft_ctx *c = lp->c;
    bind_device(c);
    const int64_t n = std::min<int64_t>(maxJoins, c->N - 3);
    if (cudaMemcpyAsync(out, lp->d.joins, (size_t) n * 16, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || sync_stream(c) != cudaSuccess) return 0;
    return n;
}

static int loop_write_scalars(vftx_loop *lp) {
    CK(cudaMemcpyAsync(lp->d.sc, lp->hSc, sizeof(Scalars), cudaMemcpyHostToDevice, lp->c->stream));
    return VFT_OK;
}

// launch geometry shared by the loop's launches
struct LoopGeom {
    unsigned evalBlocks, wideBlocks, avgBlocks;
    int AVG_T;
    size_t smemAvg;
};
static LoopGeom loop_geom(vftx_loop *lp) {
    vft_ctx *c = lp->c;
    LoopGeom g;
    const size_t evalItems = (size_t) (lp->capOut + lp->capPair);
    g.evalBlocks = (unsigned) std::min<size_t>((evalItems + 3) / 4, 148 * 8);
    g.wideBlocks = (unsigned) std::min<size_t>(evalItems, 2048);
    g.AVG_T = c->Lp <= 512 ? 256 : 64;
    g.avgBlocks = c->Lp <= 512 ? 1u : (unsigned) ((c->Lp + g.AVG_T - 1) / g.AVG_T);
    g.smemAvg = (size_t) c->Lp * 16;
    return g;
}
#define LOOP_STEP(P, A_, MX) k_nj_step<P, A_, MX><<<1, NJ_STEP_T, ((njl::scratch_bytes<P>((int) lp->cap, NJ_STEP_T) + 15) & ~(size_t) 15) + (NJ_STEP_T / 32) * group_smem_bytes<P, A_, MX>(1), c->stream>>>(make_store<P>(c), loop_state<P>(lp), hintOnly)
#define LOOP_AVG(P, A_, MX) do { if (g.smemAvg > 48 * 1024) cudaFuncSetAttribute(k_average<P, A_, MX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) g.smemAvg); \
        k_average<P, A_, MX, true><<<g.avgBlocks, g.AVG_T, g.smemAvg, c->stream>>>(make_store<P>(c), 0, 0, 0, 0.5, (P) 0, 0, c->d_terms, c->d_doneCount, (P *) c->ow, (P *) c->ov, (P *) c->ocd, (P *) nullptr, (Scalars *) lp->d.sc); } while (0)
// the request list: one CTA per item while the list is short and the rows are long, G items per warp otherwise (each kernel
// returns at once outside its regime: the list size is only known on the device)
#define LOOP_EVAL(P, A_, MX) do { if (lp->wide) k_nj_eval_wide<P, A_, MX><<<g.wideBlocks, 256, wide_smem_bytes<P, A_, MX>(c->Lp), c->stream>>>(make_store<P>(c), loop_state<P>(lp), 2048); \
        k_nj_eval<P, A_, MX><<<g.evalBlocks, 128, 4 * group_smem_bytes<P, A_, MX>(TileShape<A_, MX>::R), c->stream>>>(make_store<P>(c), loop_state<P>(lp), lp->wide ? 2048 : -1); } while (0)

// resetTopVisible on the device (kernels above); the loop's scalars in lp->hSc are current
static int loop_reset_topvisible(vftx_loop *lp) {
    vft_ctx *c = lp->c;
    Scalars &sc = *lp->hSc;
    const LoopGeom g = loop_geom(lp);
    sc.status = njl::ST_RUNNING; sc.nOutReq = 0; sc.nPairReq = 0; sc.stamp++; sc.rtvVisible = 0; sc.visfixPending = 0;
    int rc = loop_write_scalars(lp); if (rc) return rc;
    const int64_t n = sc.maxnode;
    const unsigned nb = (unsigned) ((n + 255) / 256);
    const int K = (int) std::min<int64_t>(n, RTV_K);
    const size_t finSmem = (size_t) RTV_K * 20 + (size_t) RTV_HS * 8;
#define LOOP_RTV(P, A_, MX) do { \
        prof_begin(c, CLS_SELECT, K_NJ_STEP); k_nj_rtv_mark<P><<<nb, 256, 0, c->stream>>>(loop_state<P>(lp), (int32_t *) lp->d.touchStamp); prof_end(c); \
        prof_begin(c, CLS_DIST, K_EVAL_LARGE); LOOP_EVAL(P, A_, MX); prof_end(c); \
        prof_begin(c, CLS_SELECT, K_NJ_STEP); k_nj_rtv_keys<P><<<nb, 256, 0, c->stream>>>(loop_state<P>(lp), c->d_keys); \
        k_nj_rtv_commit<P><<<nb, 256, 0, c->stream>>>(loop_state<P>(lp), (const int32_t *) lp->d.touchStamp); prof_end(c); \
        prof_begin(c, CLS_SELECT, K_SELECT); launch_topk<P, (int) sizeof(P)>(c, c->d_keys, n, K, (Rec<P> *) lp->d.rec); prof_end(c); \
        cudaFuncSetAttribute(k_nj_rtv_finish<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) finSmem); \
        prof_begin(c, CLS_SELECT, K_NJ_STEP); k_nj_rtv_finish<P><<<1, 1024, finSmem, c->stream>>>(loop_state<P>(lp), (const Rec<P> *) lp->d.rec, c->d_keys, K); prof_end(c); } while (0)
    VFT_DISPATCH(c, LOOP_RTV);
    CK(cudaGetLastError());
    c->cnt.launches += lp->wide ? 7 : 6;
    lp->nDevReset++;
    return VFT_OK;
}

// the refresh branch of topHitJoin on the device; false in *done when this one is left to the host (the end game: fewer
// than 2m active nodes, or lists too long for the merge kernels' shared memory)
static int loop_refresh(vftx_loop *lp, bool *done) {
    vft_ctx *c = lp->c;
    Scalars &sc = *lp->hSc;
    const LoopGeom g = loop_geom(lp);
    const int64_t m = lp->m, nActive = sc.nActive, n = sc.maxnode, newnode = sc.jdNew;
    const int64_t cap = 3 * m;
    int np2 = 32;
    while (np2 < cap) np2 <<= 1;
    *done = false;
    if (nActive < 2 * m || np2 > MRG_MAX || 2 * m > SEL_MAXK) return VFT_OK;
    const size_t ps = c->ps;
    const size_t slots = (size_t) m * cap;
    const size_t need = slots * (12 + 3 * ps) + (size_t) m * 4 + 64;
    if (need > c->mrgCap) {
        mem_free(c->d_mrg);
        c->d_mrg = nullptr; c->mrgCap = 0;
        CK(mem_alloc((void **) &c->d_mrg, need * 2, MEM_DEVICE));
        c->mrgCap = need * 2;
    }
    char *dm = (char *) c->d_mrg;
    void *uD = dm, *r0 = dm + slots * ps, *r1 = dm + 2 * slots * ps;
    int32_t *uJ = (int32_t *) (dm + 3 * slots * ps), *reqA = uJ + slots, *reqB = reqA + slots, *cnt = reqB + slots;
    const unsigned nb = (unsigned) ((n + 255) / 256);
    const int Gall = pick_group(c, n);
    const int64_t warpsAll = (n + Gall - 1) / Gall;
    const int Gm = pick_group(c, (int64_t) slots);
    const int64_t warpsM = ((int64_t) slots + Gm - 1) / Gm;
    const size_t smemSort = (size_t) np2 * 12;
    InlineItems inl;
    inl.a[0] = 0;
    LoopArrays &d = lp->d;
#define LOOP_REFRESH(P, A_, MX) do { \
        prof_begin(c, CLS_DIST, K_OUT_DIST_ALL); \
        if (c->stagedOk) k_out_distance_all_staged<float, 20, true><<<148, STG_T, c->stagedSmem, c->stream>>>(make_store<float>(c), n, std::min(Gall, STG_G), nActive, sc.totdiam); \
        else k_out_distance_all<P, A_, MX><<<(unsigned) ((warpsAll + 3) / 4), 128, 4 * group_smem_bytes<P, A_, MX>(Gall), c->stream>>>(make_store<P>(c), n, Gall, nActive, sc.totdiam); \
        prof_end(c); \
        k_nj_commit_all<P><<<nb, 256, 0, c->stream>>>(make_store<P>(c), loop_state<P>(lp)); \
        prof_begin(c, CLS_DIST, K_ONE_VS_ALL); \
        if (c->stagedOk) k_one_vs_all_staged<float, 20, true><<<148, STG_T, c->stagedSmem, c->stream>>>(make_store<float>(c), newnode, nActive, n, 0, n, std::min(Gall, STG_G), (float *) c->d_dist, (float *) c->d_weight, (float *) c->d_crit, c->d_keys); \
        else k_one_vs_all_warp<P, A_, MX><<<(unsigned) ((warpsAll + 3) / 4), 128, 4 * group_smem_bytes<P, A_, MX>(Gall), c->stream>>>(make_store<P>(c), newnode, nActive, n, 0, n, Gall, (P *) c->d_dist, (P *) c->d_weight, (P *) c->d_crit, c->d_keys); \
        prof_end(c); \
        prof_begin(c, CLS_SELECT, K_SELECT); launch_topk<P, (int) sizeof(P)>(c, c->d_keys, n, (int) (2 * m), (Rec<P> *) d.rec); prof_end(c); \
        prof_begin(c, CLS_SELECT, K_MERGE); \
        k_nj_refresh_self<P><<<1, 1024, (size_t) 2 * m * 4, c->stream>>>(loop_state<P>(lp), (const Rec<P> *) d.rec, (int32_t *) d.mNode, (int32_t *) d.mOff, (int32_t *) d.allJ, (P *) d.allDist); \
        k_nj_refresh_pack<P><<<(unsigned) m, 128, 0, c->stream>>>(loop_state<P>(lp), (const int32_t *) d.mNode, (const int32_t *) d.mOff, (int32_t *) d.ownJ, (P *) d.ownDist); \
        cudaFuncSetAttribute(k_merge_prep<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, MRG_MAX * 12); \
        cudaFuncSetAttribute(k_merge_finish<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, MRG_MAX * 12); \
        k_merge_prep<P><<<(unsigned) m, MRG_T, smemSort, c->stream>>>((const int32_t *) d.mNode, (const int32_t *) d.mOff, (const int32_t *) d.ownJ, (const P *) d.ownDist, (int) (2 * m), \
            (const int32_t *) d.allJ, (const P *) d.allDist, (int) newnode, (int) cap, np2, (int) c->N, uJ, (P *) uD, reqA, reqB, cnt, (unsigned long long *) d.acct); prof_end(c); \
        prof_begin(c, CLS_DIST, K_EVAL_LARGE); \
        k_eval<P, A_, MX, true><<<(unsigned) ((warpsM + 3) / 4), 128, 4 * group_smem_bytes<P, A_, MX>(Gm), c->stream>>>(make_store<P>(c), inl, reqA, reqB, (int64_t) slots, 0, Gm, 0, nActive, 0.0, (P *) r0, (P *) r1, c->d_doneCount, (P *) nullptr); prof_end(c); \
        prof_begin(c, CLS_SELECT, K_MERGE); \
        k_merge_finish<P><<<(unsigned) m, MRG_T, smemSort, c->stream>>>(make_store<P>(c), (const int32_t *) d.mNode, nActive, (int) m, (int) cap, np2, uJ, (P *) uD, reqA, (const P *) r0, cnt, (int32_t *) d.outCount, (int32_t *) d.outJ, (P *) d.outDistL); \
        k_nj_refresh_unpack<P><<<(unsigned) m, 128, 0, c->stream>>>(loop_state<P>(lp), (const int32_t *) d.mNode, (const int32_t *) d.outCount, (const int32_t *) d.outJ, (const P *) d.outDistL); prof_end(c); } while (0)
    VFT_DISPATCH(c, LOOP_REFRESH);
    CK(cudaGetLastError());
    c->cnt.launches += 10;
    // accounting of the two sweeps (one profile distance per active node each); the merge's own comes from its counters
    const int64_t leafB = c->L, profB = c->L * ((int64_t) c->A * (int64_t) ps + (int64_t) ps + 1);
    const int64_t sweep = c->nActLeaf * leafB + c->nActInternal * profB + profB;
    c->cnt.profileOps += 2 * nActive; c->cnt.outprofileOps += nActive; c->cnt.algoBytes += 2 * sweep;
    c->cnt.bytesKernel[K_OUT_DIST_ALL] += sweep; c->cnt.bytesKernel[K_ONE_VS_ALL] += sweep;
    lp->nDevRefresh++;
    *done = true;
    return loop_reset_topvisible(lp);
}

extern "C" int vftx_loop_run(vftx_loop *lp, vftx_loop_status *out) {
    if (!lp) return fail(VFT_EINVAL, "null argument");
    vft_ctx *c = lp->c;
    bind_device(c);
    static const int ROUND = [] { const char *e = std::getenv("VFT_LOOP_ROUND"); return e ? std::max(1, std::atoi(e)) : 8; }();
    static const bool deviceDetours = [] { const char *e = std::getenv("VFT_LOOP_DETOURS"); return !e || e[0] != '0'; }();
    const LoopGeom g = loop_geom(lp);
    if (g.smemAvg > 200 * 1024) return fail(VFT_EINVAL, "alignment too long for the average kernel's term buffer");
    Scalars &sc = *lp->hSc;
    int hintOnly = 0;
    auto launchHint = [&] {
        hintOnly = 1;
        VFT_DISPATCH(c, LOOP_STEP);
        VFT_DISPATCH(c, LOOP_EVAL);
        hintOnly = 0;
        c->cnt.launches += lp->wide ? 3 : 2;
    };
    if (sc.hintEpoch != sc.epoch && sc.resume == njl::RS_SEARCH && sc.status == njl::ST_RUNNING) launchHint();   // a fresh image: the hints of the first search
    for (;;) {
        for (int r = 0; r < ROUND; r++) {
            prof_begin(c, CLS_SELECT, K_NJ_STEP); VFT_DISPATCH(c, LOOP_STEP); prof_end(c);
            prof_begin(c, CLS_PROFILE, K_AVERAGE); VFT_DISPATCH(c, LOOP_AVG); prof_end(c);
            prof_begin(c, CLS_DIST, K_EVAL_SMALL); VFT_DISPATCH(c, LOOP_EVAL); prof_end(c);
        }
        c->cnt.launches += (lp->wide ? 4 : 3) * ROUND;
        lp->nSteps += ROUND;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&sc, lp->d.sc, sizeof sc, cudaMemcpyDeviceToHost, c->stream));
        CK(sync_stream(c));
        int rc = loop_note_joins(lp);
        if (rc != VFT_OK) return rc;
        if (sc.status == njl::ST_NEED_REBUILD) {
            // NJ.tcc:3012-3033: the out-profile from scratch; totdiam summed in node order as the reference does
            std::vector<char> diam((size_t) c->maxnode * c->ps);
            CK(cudaMemcpyAsync(diam.data(), lp->d.diameter, diam.size(), cudaMemcpyDeviceToHost, c->stream));
            CK(sync_stream(c));
            double td = 0;
            for (int64_t i = 0; i < c->maxnode; i++)
                if (c->activeHost[i]) td += c->ps == 4 ? (double) ((const float *) diam.data())[i] : ((const double *) diam.data())[i];
            rc = vft_outprofile_rebuild(c, nullptr, sc.nActive);
            if (rc != VFT_OK) return rc;
            sc.totdiam = td; sc.nActiveOutProfileReset = sc.nActive; sc.status = njl::ST_RUNNING;
            rc = loop_write_scalars(lp); if (rc) return rc;
            VFT_DISPATCH(c, LOOP_EVAL);
            c->cnt.launches += lp->wide ? 2 : 1;
            continue;
        }
        if (deviceDetours && sc.status == njl::ST_NEED_RESET && !sc.visfixPending) {
            rc = loop_reset_topvisible(lp); if (rc) return rc;
            launchHint();
            continue;
        }
        if (deviceDetours && sc.status == njl::ST_NEED_REFRESH) {
            bool done = false;
            rc = loop_refresh(lp, &done); if (rc) return rc;
            if (done) { launchHint(); continue; }
        }
        if (sc.status == njl::ST_NEED_HOST) { sc.status = njl::ST_NEED_RESET; sc.visfixPending = 0; }      // the device rebuild ran out of sorted candidates
        if (sc.status != njl::ST_RUNNING) break;
    }
    if (out) {
        out->status = sc.status; out->resume = sc.resume; out->visfixPending = sc.visfixPending; out->newnode = sc.jdNew;
        out->nActive = sc.nActive; out->maxnode = sc.maxnode; out->nJoins = sc.nJoins; out->nRefresh = sc.nRefresh;
        out->nVisibleUpdate = sc.nVisibleUpdate; out->nHillBetter = sc.nHillBetter; out->nReset = sc.nReset; out->nInlineOut = sc.nInlineOut;
        out->nInlinePair = sc.nInlinePair; out->nPairHit = sc.nPairHit; out->nRebuild = sc.nRebuild; out->nSteps = lp->nSteps;
    }
    if (std::getenv("VFT_LOOP_TIMING") && sc.status != njl::ST_RUNNING) {
        static const char *nm[12] = {"thjFinish: resolvePairs", "thjFinish: ensureCommit", "thjFinish: crit+sort+save", "updateTopVisible(new)", "updateVisible: flags",
                                     "updateVisible: replacements", "search: scan", "search: hill-climb (getBest x2)", "(search exit)", "joinBookkeeping", "thjPrepare", "hintSearch"};
        double tot = 0;
        for (int k = 0; k < 12; k++) tot += (double) sc.tPhase[k];
        std::fprintf(stderr, "[k_nj_step phases] joins %lld, cycles per join by phase (tid 0's clock64):\n", (long long) sc.nJoins);
        for (int k = 0; k < 12; k++) std::fprintf(stderr, "   %-34s %9.0f  %5.1f %%\n", nm[k], (double) sc.tPhase[k] / (double) std::max<int64_t>(1, sc.nJoins), 100.0 * (double) sc.tPhase[k] / std::max(1.0, tot));
    }
    // the loop's own distance work, for the context's counters (Debug.h:12-15)
    c->cnt.seqOps += sc.seqOps - lp->seenSeqOps; c->cnt.profileOps += (sc.profileOps + sc.outprofileOps) - lp->seenProfileOps;
    c->cnt.outprofileOps += sc.outprofileOps - lp->seenOutOps; c->cnt.algoBytes += sc.algoBytes - lp->seenBytes;
    c->cnt.bytesKernel[K_EVAL_SMALL] += sc.algoBytes - lp->seenBytes;
    lp->seenSeqOps = sc.seqOps; lp->seenProfileOps = sc.profileOps + sc.outprofileOps; lp->seenOutOps = sc.outprofileOps; lp->seenBytes = sc.algoBytes;
    if (lp->nDevRefresh > 0) {      // the merges of the device-side refreshes (k_merge_prep's counters, cumulative)
        unsigned long long acct[4];
        CK(cudaMemcpyAsync(acct, lp->d.acct, 32, cudaMemcpyDeviceToHost, c->stream));
        CK(sync_stream(c));
        const int64_t profB = c->L * ((int64_t) c->A * (int64_t) c->ps + (int64_t) c->ps + 1);
        const int64_t by = (int64_t) (acct[2] - lp->seenAcct[2]) * c->L + (int64_t) (acct[3] - lp->seenAcct[3]) * profB;
        c->cnt.seqOps += (int64_t) (acct[0] - lp->seenAcct[0]); c->cnt.profileOps += (int64_t) (acct[1] - lp->seenAcct[1]);
        c->cnt.algoBytes += by; c->cnt.bytesKernel[K_EVAL_LARGE] += by;
        for (int k = 0; k < 4; k++) lp->seenAcct[k] = acct[k];
    }
    return VFT_OK;
}
