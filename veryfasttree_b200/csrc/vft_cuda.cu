// vft_cuda.cu -- the CUDA (sm_100a) implementation of the C-ABI in include/vft_b200.h.
//
// Data layout in HBM (one context = one GPU):
//   codes    uint8 [2N][Lp]        every node; leaves are nothing else (1 B/position)
//   weights  P     [N][Lp]         internal nodes only (row = id - N)
//   vecs     P     [N][Lp][A]      internal nodes only, dense (unused where the code is known)
//   out-profile ow[Lp], ov[Lp][A], ocd[Lp][A]; per-node scalars diameter/selfdist/selfweight/outDist/active
// Lp = nPos rounded up to 32; rows are 16-byte aligned and read with 128-bit loads.
//
// Kernels (the ordered accumulation they share is in vft_device.cuh):
//   k_eval / k_eval_wide  candidate lists + lazy out-distances (transferBestHits / uniqueBestHits / getBestFromTopHits):
//                         up to R pairs per warp, or one CTA per pair for the per-join lists of long alignments
//   k_sweep20 (vft_sweep.cuh)  the three all-candidate sweeps of the 20-state matrix mode (setBestHit, the all-node out-distances,
//                         the list merges of a refresh): per-query tables, producer warps + a consumer warp per 32 candidates
//   k_one_vs_all_*        query vs every active node + criterion -> 64-bit sort keys   (setBestHit; the other modes)
//   k_topk_select         radix select + bitonic sort: the K best in the reference's psort order
//   k_merge_prep/finish   the m list merges of a top-hits refresh (vft_tophits_merge)
//   k_out_distance_all    profileDist(node, out-profile) + the setOutDistance algebra, every active node
//   k_average             averageProfile (+ fused updateOutProfile); the self distance of the new node is added up here, or --
//                         deferred -- by an extra CTA of the next k_eval_wide / by k_self_sum;  k_average_batch: recomputeProfiles
//   k_peer_allgather, k_rank_merge, k_scatter_outdist (vft_dist.cuh)   one tree sharded over GPUs: the exchange step
//   k_ingest, k_ingest_gather (vft_ingest.cuh)   character decoding + row hashes + compaction of the distinct rows
//   k_outprofile_*        updateOutProfile / outProfile + setCodeDist
//   k_pair_loglk, k_posterior   pairLogLk / posteriorProfile (vft_ml.cuh): one tree level, or one lock-step round of the
//                         branch-length / NNI optimisers (ml_opt.cpp), per launch; 1-8 warps per (pair, length) item
//   k_sh_support          SHSupport: ordered gather-sums over resampled columns, one thread per (quartet, resample)
//   k_spec_commit         state commit of a speculative join (vft_spec_join_*: k_average into a shadow out-profile +
//                         k_eval in raw mode, launched ahead of the host's decision)
// Scratch profile rows (cfg.nScratch, ids 2N ...) extend codes / weights / vecs for the temporaries of the ML phase.
// There is no CPU fallback: without a usable device vft_ctx_create returns VFT_ENODEVICE.
#include "../../include/vft_b200.h"
#include "vft_device.cuh"
#include "vft_bulk.cuh"
#include "vft_ml.cuh"
#include "nj_loop.h"
#include "nj_loop_logic.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace vft;

static thread_local char g_err[512] = "";
static int fail(int code, const char *msg) { std::snprintf(g_err, sizeof g_err, "%s", msg); return code; }
static int cuda_fail(cudaError_t e, const char *what) {
    std::snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return VFT_ECUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

// ---- allocation cache -----------------------------------------------------------------------------
// vft_nj_build creates and destroys a context per tree; cudaMalloc / cudaHostAlloc / cudaFree cost
// milliseconds to hundreds of milliseconds each (cudaFree also synchronises the device).  Freed blocks
// are therefore parked in a small process-wide cache (exact-size reuse, per device) instead of going
// back to the driver: a caching allocator, nothing else -- contents are never reused, every context
// re-initialises its buffers.  VFT_POOL_MB caps the parked device bytes (default 8192; 0 disables).
#include <mutex>
#include <unordered_map>
namespace {
enum { MEM_DEVICE = 0, MEM_PINNED = 1 };
struct MemBlock { void *p; size_t bytes; int kind, device; };
std::mutex g_memMu;
std::unordered_map<void *, MemBlock> g_memLive;
std::vector<MemBlock> g_memParked;
size_t g_parkedBytes[2] = {0, 0};

size_t pool_cap(int kind) {
    static const long mb = [] { const char *e = std::getenv("VFT_POOL_MB"); return e ? std::atol(e) : 8192l; }();
    return kind == MEM_DEVICE ? (size_t) mb << 20 : std::min<size_t>((size_t) mb << 20, (size_t) 512 << 20);
}
void raw_free(const MemBlock &b) { if (b.kind == MEM_DEVICE) cudaFree(b.p); else cudaFreeHost(b.p); }

void mem_release_all() {
    std::lock_guard<std::mutex> lk(g_memMu);
    for (const MemBlock &b : g_memParked) raw_free(b);
    g_memParked.clear();
    g_parkedBytes[0] = g_parkedBytes[1] = 0;
}

cudaError_t mem_alloc(void **out, size_t bytes, int kind) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (bytes == 0) bytes = 1;
    {
        std::lock_guard<std::mutex> lk(g_memMu);
        for (size_t k = 0; k < g_memParked.size(); k++) {
            const MemBlock b = g_memParked[k];
            if (b.bytes == bytes && b.kind == kind && b.device == dev) {
                g_memParked[k] = g_memParked.back(); g_memParked.pop_back();
                g_parkedBytes[kind] -= bytes;
                g_memLive[b.p] = b;
                *out = b.p;
                return cudaSuccess;
            }
        }
    }
    cudaError_t e = kind == MEM_DEVICE ? cudaMalloc(out, bytes) : cudaHostAlloc(out, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) {                  // out of memory: give the parked blocks back and retry once
        cudaGetLastError();
        mem_release_all();
        e = kind == MEM_DEVICE ? cudaMalloc(out, bytes) : cudaHostAlloc(out, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_memMu);
        g_memLive[*out] = MemBlock{*out, bytes, kind, dev};
    }
    return e;
}

void mem_free(void *p) {
    if (!p) return;
    MemBlock b;
    {
        std::lock_guard<std::mutex> lk(g_memMu);
        auto it = g_memLive.find(p);
        if (it == g_memLive.end()) return;
        b = it->second;
        g_memLive.erase(it);
        if (g_parkedBytes[b.kind] + b.bytes <= pool_cap(b.kind) && g_memParked.size() < 256) {
            g_memParked.push_back(b);
            g_parkedBytes[b.kind] += b.bytes;
            return;
        }
    }
    raw_free(b);
}
}  // namespace

extern "C" int vft_release_cached_memory(void) { mem_release_all(); return VFT_OK; }

// =================================================================================================
// kernels
// =================================================================================================

// One launch evaluates a mixed request list (vft_eval_batch): items [0,nOutItems) are
// setOutDistance requests (a = node), the rest are pair-distance requests (a,b).  Each warp owns G
// consecutive items: leaf x leaf pairs are done by their own lane (seqDist is byte work), every other
// item is evaluated by the whole warp, one after the other (profile_dist_warp).  G=1 for the small
// per-join lists (latency), larger for the refresh batches (throughput).
// ia/ib/r0/r1 live in pinned host memory mapped into the device address space (zero-copy): a
// request costs one launch and one stream synchronisation, no separate memcpy.
constexpr int INLINE_ITEMS = 896;
constexpr int WIDE_THREADS[6] = {384, 192, 128, 96, 64, 32};      // CTA sizes of the CTA-per-pair kernel (k_eval_wide), largest first
constexpr size_t SPEC_MAX = 4096;                                     // items of one speculative join request
struct InlineItems { int32_t a[INLINE_ITEMS], b[INLINE_ITEMS]; };     // 7 KB of kernel parameters

// profileDist(new, new) from the per-position terms k_average left in global memory (deferred self distance): the ordered
// sums of NJ.tcc:1172-1183 by one lane, terms staged through shared memory by the warp
template<typename P>
__device__ __forceinline__ void self_sum_from_terms(const Store<P> &s, const double *__restrict__ gTerms, int64_t oid, double *sT) {
    for (int64_t k = threadIdx.x; k < 2 * s.Lp; k += blockDim.x) sT[k] = __ldcg(gTerms + k);
    __syncthreads();
    if (threadIdx.x == 0) {
        double top = 0, denom = 0;
        for (int64_t pos = 0; pos < s.Lp; pos += 8) {
            double w8[8], t8[8];
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const double2 a = *reinterpret_cast<const double2 *>(sT + pos + k), b = *reinterpret_cast<const double2 *>(sT + s.Lp + pos + k);
                w8[k] = a.x; w8[k + 1] = a.y; t8[k] = b.x; t8[k + 1] = b.y;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) { denom = xadd(denom, w8[k]); top = xadd(top, t8[k]); }
        }
        s.selfweight[oid] = (P) (denom > 0 ? denom : 0.01);
        s.selfdist[oid] = (P) (denom > 0 ? top / denom : 1.0);
    }
}
template<typename P>
__global__ void __launch_bounds__(128)
k_self_sum(Store<P> s, const double *__restrict__ gTerms, int64_t oid) {
    extern __shared__ __align__(16) unsigned char smem[];
    self_sum_from_terms<P>(s, gTerms, oid, reinterpret_cast<double *>(smem));
}

// DENSE = the throughput build for batches (fp32: capped at 128 registers -> 16 warps/SM, a few bytes of spill);
// the per-join lists (one item per warp, latency bound) use the uncapped build.
template<typename P, int A, bool MATRIX, bool DENSE>
__global__ void __launch_bounds__(128, (DENSE && sizeof(P) == 4) ? 4 : 2)
k_eval(Store<P> s, const __grid_constant__ InlineItems inl, const int32_t *__restrict__ ia, const int32_t *__restrict__ ib,
       int64_t n, int64_t nOutItems, int G, int raw, int64_t nActive, double totdiam, P *__restrict__ r0, P *__restrict__ r1,
       unsigned int *__restrict__ doneCount, P *__restrict__ hostOut, volatile unsigned int *hostFlag = nullptr, unsigned int flagVal = 0) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    unsigned char *smw = smemRaw + (threadIdx.x >> 5) * group_smem_bytes<P, A, MATRIX>(G);
    const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
    const int64_t item = warp * G + lane;
    const bool inRange = lane < G && item < n;
    // small lists travel in the kernel parameters (no PCIe read at kernel start), large ones are
    // read from the mapped pinned buffer (or, for vft_tophits_merge, from device memory)
    const int64_t a = !inRange ? -1 : (ia ? ia[item] : inl.a[item]), b = !inRange ? -1 : (ib ? ib[item] : inl.b[item]);
    const bool valid = inRange && a >= 0;                // a < 0: empty slot (vft_tophits_merge), nothing is written
    const bool isOut = valid && item < nOutItems;
    const bool isSeq = valid && !isOut && !raw && a < s.nSeqs && b < s.nSeqs;
    P d = 0, w = 0;
    if (isSeq) {
        seq_dist<P, MATRIX>(s, s.codes + a * s.Lp, s.codes + b * s.Lp, d, w);
        d = (P) xadd((double) d, 0.0);                                                    // NJ.tcc:1122
    }
    const bool isProf = valid && !isSeq;
    const unsigned mask = __ballot_sync(full, isProf);
    double den, top;
    group_profile_dist<P, A, MATRIX>(s, a, isOut ? (int64_t) -1 : b, mask, G, smw, den, top);
    if (isProf) {
        P dd, ww;
        finish_dist<P>(den, top, dd, ww);
        if (isOut && !raw) d = out_distance_finish<P>(s, a, nActive, totdiam, dd, ww);
        else { d = raw ? dd : join_correct<P>(s, a, b, dd); w = ww; }      // raw out item: bare profileDist(node, out-profile)
    }
    if (valid) { r0[item] = d; r1[item] = w; }
    // Results go to DEVICE memory; the last CTA to finish copies both arrays to the mapped host buffer with
    // wide, coalesced stores.  (One 4-byte store per item straight into host memory costs a PCIe
    // transaction each -- several microseconds for a 250-item list, more than the distances themselves.)
    if (hostOut == nullptr) return;
    __shared__ bool amLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) amLast = atomicAdd(doneCount, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) { hostOut[k] = __ldcg(r0 + k); hostOut[n + k] = __ldcg(r1 + k); }
    if (threadIdx.x == 0) *doneCount = 0;
    if (hostFlag != nullptr) {                           // the host spins on this word instead of synchronising the stream
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *hostFlag = flagVal;
    }
}

// The same request list, ONE CTA PER ITEM (the per-join lists of long alignments: a few hundred pairs of
// ~100 KB each cannot fill the machine with one warp per pair, and one warp walking 40+ chunks is latency
// bound).  The CTA's warps split the chunks of the pair, write the per-position terms into one full-length
// row in shared memory, and one lane adds the row in position order.  Leaf x leaf pairs take the same
// path: with unit weights profileDist's terms ARE seqDist's (NJ.tcc:1601-1624), only the no-overlap
// weight differs (0 instead of 0.01).
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(sizeof(P) == 4 ? 384 : 256)
k_eval_wide(Store<P> s, const __grid_constant__ InlineItems inl, const int32_t *__restrict__ ia, const int32_t *__restrict__ ib,
            int64_t n, int64_t nOutItems, int raw, int64_t nActive, double totdiam, P *__restrict__ r0, P *__restrict__ r1,
            unsigned int *__restrict__ doneCount, P *__restrict__ hostOut, volatile unsigned int *hostFlag = nullptr, unsigned int flagVal = 0,
            int64_t selfNode = -1, const double *__restrict__ gTerms = nullptr, unsigned int *selfFlag = nullptr, unsigned int selfSeq = 0) {
    // selfNode >= 0: the newest node's self distance is still pending (k_average left its terms in gTerms): CTA n -- one more than
    // the items -- adds them up while the item CTAs work; the item that needs the result (the node's out-distance) waits for the
    // flag at the very end of its own chain.  Every CTA of the launch is co-resident (the host sizes the CTAs for that), so the
    // wait cannot deadlock.
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int64_t item = blockIdx.x;
    const bool selfCta = item == n;
    const int64_t a = selfCta ? -1 : (ia ? ia[item] : inl.a[item]), b = selfCta ? -1 : (ib ? ib[item] : inl.b[item]);
    if (selfCta) {
        self_sum_from_terms<P>(s, gTerms, selfNode, reinterpret_cast<double *>(smemRaw));
        if (threadIdx.x == 0) { __threadfence(); atomicExch(selfFlag, selfSeq); }
    }
    if (a >= 0) {
        const bool isOut = item < nOutItems;
        const bool isSeq = !isOut && !raw && a < s.nSeqs && b < s.nSeqs;
        double den, top;
        group_profile_dist<P, A, MATRIX, true>(s, a, isOut ? (int64_t) -1 : b, 1u, 1, smemRaw, den, top, (int) (threadIdx.x >> 5),
                                               (int) (blockDim.x >> 5));
        __syncthreads();
        if (threadIdx.x < 32) {
            cta_ordered_sum<P, A, MATRIX>(s, smemRaw, den, top);
            if (threadIdx.x == 0) {
                P dd, ww, d, w = 0;
                finish_dist<P>(den, top, dd, ww);
                if (isOut && a == selfNode) {                // needs selfdist / selfweight of the node whose self distance CTA n computes
                    while (atomicAdd(selfFlag, 0u) != selfSeq) { }
                    __threadfence();
                }
                if (isOut && !raw) d = out_distance_finish<P>(s, a, nActive, totdiam, dd, ww);
                else if (isSeq) { d = (P) xadd((double) dd, 0.0); w = den > 0 ? ww : (P) 0; }     // :1621-1622, :1122
                else { d = raw ? dd : join_correct<P>(s, a, b, dd); w = ww; }
                r0[item] = d; r1[item] = w;
            }
        }
    }
    if (hostOut == nullptr) return;
    __shared__ bool amLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) amLast = atomicAdd(doneCount, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) { hostOut[k] = __ldcg(r0 + k); hostOut[n + k] = __ldcg(r1 + k); }
    if (threadIdx.x == 0) *doneCount = 0;
    if (hostFlag != nullptr) {                           // the host spins on this word instead of synchronising the stream
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *hostFlag = flagVal;
    }
}

// setBestHit (NJ.tcc:3571-3639) for a LEAF query: one thread per node slot (leaf x leaf = seqDist).
// Candidate addressing of the three all-candidate sweeps: list == nullptr -- slot k IS node k (inactive nodes and nodes
// outside [jBegin, jEnd) get a never-selected key); list != nullptr -- the COMPACT form: slot k is node
// list[k * stride + offset] of the ascending active list, every slot is live, results are indexed by slot.  stride/offset
// pick the strided share of one rank when a tree is sharded over GPUs (vft_dist_init): slot order == node order either way,
// so the (criterion, node descending) tie rule of the select is unchanged.
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128)
k_one_vs_all_leaf(Store<P> s, int64_t query, int64_t nActive, int64_t maxnode, int64_t jBegin, int64_t jEnd,
                  P *__restrict__ dist, P *__restrict__ weight, P *__restrict__ crit, uint64_t *__restrict__ keys,
                  const int32_t *__restrict__ list = nullptr, int stride = 1, int offset = 0) {
    const int64_t k = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (k >= maxnode) return;
    int64_t j = k;
    if (list != nullptr) j = list[k * stride + offset];
    else if (!s.active[j] || j < jBegin || j >= jEnd) { keys[k] = ~0ull; return; }
    P d, w;
    join_dist<P, A, MATRIX>(s, query, j, false, d, w);
    // setCriterion (NJ.tcc:1099-1107) with every out-distance fresh at this nActive
    const double outI = (double) s.outDist[query], outJ = (double) s.outDist[j];
    const P c = (P) xsub((double) d, xadd(outI, outJ) / (double) (nActive - 2));
    dist[k] = d; weight[k] = w; crit[k] = c;
    keys[k] = order_key(c);
}

// setBestHit for an INTERNAL query: every distance is a profileDist; each warp owns G node slots
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128, sizeof(P) == 4 ? 4 : 2)
k_one_vs_all_warp(Store<P> s, int64_t query, int64_t nActive, int64_t maxnode, int64_t jBegin, int64_t jEnd, int G,
                  P *__restrict__ dist, P *__restrict__ weight, P *__restrict__ crit, uint64_t *__restrict__ keys,
                  const int32_t *__restrict__ list = nullptr, int stride = 1, int offset = 0) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    unsigned char *smw = smemRaw + (threadIdx.x >> 5) * group_smem_bytes<P, A, MATRIX>(G);
    const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
    const int64_t k = warp * G + lane;
    const bool valid = lane < G && k < maxnode;
    int64_t j = k;
    bool act;
    if (list != nullptr) { j = valid ? (int64_t) list[k * stride + offset] : 0; act = valid; }
    else act = valid && s.active[j] && j >= jBegin && j < jEnd;
    const unsigned mask = __ballot_sync(full, act);
    double den, top;
    group_profile_dist<P, A, MATRIX>(s, query, j, mask, G, smw, den, top);
    if (!valid) return;
    if (!act) { keys[k] = ~0ull; return; }
    P d, w;
    finish_dist<P>(den, top, d, w);
    d = join_correct<P>(s, query, j, d);
    const double outI = (double) s.outDist[query], outJ = (double) s.outDist[j];
    const P c = (P) xsub((double) d, xadd(outI, outJ) / (double) (nActive - 2));
    dist[k] = d; weight[k] = w; crit[k] = c;
    keys[k] = order_key(c);
}

// The two all-candidate sweeps with the shared profile staged in shared memory by TMA bulk copies (vft_bulk.cuh): 512-thread
// CTAs (16 warps x up to 16 pairs each), one CTA per SM next to ~110 KB of staged rows.  Same arithmetic, same results.
constexpr int STG_T = 512, STG_G = 12;
template<typename P, int A, bool MATRIX>
__host__ __device__ inline size_t staged_tile_bytes() { return (size_t) (STG_T / 32) * group_smem_bytes<P, A, MATRIX>(STG_G); }

template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(STG_T, 1)
k_out_distance_all_staged(Store<P> s, int64_t maxnode, int G, int64_t nActive, double totdiam,
                          const int32_t *__restrict__ list = nullptr, int stride = 1, int offset = 0, P *__restrict__ res = nullptr) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ __align__(8) uint64_t bar;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    unsigned char *stage = smemRaw + staged_tile_bytes<P, A, MATRIX>();
    const uint32_t ocdBytes = (uint32_t) (s.Lp * A * sizeof(P)), owBytes = (uint32_t) (s.Lp * sizeof(P));
    const BulkSrc src[3] = {{s.ocd, ocdBytes}, {s.ow, owBytes}, {nullptr, 0}};
    cta_bulk_stage(stage, src, &bar);
    Store<P> s2 = s;
    s2.ocd = reinterpret_cast<P *>(stage); s2.ow = reinterpret_cast<P *>(stage + ocdBytes);
    unsigned char *smw = smemRaw + (threadIdx.x >> 5) * group_smem_bytes<P, A, MATRIX>(STG_G);
    const int64_t totalWarps = (int64_t) gridDim.x * (STG_T / 32);
    for (int64_t warp = (blockIdx.x * (int64_t) STG_T + threadIdx.x) >> 5; warp * G < maxnode; warp += totalWarps) {
        const int64_t k = warp * G + lane;
        int64_t j = k;
        bool act = lane < G && k < maxnode;
        if (list != nullptr) j = act ? (int64_t) list[k * stride + offset] : 0;
        else act = act && s.active[j];
        const unsigned mask = __ballot_sync(full, act);
        double den, top;
        group_profile_dist<P, A, MATRIX>(s2, j, (int64_t) -1, mask, G, smw, den, top);
        if (act) {
            P dd, ww;
            finish_dist<P>(den, top, dd, ww);
            const P v = out_distance_finish<P>(s, j, nActive, totdiam, dd, ww);
            if (res != nullptr) res[k] = v; else s.outDist[j] = v;
        }
        __syncwarp();
    }
}

template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(STG_T, 1)
k_one_vs_all_staged(Store<P> s, int64_t query, int64_t nActive, int64_t maxnode, int64_t jBegin, int64_t jEnd, int G,
                    P *__restrict__ dist, P *__restrict__ weight, P *__restrict__ crit, uint64_t *__restrict__ keys,
                    const int32_t *__restrict__ list = nullptr, int stride = 1, int offset = 0) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ __align__(8) uint64_t bar;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    unsigned char *stage = smemRaw + staged_tile_bytes<P, A, MATRIX>();
    // the query is an internal node: codes, weights, vectors (the vectors first: 16-byte alignment of every region)
    const int64_t row = query - s.nSeqs;
    const uint32_t vBytes = (uint32_t) (s.Lp * A * sizeof(P)), wBytes = (uint32_t) (s.Lp * sizeof(P)), cBytes = (uint32_t) s.Lp;
    const BulkSrc src[3] = {{s.vecs + row * s.Lp * A, vBytes}, {s.weights + row * s.Lp, wBytes}, {s.codes + query * s.Lp, cBytes}};
    cta_bulk_stage(stage, src, &bar);
    Store<P> s2 = s;
    s2.ovId = (int32_t) query; s2.ovV = reinterpret_cast<const P *>(stage); s2.ovW = reinterpret_cast<const P *>(stage + vBytes);
    s2.ovCodes = stage + vBytes + wBytes;
    unsigned char *smw = smemRaw + (threadIdx.x >> 5) * group_smem_bytes<P, A, MATRIX>(STG_G);
    const double outI = (double) s.outDist[query];
    const int64_t totalWarps = (int64_t) gridDim.x * (STG_T / 32);
    for (int64_t warp = (blockIdx.x * (int64_t) STG_T + threadIdx.x) >> 5; warp * G < maxnode; warp += totalWarps) {
        const int64_t k = warp * G + lane;
        const bool valid = lane < G && k < maxnode;
        int64_t j = k;
        bool act;
        if (list != nullptr) { j = valid ? (int64_t) list[k * stride + offset] : 0; act = valid; }
        else act = valid && s.active[j] && j >= jBegin && j < jEnd;
        const unsigned mask = __ballot_sync(full, act);
        double den, top;
        group_profile_dist<P, A, MATRIX>(s2, query, j, mask, G, smw, den, top);
        if (valid) {
            if (!act) keys[k] = ~0ull;
            else {
                P d, w;
                finish_dist<P>(den, top, d, w);
                d = join_correct<P>(s, query, j, d);
                const P c = (P) xsub((double) d, xadd(outI, (double) s.outDist[j]) / (double) (nActive - 2));
                dist[k] = d; weight[k] = w; crit[k] = c;
                keys[k] = order_key(c);
            }
        }
        __syncwarp();
    }
}

// setOutDistance for every active node (NJ.tcc:257-260 / :4451-4464), committed to s.outDist -- or, in the compact form of a
// sharded sweep, written to res[slot] (the rank's share; committed everywhere after the exchange)
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128, sizeof(P) == 4 ? 4 : 2)
k_out_distance_all(Store<P> s, int64_t maxnode, int G, int64_t nActive, double totdiam,
                   const int32_t *__restrict__ list = nullptr, int stride = 1, int offset = 0, P *__restrict__ res = nullptr) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    unsigned char *smw = smemRaw + (threadIdx.x >> 5) * group_smem_bytes<P, A, MATRIX>(G);
    const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
    const int64_t k = warp * G + lane;
    int64_t j = k;
    bool act = lane < G && k < maxnode;
    if (list != nullptr) j = act ? (int64_t) list[k * stride + offset] : 0;
    else act = act && s.active[j];
    const unsigned mask = __ballot_sync(full, act);
    double den, top;
    group_profile_dist<P, A, MATRIX>(s, j, (int64_t) -1, mask, G, smw, den, top);
    if (act) {
        P dd, ww;
        finish_dist<P>(den, top, dd, ww);
        const P v = out_distance_finish<P>(s, j, nActive, totdiam, dd, ww);
        if (res != nullptr) res[k] = v; else s.outDist[j] = v;
    }
}

// ---- top-K in psort order: key ascending, ties by index DESCENDING -----------------------------
struct KV { uint64_t key; uint32_t idx; };
__device__ __forceinline__ bool kv_less(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) {
    return ka < kb || (ka == kb && ia > ib);
}

template<typename P>
struct Rec { int64_t j; P dist, weight, crit; };

// ---- top-K in ONE kernel: radix select + small sort -------------------------------------------------
// The K best of n keys in psort order = the K smallest COMPOSITE keys (criterion bits, 0xFFFFFFFF - index):
// composites are unique, so ties need no special casing.  One CTA finds the K-th composite with an
// 8-bit-per-pass MSB radix select over the keys (L2-resident: n*8 bytes per pass), compacts the K
// selected elements into shared memory, sorts them with a bitonic network and writes the result
// records straight into mapped host memory.  KEYBYTES = 4 (float criterion) or 8 (double).
constexpr int SEL_T = 1024;
constexpr int SEL_MAXK = 4096;
constexpr int SEL_CTAS = 32;                 // chunks of the multi-CTA select

template<typename P, int KEYBYTES>
__global__ void __launch_bounds__(SEL_T)
k_topk_select(const uint64_t *__restrict__ keysAll, int64_t nAll, int K, const P *__restrict__ dist, const P *__restrict__ weight,
              const P *__restrict__ crit, Rec<P> *__restrict__ out, const uint32_t *__restrict__ idxIn = nullptr,
              uint64_t *__restrict__ candK = nullptr, uint32_t *__restrict__ candI = nullptr,
              const int32_t *__restrict__ list = nullptr, int lstride = 1, int loffset = 0) {
    // list != nullptr: the keys are indexed by SLOT of a compact candidate list (see k_one_vs_all_leaf); slot order == node
    // order, so the selection works on slots and only the record written at the end carries the node id
    // Multi-CTA use (long key arrays): stage 1 -- gridDim.x CTAs, each selects the K best of ITS chunk and writes them as
    // (key, original index) candidates (candK/candI, K per CTA, padded with never-selected keys); stage 2 -- one CTA runs the
    // same selection over the candidates (idxIn = their original indices).  The K best overall are among the per-chunk K
    // best, and the composite (key, index) order is the same at both stages.
    const int64_t chunk = (nAll + gridDim.x - 1) / gridDim.x;
    const int64_t chunkBase = (int64_t) blockIdx.x * chunk;
    const int64_t n = chunkBase >= nAll ? 0 : (nAll - chunkBase < chunk ? nAll - chunkBase : chunk);
    const uint64_t *__restrict__ keys = keysAll + chunkBase;
    if (idxIn != nullptr) idxIn += chunkBase;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned int *hist = reinterpret_cast<unsigned int *>(smem);                    // [32 warps][256]
    uint64_t *sk = reinterpret_cast<uint64_t *>(smem + 32 * 256 * 4);              // [SEL_MAXK] sort keys
    uint32_t *sv = reinterpret_cast<uint32_t *>(smem + 32 * 256 * 4 + SEL_MAXK * 8);   // [SEL_MAXK] node ids
    __shared__ uint64_t prefKey, maskKey;
    __shared__ uint32_t prefIdx, maskIdx;
    __shared__ unsigned int need, cnt, allTies;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int PASSES = KEYBYTES + 4;
    if (tid == 0) { prefKey = 0; maskKey = 0; prefIdx = 0; maskIdx = 0; need = (unsigned) K; cnt = 0; allTies = 0; }
    __syncthreads();
    for (int p = 0; p < PASSES; p++) {
        for (int t = tid; t < 32 * 256; t += SEL_T) hist[t] = 0;
        __syncthreads();
        const uint64_t pk = prefKey, mk = maskKey;
        const uint32_t pi = prefIdx, mi = maskIdx;
        const bool onKey = p < KEYBYTES;
        const int shift = onKey ? 8 * (KEYBYTES - 1 - p) : 8 * (3 - (p - KEYBYTES));
        for (int64_t base = 0; base < n; base += SEL_T) {
            const int64_t i = base + tid;
            bool in = false;
            unsigned int digit = 0;
            if (i < n) {
                const uint64_t k = keys[i];
                const uint32_t ni = 0xFFFFFFFFu - (idxIn ? idxIn[i] : (uint32_t) (chunkBase + i));
                in = ((k & mk) == pk) && ((ni & mi) == pi);
                digit = onKey ? (unsigned int) ((k >> shift) & 0xFFu) : ((ni >> shift) & 0xFFu);
            }
            // warp-aggregated increment of the warp-private histogram; fast path: the whole warp
            // agrees on the digit (the usual case: criteria share their leading bits)
            const unsigned act = __ballot_sync(0xFFFFFFFFu, in);
            if (act) {
                const int leader = __ffs(act) - 1;
                const unsigned int d0 = __shfl_sync(0xFFFFFFFFu, digit, leader);
                const unsigned same = __ballot_sync(0xFFFFFFFFu, in && digit == d0);
                if (same == act) { if (lane == leader) hist[wid * 256 + d0] += __popc(act); }
                else if (in) atomicAdd(&hist[wid * 256 + digit], 1u);
            }
        }
        __syncthreads();
        // fold the 32 private histograms, then thread 0 walks the 256 bins
        if (tid < 256) {
            unsigned int tot = 0;
            for (int w = 0; w < 32; w++) tot += hist[w * 256 + tid];
            hist[tid] = tot;
        }
        __syncthreads();
        if (wid == 0) {
            // warp 0 scans the 256 bins: 8 bins per lane, exclusive prefix over lanes, then the lane
            // whose range contains the `need`-th element walks its 8 bins
            unsigned int loc[8], sum = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) { loc[b] = hist[lane * 8 + b]; sum += loc[b]; }
            unsigned int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
            const unsigned int excl = incl - sum, want = need;
            const bool mine = excl < want && want <= incl;
            const unsigned who = __ballot_sync(0xFFFFFFFFu, mine);
            const int owner = who ? __ffs(who) - 1 : 31;
            if (lane == owner) {
                unsigned int cum = excl;
                int b = 0;
                for (; b < 7; b++) { if (cum + loc[b] >= want) break; cum += loc[b]; }
                const int d = lane * 8 + b;
                need = want - cum;
                if (onKey) { prefKey |= (uint64_t) d << shift; maskKey |= (uint64_t) 0xFFu << shift; }
                else { prefIdx |= (uint32_t) d << shift; maskIdx |= 0xFFu << shift; }
                // last key pass: the bin holds exactly the elements that tie with the K-th criterion.  If all of them
                // are needed (the usual case: no tie is cut) the four index passes have nothing left to decide.
                if (p == KEYBYTES - 1 && want - cum == loc[b]) { allTies = 1; prefIdx = 0xFFFFFFFFu; }
            }
        }
        __syncthreads();
        if (allTies) break;
    }
    // the K-th composite is (prefKey, prefIdx): select everything <= it
    const uint64_t tk = prefKey;
    const uint32_t ti = prefIdx;
    for (int64_t base = 0; base < n; base += SEL_T) {
        const int64_t i = base + tid;
        if (i < n) {
            const uint64_t k = keys[i];
            const uint32_t oi = idxIn ? idxIn[i] : (uint32_t) (chunkBase + i);
            const uint32_t ni = 0xFFFFFFFFu - oi;
            if (k < tk || (k == tk && ni <= ti)) {
                const unsigned int slot = atomicAdd(&cnt, 1u);
                if (slot < SEL_MAXK) { sk[slot] = k; sv[slot] = oi; }
            }
        }
    }
    __syncthreads();
    const int have = min((int) cnt, SEL_MAXK);
    int np2 = 1;
    while (np2 < have) np2 <<= 1;
    for (int t = tid; t < np2; t += SEL_T) if (t >= have) { sk[t] = ~0ull; sv[t] = 0u; }
    // bitonic sort of np2 elements: key ascending, index descending
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < np2 / 2; t += SEL_T) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint64_t ka = sk[lo], kb = sk[hi];
                const uint32_t ia = sv[lo], ib = sv[hi];
                const bool sw = up ? kv_less(kb, ib, ka, ia) : kv_less(ka, ia, kb, ib);
                if (sw) { sk[lo] = kb; sk[hi] = ka; sv[lo] = ib; sv[hi] = ia; }
            }
        }
    }
    __syncthreads();
    if (candK != nullptr) {                              // stage 1: this chunk's candidates
        for (int t = tid; t < K; t += SEL_T) {
            candK[(size_t) blockIdx.x * K + t] = t < have ? sk[t] : ~0ull;
            candI[(size_t) blockIdx.x * K + t] = t < have ? sv[t] : 0xFFFFFFFFu;
        }
        return;
    }
    for (int t = tid; t < K && t < have; t += SEL_T) {
        const uint32_t j = sv[t];
        if (j != 0xFFFFFFFFu) {                                                                                 // (padding of a short chunk: never a real node)
            out[t].j = list != nullptr ? (int64_t) list[(int64_t) j * lstride + loffset] : (int64_t) j;
            out[t].dist = dist[j]; out[t].weight = weight[j]; out[t].crit = crit[j];
        } else out[t].j = j;
    }
}

// the K best of n keys in psort order: one CTA for short arrays, chunked over CTAs + a merge stage for long ones
template<typename P, int KEYBYTES>
static void launch_topk(vft_ctx *c, const uint64_t *keys, int64_t n, int K, Rec<P> *out, const int32_t *list = nullptr, int lstride = 1, int loffset = 0);

// ---- top-hits refresh on the device: vft_tophits_merge -------------------------------------------------
// One CTA per list.  psort order (key ascending, ties in reverse input order) = ascending order of the
// unique composite (key, n-1-inputIndex): a bitonic network over (uint64 key, uint32 tie) pairs in shared memory.
constexpr int MRG_T = 256;
constexpr int MRG_MAX = 4096;

__device__ __forceinline__ void block_bitonic_sort(uint64_t *key, uint32_t *tie, int np2) {
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < np2 / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint64_t ka = key[lo], kb = key[hi];
                const uint32_t ta = tie[lo], tb = tie[hi];
                const bool aFirst = ka < kb || (ka == kb && ta < tb);
                if (up ? !aFirst : aFirst) { key[lo] = kb; key[hi] = ka; tie[lo] = tb; tie[hi] = ta; }
            }
        }
    }
    __syncthreads();
}

// transferBestHits(updateDistances=false) + the psort by (i,j) + dedupe of uniqueBestHits (NJ.tcc:4580-4613,
// :4786-4822): writes the surviving candidates of list l (ascending j) and, for those without a distance,
// a request slot for k_eval
template<typename P>
__global__ void __launch_bounds__(MRG_T)
k_merge_prep(const int32_t *__restrict__ iNode, const int32_t *__restrict__ ownOffset, const int32_t *__restrict__ ownJ,
             const P *__restrict__ ownDist, int nAvail, const int32_t *__restrict__ allJ, const P *__restrict__ allDist,
             int newnode, int cap, int np2, int nSeqs, int32_t *__restrict__ uJ, P *__restrict__ uD, int32_t *__restrict__ reqA,
             int32_t *__restrict__ reqB, int32_t *__restrict__ count, unsigned long long *__restrict__ acct) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *key = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tie = reinterpret_cast<uint32_t *>(smem + (size_t) np2 * 8);
    __shared__ int warpTot[MRG_T / 32];
    __shared__ int total;
    const int l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int iN = iNode[l], o0 = ownOffset[l], nOwn = ownOffset[l + 1] - o0, n = nOwn + nAvail;
    for (int e = tid; e < np2; e += MRG_T) {
        uint64_t k = ~0ull;
        uint32_t t = 0xFFFFFFFFu;
        if (e < n) {
            const int j = e < nOwn ? ownJ[o0 + e] : allJ[e - nOwn];
            k = (j >= 0 && j != iN) ? (uint64_t) (j + 1) : 0;                              // :1631-1637: no ancestor / self -> dropped
            t = (uint32_t) (n - 1 - e);
        }
        key[e] = k; tie[e] = t;
    }
    block_bitonic_sort(key, tie, np2);
    // first of each run of equal j survives (:4802-4821); block-wide compaction
    const int per = max(1, np2 / MRG_T), t0 = tid * per;
    int mine = 0;
    for (int t = t0; t < t0 + per && t < n; t++) mine += (key[t] != 0 && (t == 0 || key[t - 1] != key[t])) ? 1 : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warpTot[wid] = incl;
    __syncthreads();
    int base = incl - mine;
    for (int w = 0; w < wid; w++) base += warpTot[w];
    if (tid == MRG_T - 1) total = base + mine;
    unsigned long long nSeq = 0, nProf = 0, bytesLeaf = 0, bytesInt = 0;
    for (int t = t0; t < t0 + per && t < n; t++) {
        if (key[t] != 0 && (t == 0 || key[t - 1] != key[t])) {
            const int e = n - 1 - (int) tie[t], j = (int) key[t] - 1;
            const P d = e < nOwn ? ownDist[o0 + e] : (iN == newnode ? allDist[e - nOwn] : (P) -1e20);   // :4601-4606
            const size_t slot = (size_t) l * cap + base;
            const bool need = d < (P) 0;                                                      // :4826
            uJ[slot] = j; uD[slot] = d; reqA[slot] = need ? iN : -1; reqB[slot] = j;
            if (need) { if (iN < nSeqs && j < nSeqs) nSeq++; else { nProf++; if (j < nSeqs) bytesLeaf++; else bytesInt++; } }
            base++;
        }
    }
    __syncthreads();
    const int nu = total;
    for (int t = nu + tid; t < cap; t += MRG_T) reqA[(size_t) l * cap + t] = -1;
    if (tid == 0) count[l] = nu;
    // accounting: warp-reduced, one atomic per warp and counter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nSeq += __shfl_xor_sync(0xFFFFFFFFu, nSeq, o); nProf += __shfl_xor_sync(0xFFFFFFFFu, nProf, o);
        bytesLeaf += __shfl_xor_sync(0xFFFFFFFFu, bytesLeaf, o); bytesInt += __shfl_xor_sync(0xFFFFFFFFu, bytesInt, o);
    }
    if (lane == 0) {
        if (nSeq) atomicAdd(&acct[0], nSeq);
        if (nProf) atomicAdd(&acct[1], nProf);
        if (bytesLeaf + nSeq) atomicAdd(&acct[2], bytesLeaf + nSeq);
        if (bytesInt) atomicAdd(&acct[3], bytesInt);
    }
}

// the tail of uniqueBestHits (distances in, criterion from the fresh out-distance table, NJ.tcc:4823-4831,
// :1099-1107) + sortSaveBestHits (psort by criterion, first m; NJ.tcc:4535-4578)
template<typename P>
__global__ void __launch_bounds__(MRG_T)
k_merge_finish(Store<P> s, const int32_t *__restrict__ iNode, int64_t nActive, int m, int cap, int np2,
               const int32_t *__restrict__ uJ, P *__restrict__ uD, const int32_t *__restrict__ reqA, const P *__restrict__ r0,
               const int32_t *__restrict__ count, int32_t *__restrict__ outCount, int32_t *__restrict__ outJ, P *__restrict__ outDist) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *key = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tie = reinterpret_cast<uint32_t *>(smem + (size_t) np2 * 8);
    const int l = blockIdx.x, tid = threadIdx.x;
    const int iN = iNode[l], cnt = count[l];
    const double outI = (double) s.outDist[iN];
    for (int t = tid; t < np2; t += MRG_T) {
        uint64_t k = ~0ull;
        uint32_t ti = 0xFFFFFFFFu;
        if (t < cnt) {
            const size_t slot = (size_t) l * cap + t;
            const int j = uJ[slot];
            P d = uD[slot];
            if (reqA[slot] >= 0) { d = r0[slot]; uD[slot] = d; }
            const P c = (P) xsub((double) d, xadd(outI, (double) s.outDist[j]) / (double) (nActive - 2));
            k = order_key(c); ti = (uint32_t) (cnt - 1 - t);
        }
        key[t] = k; tie[t] = ti;
    }
    block_bitonic_sort(key, tie, np2);
    const int nSave = min(m, cnt);                       // every j is distinct and != iNode here
    for (int t = tid; t < nSave; t += MRG_T) {
        const size_t src = (size_t) l * cap + (cnt - 1 - (int) tie[t]);
        outJ[(size_t) l * m + t] = uJ[src]; outDist[(size_t) l * m + t] = uD[src];
    }
    if (tid == 0) outCount[l] = nSave;
}

// averageProfile (NJ.tcc:2067-2135) + profileDist(new,new) (NJ.tcc:3040-3043); ONE CTA:
// positions in parallel, then thread 0 adds the per-position self-distance terms in order.
template<typename P, int A, bool MATRIX, bool UPDATE>
__global__ void __launch_bounds__(256)
k_average(Store<P> s, int64_t oid, int64_t id1, int64_t id2, double bionjWeight, P diameterOut, int64_t nActiveOld,
          double *__restrict__ gTerms, unsigned int *__restrict__ doneCount, P *dOw, P *dOv, P *dOcd, P *specSelf,
          const njl::Scalars *jd = nullptr, int ppc = 0, int noSelf = 0) {
    // noSelf: the self distance is DEFERRED -- the terms go to gTerms and are added by the next per-join kernel (k_eval_wide's
    // extra CTA) or by k_self_sum, off this kernel's critical path
    // ppc: positions per CTA and pass (0: one per thread).  With ppc < blockDim.x the first ppc threads build the profile
    // (and the out-profile's new vector); then ALL threads share the 20 codeDist entries of each position (setCodeDist,
    // 20 three-way products of 20 per position: 85 % of this kernel's arithmetic, serial per position otherwise)
    // dOw/dOv/dOcd: where the UPDATEd out-profile goes (the live arrays, or the shadow copy of a speculative join);
    // specSelf != nullptr: speculative -- the self distance goes there and no per-node state is committed
    // jd != nullptr: the device-resident join loop (nj_loop_logic.h) -- the join is read from the loop's scalars
    extern __shared__ __align__(16) unsigned char smem[];
    bool doUpdate = UPDATE;
    if (jd != nullptr) {
        if (!jd->jdValid) return;
        oid = jd->jdNew; id1 = jd->jdI; id2 = jd->jdJ; diameterOut = (P) jd->jdDiameter; nActiveOld = jd->jdNActiveOld;
        doUpdate = UPDATE && jd->jdUpdate != 0;
    }
    const bool single = gridDim.x == 1 && !noSelf;               // short alignments: one CTA, the terms never leave shared memory
    double *termW = single ? reinterpret_cast<double *>(smem) : gTerms;   // [Lp] w*w   (global when the CTAs split the positions)
    double *termT = termW + s.Lp;                                // [Lp] w*w*piece
    const View<P, A> p1 = make_view<P, A>(s, id1), p2 = make_view<P, A>(s, id2);
    const int64_t row = oid - s.nSeqs;
    uint8_t *oc = s.codes + oid * s.Lp;
    P *ow = s.weights + row * s.Lp;
    P *ov = s.vecs + row * s.Lp * A;
    const int PPC = ppc > 0 ? ppc : (int) blockDim.x;
    for (int64_t base = blockIdx.x * (int64_t) PPC; base < s.Lp; base += (int64_t) gridDim.x * PPC) {
      const int64_t pos = base + threadIdx.x;
      if ((int) threadIdx.x < PPC && pos < s.Lp) {
        double tw = 0, tt = 0;
        if (pos < s.L) {
            const uint32_t c1 = p1.codes[pos], c2 = p2.codes[pos];
            const P w1 = p1.w ? p1.w[pos] : (c1 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
            const P w2 = p2.w ? p2.w[pos] : (c2 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
            const P wo = (P) xadd(xmul(bionjWeight, (double) w1), xmul(xsub(1.0, bionjWeight), (double) w2));   // :2075
            uint32_t co = VFT_DEV_NOCODE;
            if (wo > 0) {                                                                  // :2077-2085
                if (w1 > 0 && c1 != VFT_DEV_NOCODE && (w2 <= 0 || c1 == c2)) co = c1;
                else if (w1 <= 0 && w2 > 0 && c2 != VFT_DEV_NOCODE) co = c2;
            }
            P f[A];
#pragma unroll
            for (int k = 0; k < A; k++) f[k] = 0;
            const bool hasVec = wo > 0 && co == VFT_DEV_NOCODE;
            if (hasVec) {                                                                  // :2104-2112
                if (w1 > 0) add_to_freq<P, A, MATRIX>(s, f, xmul((double) w1, bionjWeight), c1,
                                                      (c1 == VFT_DEV_NOCODE && p1.v) ? p1.v + pos * A : nullptr);
                if (w2 > 0) add_to_freq<P, A, MATRIX>(s, f, xmul((double) w2, xsub(1.0, bionjWeight)), c2,
                                                      (c2 == VFT_DEV_NOCODE && p2.v) ? p2.v + pos * A : nullptr);
                normalize_freq<P, A, MATRIX>(s, f);
            }
            ow[pos] = wo;
            oc[pos] = (uint8_t) co;
#pragma unroll
            for (int k = 0; k < A; k++) ov[pos * A + k] = f[k];
            if (UPDATE && doUpdate) {
                // updateOutProfile for this position (NJ.tcc:943-1010), fused: the three profiles it
                // needs are already in registers
                P g[A];
#pragma unroll
                for (int k = 0; k < A; k++) g[k] = s.ov[pos * A + k];
                const double originalMult = (double) pmul(s.ow[pos], (P) nActiveOld);     // :962
                const double newMult = xsub(xsub(xadd(originalMult, (double) wo), (double) w1), (double) w2);   // :963
                P wout = (P) (newMult / (double) (nActiveOld - 1));                        // :964
                if (wout <= 0) wout = (P) 1e-20;
                dOw[pos] = wout;
#pragma unroll
                for (int k = 0; k < A; k++) g[k] = (P) xmul((double) g[k], originalMult);  // :969-971
                if (w1 > 0) add_to_freq<P, A, MATRIX>(s, g, (double) (-w1), c1, (c1 == VFT_DEV_NOCODE && p1.v) ? p1.v + pos * A : nullptr);
                if (w2 > 0) add_to_freq<P, A, MATRIX>(s, g, (double) (-w2), c2, (c2 == VFT_DEV_NOCODE && p2.v) ? p2.v + pos * A : nullptr);
                if (wo > 0) add_to_freq<P, A, MATRIX>(s, g, (double) wo, co, hasVec ? f : nullptr);
                normalize_freq<P, A, MATRIX>(s, g);                                        // :984
#pragma unroll
                for (int k = 0; k < A; k++) dOv[pos * A + k] = g[k];
                if (MATRIX && ppc == 0) code_dist_row<P, A, MATRIX>(s, g, dOcd + pos * A); // :1001-1003 (ppc > 0: by all threads, below)
            }
            if (wo > 0) {                                       // self-distance term, profileDist(out,out)
                const double wt = (double) pmul(wo, wo);
                tw = wt;
                tt = xmul(wt, piece<P, A, MATRIX>(s, co, co, f, f, nullptr));
            }
        } else {
            ow[pos] = 0; oc[pos] = (uint8_t) VFT_DEV_NOCODE;
#pragma unroll
            for (int k = 0; k < A; k++) ov[pos * A + k] = 0;
        }
        if (jd == nullptr) { termW[pos] = tw; termT[pos] = tt; }
      }
      if (MATRIX && UPDATE && ppc > 0) {
        // setCodeDist of the updated out-profile (NJ.tcc:873-898, :1001-1003): thread per (position, code)
        __syncthreads();
        if (doUpdate) {
            P e[A];
#pragma unroll
            for (int k = 0; k < A; k++) e[k] = s.eigenval[k];
            for (int task = threadIdx.x; task < PPC * A; task += blockDim.x) {
                const int64_t p2 = base + task / A;
                const int cc = task % A;
                if (p2 < s.L) {
                    P g[A], b[A];
                    load_vec<P, A>(dOv + p2 * A, g);
#pragma unroll
                    for (int k = 0; k < A; k++) b[k] = s.codeFreq[cc * 20 + k];
                    dOcd[p2 * A + cc] = (P) (double) vec_mul3_sum<P, A>(g, b, e, s.reduction);
                }
            }
        }
        __syncthreads();
      }
    }
    if (jd != nullptr) {
        // device-resident loop: profileDist(new, new) is evaluated with the join's request list (k_nj_eval), in parallel with
        // the other distances of the new node, instead of as a second chain behind this kernel
        if (blockIdx.x == 0 && threadIdx.x == 0) { s.diameter[oid] = diameterOut; s.active[id1] = 0; s.active[id2] = 0; s.active[oid] = 1; }
        return;
    }
    if (noSelf) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { s.diameter[oid] = diameterOut; s.active[id1] = 0; s.active[id2] = 0; s.active[oid] = 1; }
        return;
    }
    // the last CTA to finish adds the self-distance terms in position order
    double *sT = reinterpret_cast<double *>(smem);               // [2*Lp] both term arrays
    if (!single) {
        __shared__ bool amLast;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) amLast = atomicAdd(doneCount, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!amLast) return;
        __threadfence();
        for (int64_t k = threadIdx.x; k < 2 * s.Lp; k += blockDim.x) sT[k] = __ldcg(gTerms + k);
        if (threadIdx.x == 0) *doneCount = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // ordered sums over the positions (skipped positions hold +0.0, which is exact to add): 8 terms of each
        // chain are fetched ahead of the 16 dependent additions
        double top = 0, denom = 0;
        for (int64_t pos = 0; pos < s.Lp; pos += 8) {
            double w8[8], t8[8];
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const double2 a = *reinterpret_cast<const double2 *>(sT + pos + k), b = *reinterpret_cast<const double2 *>(sT + s.Lp + pos + k);
                w8[k] = a.x; w8[k + 1] = a.y; t8[k] = b.x; t8[k + 1] = b.y;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) { denom = xadd(denom, w8[k]); top = xadd(top, t8[k]); }
        }
        const P sw = (P) (denom > 0 ? denom : 0.01), sd = (P) (denom > 0 ? top / denom : 1.0);
        if (specSelf) { specSelf[0] = sd; specSelf[1] = sw; return; }
        s.selfweight[oid] = sw;
        s.selfdist[oid] = sd;
        s.diameter[oid] = diameterOut;
        s.active[id1] = 0; s.active[id2] = 0; s.active[oid] = 1;
    }
}

// averageProfile (NJ.tcc:2067-2135) for n independent (out, child, child) items: one tree level of the ME
// recomputeProfiles (NJ.tcc:3474-3506).  Only the profile rows are written -- no diameter, self distance or active flags.
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128)
k_average_batch(Store<P> s, const int32_t *__restrict__ oids, const int32_t *__restrict__ id1s, const int32_t *__restrict__ id2s, int64_t itemBase) {
    const int64_t item = itemBase + blockIdx.y;
    const int64_t oid = oids[item], id1 = id1s[item], id2 = id2s[item];
    const int64_t pos = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (pos >= s.Lp) return;
    const View<P, A> p1 = make_view<P, A>(s, id1), p2 = make_view<P, A>(s, id2);
    const int64_t row = oid - s.nSeqs;
    uint8_t *oc = s.codes + oid * s.Lp;
    P *ow = s.weights + row * s.Lp;
    P *ov = s.vecs + row * s.Lp * A;
    P f[A];
#pragma unroll
    for (int k = 0; k < A; k++) f[k] = 0;
    P wo = 0;
    uint32_t co = VFT_DEV_NOCODE;
    if (pos < s.L) {
        const uint32_t c1 = p1.codes[pos], c2 = p2.codes[pos];
        const P w1 = p1.w ? p1.w[pos] : (c1 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
        const P w2 = p2.w ? p2.w[pos] : (c2 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
        wo = (P) xadd(xmul(0.5, (double) w1), xmul(0.5, (double) w2));                   // :2075, unweighted (-1 => 0.5)
        if (wo > 0) {                                                                      // :2077-2085
            if (w1 > 0 && c1 != VFT_DEV_NOCODE && (w2 <= 0 || c1 == c2)) co = c1;
            else if (w1 <= 0 && w2 > 0 && c2 != VFT_DEV_NOCODE) co = c2;
        }
        if (wo > 0 && co == VFT_DEV_NOCODE) {                                              // :2104-2112
            if (w1 > 0) add_to_freq<P, A, MATRIX>(s, f, xmul((double) w1, 0.5), c1, (c1 == VFT_DEV_NOCODE && p1.v) ? p1.v + pos * A : nullptr);
            if (w2 > 0) add_to_freq<P, A, MATRIX>(s, f, xmul((double) w2, 0.5), c2, (c2 == VFT_DEV_NOCODE && p2.v) ? p2.v + pos * A : nullptr);
            normalize_freq<P, A, MATRIX>(s, f);
        }
    }
    ow[pos] = wo;
    oc[pos] = (uint8_t) co;
#pragma unroll
    for (int k = 0; k < A; k++) ov[pos * A + k] = f[k];
}

// a speculative join that turned out right: its per-node state becomes real (vft_spec_join_take)
template<typename P>
__global__ void k_spec_commit(Store<P> s, int64_t oid, int64_t id1, int64_t id2, P diameterOut, const P *__restrict__ specSelf) {
    s.selfdist[oid] = specSelf[0]; s.selfweight[oid] = specSelf[1];
    s.diameter[oid] = diameterOut;
    s.active[id1] = 0; s.active[id2] = 0; s.active[oid] = 1;
}

// updateOutProfile, NJ.tcc:943-1010: one thread per position
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128)
k_outprofile_update(Store<P> s, int64_t o1, int64_t o2, int64_t nw, int64_t nActiveOld) {
    const int64_t pos = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (pos >= s.L) return;
    const View<P, A> old1 = make_view<P, A>(s, o1), old2 = make_view<P, A>(s, o2), nn = make_view<P, A>(s, nw);
    const uint32_t c1 = old1.codes[pos], c2 = old2.codes[pos], cn = nn.codes[pos];
    const P w1 = old1.w ? old1.w[pos] : (c1 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
    const P w2 = old2.w ? old2.w[pos] : (c2 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
    const P wn = nn.w[pos];
    P f[A];
#pragma unroll
    for (int k = 0; k < A; k++) f[k] = s.ov[pos * A + k];
    const double originalMult = (double) pmul(s.ow[pos], (P) nActiveOld);                 // :962
    const double newMult = xsub(xsub(xadd(originalMult, (double) wn), (double) w1), (double) w2);   // :963
    P wout = (P) (newMult / (double) (nActiveOld - 1));                                    // :964
    if (wout <= 0) wout = (P) 1e-20;
    s.ow[pos] = wout;
#pragma unroll
    for (int k = 0; k < A; k++) f[k] = (P) xmul((double) f[k], originalMult);             // :969-971
    if (w1 > 0) add_to_freq<P, A, MATRIX>(s, f, (double) (-w1), c1, (c1 == VFT_DEV_NOCODE && old1.v) ? old1.v + pos * A : nullptr);
    if (w2 > 0) add_to_freq<P, A, MATRIX>(s, f, (double) (-w2), c2, (c2 == VFT_DEV_NOCODE && old2.v) ? old2.v + pos * A : nullptr);
    if (wn > 0) add_to_freq<P, A, MATRIX>(s, f, (double) wn, cn, (cn == VFT_DEV_NOCODE) ? nn.v + pos * A : nullptr);
    normalize_freq<P, A, MATRIX>(s, f);                                                    // :984
#pragma unroll
    for (int k = 0; k < A; k++) s.ov[pos * A + k] = f[k];
    if (MATRIX) code_dist_row<P, A, MATRIX>(s, f, s.ocd + pos * A);                        // :1001-1003
}

// outProfile, NJ.tcc:729-815.  The reference accumulates node after node (ascending id) into every
// position's weight and A frequencies; the sums are P-typed and therefore order dependent, so every
// accumulator stays a sequential chain over the nodes -- but the L*A (+L) chains are independent, and
// nothing forces the LOADS to be sequential.  Producer/consumer inside one CTA of 256 threads:
//   - all threads fetch the next tile (64 nodes x PP positions: codes, weights, vectors; coalesced,
//     thousands of loads in flight) into registers,
//   - meanwhile the first PP*A threads, one per (position, state), run their chains over the current
//     tile out of shared memory,
//   - the fetched tile is stored to the other shared-memory buffer; one barrier per tile.
// Time per tile ~ max(load latency, 64 dependent additions): close to the chain bound.
constexpr int REB_T = 256, REB_TN = 64;
template<int A> struct RebShape { static constexpr int PP = (A == 4) ? 32 : 8; };
template<typename P, int A>
__host__ __device__ inline size_t rebuild_smem_bytes() {
    constexpr int PP = RebShape<A>::PP;
    return (size_t) 2 * REB_TN * PP * A * sizeof(P) + (size_t) 2 * REB_TN * PP * sizeof(double) + 2 * REB_TN * 4
           + (size_t) PP * A * sizeof(P);
}

template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(REB_T)
k_outprofile_rebuild(Store<P> s, const int64_t *__restrict__ ids, int64_t n) {
    constexpr int PP = RebShape<A>::PP, TN = REB_TN, NPAIR = TN * PP / REB_T;
    extern __shared__ __align__(16) unsigned char smem[];
    double *sW = reinterpret_cast<double *>(smem);                         // [2][TN][PP]     weight * inweight, the addend of :741
    P *sV = reinterpret_cast<P *>(sW + 2 * TN * PP);                       // [2][TN][PP][A]  the addend of every (position, state) chain
    P *fs = sV + 2 * TN * PP * A;                                          // [PP][A]
    int *sId = reinterpret_cast<int *>(fs + PP * A);                       // [2][TN]
    const int tid = threadIdx.x;
    const int64_t pos0 = (int64_t) blockIdx.x * PP;
    const int nTiles = (int) ((n + TN - 1) / TN);
    const double inweight = 1.0 / (double) n;                              // :732
    // consumer role
    const bool chain = tid < PP * A;
    const int k = tid % A, pl = (tid / A) % PP;
    P wout = 0, f = 0;
    // producer registers: NPAIR (node, position) pairs per thread
    uint32_t rc[NPAIR]; P rw[NPAIR]; P rv[NPAIR][A];
    auto fetch = [&](int tile, int buf) {
#pragma unroll
        for (int q = 0; q < NPAIR; q++) {
            const int e = q * REB_T + tid, u = e / PP, p = e % PP;
            const int64_t iu = (int64_t) tile * TN + u, pos = pos0 + p;
            rc[q] = VFT_DEV_NOCODE; rw[q] = 0;
#pragma unroll
            for (int a = 0; a < A; a++) rv[q][a] = 0;
            if (iu < n) {
                const int64_t id = sId[buf * TN + u];
                rc[q] = (uint32_t) s.codes[id * s.Lp + pos];
                if (id < s.nSeqs) rw[q] = rc[q] != VFT_DEV_NOCODE ? (P) 1 : (P) 0;
                else {
                    const int64_t row = id - s.nSeqs;
                    rw[q] = s.weights[row * s.Lp + pos];
                    load_vec<P, A>(s.vecs + (row * s.Lp + pos) * A, rv[q]);
                }
            }
        }
    };
    // The producers also do everything of addToFreq (:821-833) that does not depend on the running sums: what reaches
    // shared memory is, per (node, position), the double addend of the weight chain and the A addends of the frequency
    // chains -- v*w, codeFreq[code]*w, or w at the node's own code; +0 where the reference adds nothing (x + 0 is exact).
    // The consumer chains are then one dependent addition per node, no branch, no lookup.
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < NPAIR; q++) {
            const int e = q * REB_T + tid;                                 // = u * PP + p
            const P w = rw[q];
            const uint32_t c = rc[q];
            sW[buf * TN * PP + e] = xmul((double) w, inweight);
            P add[A];
#pragma unroll
            for (int a = 0; a < A; a++) {
                add[a] = 0;
                if (w > 0) {
                    if (c == VFT_DEV_NOCODE) add[a] = pmul(rv[q][a], w);
                    else if (MATRIX) add[a] = pmul(s.codeFreq[c * 20 + a], w);
                    else if (c == (uint32_t) a) add[a] = w;
                }
            }
            store_vec<P, A>(sV + ((size_t) buf * TN * PP + e) * A, add);      // 128-bit stores: scalar ones conflict 4 ways on the A-strided rows
        }
    };
    if (tid < TN) sId[tid] = tid < n ? (int) ids[tid] : 0;
    if (tid < TN) sId[TN + tid] = (int64_t) TN + tid < n ? (int) ids[TN + tid] : 0;
    __syncthreads();
    fetch(0, 0);
    stash(0);
    __syncthreads();
    for (int t = 0; t < nTiles; t++) {
        const int cur = t & 1;
        if (t + 1 < nTiles) fetch(t + 1, cur ^ 1);                         // in flight during the chains below
        int idNext = 0;
        const bool haveId = t + 2 < nTiles && tid < TN;
        if (haveId) { const int64_t iu = (int64_t) (t + 2) * TN + tid; idNext = iu < n ? (int) ids[iu] : 0; }
        if (chain) {
            const double *cW = sW + cur * TN * PP + pl;
            const P *cV = sV + ((size_t) cur * TN * PP + pl) * A + k;
#pragma unroll 16
            for (int u = 0; u < TN; u++) {
                wout = (P) xadd((double) wout, cW[u * PP]);                                // :741
                // addToFreq, :821-833.  At a known code without a matrix the reference adds in double and narrows (:831);
                // with a 53-bit intermediate that is the same value as the P addition (2p+2 <= 53 for p = 24; P = double trivially)
                f = padd(f, cV[(size_t) u * PP * A]);
            }
        }
        if (t + 1 < nTiles) stash(cur ^ 1);
        if (haveId) sId[cur * TN + tid] = idNext;
        __syncthreads();
    }
    if (chain) fs[pl * A + k] = f;
    __syncthreads();
    const int64_t pos = pos0 + pl;
    if (chain && k == 0 && pos < s.L) {
        if (wout <= 0) wout = (P) 1e-20;                                                   // :743-745
        P fa[A];
#pragma unroll
        for (int q = 0; q < A; q++) fa[q] = fs[pl * A + q];
        normalize_freq<P, A, MATRIX>(s, fa);                                               // :789-794
        s.ow[pos] = wout;
#pragma unroll
        for (int q = 0; q < A; q++) s.ov[pos * A + q] = fa[q];
        if (MATRIX) code_dist_row<P, A, MATRIX>(s, fa, s.ocd + pos * A);                   // :801-803
    }
}

// pairLogLk (NJ.tcc:1192-1447): W warps per (pair, length) item, 8 / W items per CTA.  W = 1 for whole-tree batches
// (treeLogLk: one item per internal node); the lock-step Brent rounds of ml_opt.cpp carry tens to hundreds of items, far
// fewer than the machine has warps, and get up to 8 warps each.
template<typename P, int A>
__global__ void __launch_bounds__(256)
k_pair_loglk(Store<P> s, MLModel<P> m, const int32_t *__restrict__ ia, const int32_t *__restrict__ ib,
             const double *__restrict__ len, int64_t n, double *__restrict__ loglk, double *__restrict__ siteLk, int W, int tableBytes) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int nT = 32 * W, perCta = 8 / W, local = threadIdx.x / nT, tid = threadIdx.x - local * nT;
    const int64_t item = blockIdx.x * (int64_t) perCta + local;
    const size_t perItem = (size_t) s.Lp * 8 + (size_t) tableBytes;
    unsigned char *mine = smemRaw + local * perItem;
    const bool valid = item < n;
    pair_loglk_group<P, A>(s, m, valid, valid ? ia[item] : 0, valid ? ib[item] : 0, valid ? len[item] : 0.0,
                           reinterpret_cast<double *>(mine), mine + (size_t) s.Lp * 8, valid && siteLk ? siteLk + item * s.L : nullptr,
                           tid, nT, valid ? loglk + item : nullptr);
}

// posteriorProfile (NJ.tcc:2137-2447): one thread per position; the two expEigenRates tables (or the
// JC pSame/pDiff vectors) are built once per CTA in shared memory
template<typename P, int A>
__global__ void __launch_bounds__(128)
k_posterior(Store<P> s, MLModel<P> m, int64_t oid, int64_t id1, int64_t id2, double len1, double len2,
            const int32_t *__restrict__ items, const double *__restrict__ lens) {
    __shared__ __align__(16) unsigned char tab[2 * 64 * 20 * 8];
    if (items != nullptr) {          // batched (one tree level): item blockIdx.y = (out, id1, id2), (len1, len2)
        oid = items[3 * blockIdx.y]; id1 = items[3 * blockIdx.y + 1]; id2 = items[3 * blockIdx.y + 2];
        len1 = lens[2 * blockIdx.y]; len2 = lens[2 * blockIdx.y + 1];
    }
    if (len1 < m.MLMinBranchLength) len1 = m.MLMinBranchLength;                            // :2139-2144
    if (len2 < m.MLMinBranchLength) len2 = m.MLMinBranchLength;
    P *ee1 = reinterpret_cast<P *>(tab), *ee2 = reinterpret_cast<P *>(tab + 64 * 20 * 8);
    double *PS1 = reinterpret_cast<double *>(tab), *PD1 = PS1 + 64, *PS2 = reinterpret_cast<double *>(tab + 64 * 20 * 8), *PD2 = PS2 + 64;
    if (m.codeFreq) {
        exp_eigen_rates<P, A>(m, len1, ee1, threadIdx.x, blockDim.x);
        exp_eigen_rates<P, A>(m, len2, ee2, threadIdx.x, blockDim.x);
    } else {
        jc_tables<P>(m, len1, PS1, PD1, threadIdx.x, blockDim.x);
        jc_tables<P>(m, len2, PS2, PD2, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    const int64_t pos = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (pos >= s.Lp) return;
    const int64_t row = oid - s.nSeqs;
    P wOut = 0;
    uint32_t cOut = VFT_DEV_NOCODE;
    P f[A];
#pragma unroll
    for (int k = 0; k < A; k++) f[k] = 0;
    if (pos < s.L) {
        const View<P, A> p1 = make_view<P, A>(s, id1), p2 = make_view<P, A>(s, id2);
        posterior_site<P, A>(s, m, p1, p2, pos, ee1, ee2, PS1, PD1, PS2, PD2, wOut, cOut, f);
    }
    s.weights[row * s.Lp + pos] = wOut;
    s.codes[oid * s.Lp + pos] = (uint8_t) cOut;
#pragma unroll
    for (int k = 0; k < A; k++) s.vecs[(row * s.Lp + pos) * A + k] = f[k];
    if (pos == 0 && oid < 2 * s.nSeqs) s.active[oid] = 1;              // scratch rows have no per-node state
}

// SHSupport, NJ.tcc:1126-1165: one thread per (quartet, resample); the quartet's 3 x nPos site log-likelihoods sit in shared
// memory, the resampled columns are read transposed ([column][resample]: coalesced).  The three sums of a resample are ordered
// chains over the columns, as in the reference; the vote is an integer.
__global__ void __launch_bounds__(256)
k_sh_support(const double *__restrict__ siteLogLk, const double *__restrict__ loglk, const int32_t *__restrict__ colT, int64_t nPos,
             int64_t nBoot, unsigned int *__restrict__ votes) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *sl = reinterpret_cast<double *>(smem);                     // [3][nPos]
    const int64_t q = blockIdx.y;
    for (int64_t t = threadIdx.x; t < 3 * nPos; t += blockDim.x) sl[t] = siteLogLk[q * 3 * nPos + t];
    __syncthreads();
    const int64_t iBoot = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    bool vote = false;
    if (iBoot < nBoot) {
        const double l0 = loglk[3 * q], l1 = loglk[3 * q + 1], l2 = loglk[3 * q + 2];
        const double d1 = xsub(l0, l1), d2 = xsub(l0, l2), delta = d1 < d2 ? d1 : d2;
        double r0 = -l0, r1 = -l1, r2 = -l2;
        for (int64_t j = 0; j < nPos; j++) {
            const int pos = colT[j * nBoot + iBoot];
            r0 = xadd(r0, sl[pos]); r1 = xadd(r1, sl[nPos + pos]); r2 = xadd(r2, sl[2 * nPos + pos]);
        }
        const double r[3] = {r0, r1, r2};
        int best = 0;
        if (r[1] > r[best]) best = 1;
        if (r[2] > r[best]) best = 2;
        const double a = xsub(r[best], r[(best + 1) % 3]), b = xsub(r[best], r[(best + 2) % 3]);
        vote = (a < b ? a : b) < delta;
    }
    const unsigned int n = __syncthreads_count(vote);
    if (threadIdx.x == 0 && n) atomicAdd(&votes[q], n);
}

// leaves: selfweight = nPos - nGaps (NJ.tcc:249-252), active, padding of the code rows
template<typename P>
__global__ void k_init_leaves(Store<P> s) {
    const int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (i >= s.nSeqs) return;
    int64_t gaps = 0;
    for (int64_t p = 0; p < s.L; p++) gaps += s.codes[i * s.Lp + p] == VFT_DEV_NOCODE;
    s.selfweight[i] = (P) (s.L - gaps);
    s.selfdist[i] = 0; s.diameter[i] = 0; s.active[i] = 1;
}

// =================================================================================================
// host side of the ABI
// =================================================================================================

#include "vft_dist.cuh"
#include "vft_sweep.cuh"

struct vft_ctx {
    vft_config cfg;
    int A;
    int64_t N, M, L, Lp, maxnode;
    int64_t S = 0;                 // scratch profile rows (cfg.nScratch): ids M .. M+S-1, likelihood entry points only
    size_t ps;
    cudaStream_t stream = nullptr;
    // device
    void *codes, *weights, *vecs, *ow, *ov, *ocd, *diameter, *selfdist, *selfweight, *outDist, *active, *tables;
    void *d_dist, *d_weight, *d_crit;          // [M] one-vs-all scratch
    uint64_t *d_keys;                          // [M]
    uint64_t *d_candK = nullptr; uint32_t *d_candI = nullptr;   // [SEL_CTAS * SEL_MAXK] candidates of the chunked top-K select
    int64_t *d_ids, *d_pi, *d_pj;              // staging for lists
    void *d_out1, *d_out2;
    int64_t listCap;
    // vft_tophits_merge scratch
    double *d_terms;                           // [2*Lp] self-distance terms of k_average
    void *d_mrg; size_t mrgCap;
    bool wideOk;
    int wideSlots[6] = {0, 0, 0, 0, 0, 0};        // co-resident CTAs of k_eval_wide on the whole GPU at WIDE_THREADS[k] threads per CTA
    bool stagedOk = false;                     // the TMA-staged sweep kernels apply (fp32, 20 states, matrix mode, rows fit shared memory)
    size_t stagedSmem = 0;
    unsigned long long *d_acct;
    // pinned host
    void *h_in, *h_out;
    size_t hCap;
    std::vector<uint8_t> activeHost;
    int64_t nActLeaf, nActInternal;
    // the ascending list of active nodes, kept on the device for the compact all-candidate sweeps (rebuilt lazily)
    int32_t *d_act = nullptr, *h_act = nullptr; int64_t nAct = 0; bool actDirty = true;
    // query tables of the 20-state matrix sweeps (vft_sweep.cuh): [qtabCap][Lp][20] x2 + [qtabCap][Lp]
    void *d_qcd = nullptr, *d_qv = nullptr, *d_qw = nullptr; int32_t *d_qnodes = nullptr; int64_t qtabCap = 0;
    bool sweepOk = false;
    // the self distance of the newest node, deferred: its per-position terms wait in d_terms and are added up (in order) by an
    // extra CTA of the NEXT per-join distance kernel instead of at the tail of k_average (resolved by any other entry point)
    int64_t pendingSelf = -1; bool deferSelf = true; unsigned int *d_selfFlag = nullptr; unsigned int selfSeq = 0;
    int sweepModes = 7;                        // which sweeps use vft_sweep.cuh: bit 0 one-vs-all, bit 1 all-node out-distances, bit 2 refresh list merges (VFT_SWEEP_MODES)
    bool sharded = false;                      // member of the process's dist group (vft_dist.cuh): sweeps cover this rank's share
    // ML model
    void *mlTables, *mlRates;
    int32_t *mlRatecat;
    bool hasTransmat, hasRates;
    int nRateCats, fastexp;
    double MLMinRel, MLMinBr;
    unsigned int *d_doneCount;
    // speculative join (vft_spec_join_*): shadow out-profile, its own result buffers, completion event
    void *ow2 = nullptr, *ov2 = nullptr, *ocd2 = nullptr, *d_specR0 = nullptr, *d_specR1 = nullptr, *d_specSelf = nullptr;
    void *h_specIn = nullptr, *h_specOut = nullptr;
    cudaEvent_t specDone = nullptr;
    bool specPending = false;
    volatile unsigned int *h_flag = nullptr; unsigned int flagSeq = 0;     // completion word of the per-join request lists (pinned)
    int64_t specOut = -1, specId1 = -1, specId2 = -1, specNPairs = 0, specNOut = 0, specBytes = 0;
    vft_counters cnt;
    // stopwatch + optional per-kernel-class event timing (cfg.reserved & VFT_CFG_PROFILE)
    cudaEvent_t tmr0 = nullptr, tmr1 = nullptr;
    bool profile;
    struct Pending { cudaEvent_t a, b; int cls, kid; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
};

enum { CLS_DIST = 0, CLS_SELECT = 1, CLS_PROFILE = 2 };
// finer split of the same timings: index into vft_counters.msKernel / nKernel (names: VFT_KERNEL_NAMES in the header)
enum { K_EVAL_SMALL = 0, K_EVAL_LARGE, K_ONE_VS_ALL, K_OUT_DIST_ALL, K_SELECT, K_MERGE, K_AVERAGE, K_OUTPROFILE_UPDATE, K_REBUILD, K_LOGLK, K_POSTERIOR, K_NJ_STEP };

// attributes the algorithmic bytes accounted inside a scope to one kernel (vft_counters.bytesKernel)
struct BytesScope {
    vft_ctx *c; int kid; int64_t b0;
    BytesScope(vft_ctx *c, int kid);
    ~BytesScope();
};
static void prof_begin(vft_ctx *c, int cls, int kid) {
    if (!c->profile) return;
    vft_ctx::Pending p;
    for (cudaEvent_t *e : {&p.a, &p.b}) {
        if (c->pool.empty()) cudaEventCreate(e);
        else { *e = c->pool.back(); c->pool.pop_back(); }
    }
    p.cls = cls; p.kid = kid;
    cudaEventRecord(p.a, c->stream);
    c->pending.push_back(p);
}
static void prof_end(vft_ctx *c) {
    if (!c->profile) return;
    cudaEventRecord(c->pending.back().b, c->stream);
}
// call only after the stream has been synchronised
static void prof_resolve(vft_ctx *c) {
    if (!c->profile) return;
    for (auto &p : c->pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        if (p.cls == CLS_DIST) { c->cnt.msDist += ms; c->cnt.distLaunches++; }
        else if (p.cls == CLS_SELECT) c->cnt.msSelect += ms;
        else c->cnt.msProfile += ms;
        c->cnt.msKernel[p.kid] += ms; c->cnt.nKernel[p.kid]++;
        c->pool.push_back(p.a); c->pool.push_back(p.b);
    }
    c->pending.clear();
}
static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    asm volatile("pause" ::: "memory");
#else
    asm volatile("" ::: "memory");
#endif
}
// the calling thread may be a host-pool thread that never selected the context's device
static void resolve_pending_self(vft_ctx *c);
static inline void bind_device(vft_ctx *c, bool keepPendingSelf = false) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != c->cfg.device) cudaSetDevice(c->cfg.device);
    if (!keepPendingSelf && c->pendingSelf >= 0) resolve_pending_self(c);
}
static cudaError_t sync_stream(vft_ctx *c) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) prof_resolve(c);
    return e;
}

BytesScope::BytesScope(vft_ctx *c, int kid) : c(c), kid(kid), b0(c->cnt.algoBytes) {}
BytesScope::~BytesScope() { c->cnt.bytesKernel[kid] += c->cnt.algoBytes - b0; }

template<typename P>
static Store<P> make_store(vft_ctx *c) {
    Store<P> s;
    s.codes = (uint8_t *) c->codes; s.weights = (P *) c->weights; s.vecs = (P *) c->vecs;
    s.ow = (P *) c->ow; s.ov = (P *) c->ov; s.ocd = c->cfg.useMatrix ? (P *) c->ocd : nullptr;
    s.diameter = (P *) c->diameter; s.selfdist = (P *) c->selfdist; s.selfweight = (P *) c->selfweight;
    s.outDist = (P *) c->outDist; s.active = (uint8_t *) c->active;
    const P *t = (const P *) c->tables;
    s.distances = t; s.eigenval = t + 400; s.eigentot = t + 420; s.codeFreq = t + 440;
    s.nSeqs = c->N; s.L = c->L; s.Lp = c->Lp; s.reduction = c->cfg.reduction;
    s.fPostTotalTolerance = c->cfg.fPostTotalTolerance;
    s.ovId = -2; s.ovCodes = nullptr; s.ovW = nullptr; s.ovV = nullptr;
    return s;
}

// a deferred self distance that no per-join kernel picked up: one small kernel, ordered on the context's stream
static void resolve_pending_self(vft_ctx *c) {
    const int64_t oid = c->pendingSelf;
    c->pendingSelf = -1;
    const size_t smem = (size_t) c->Lp * 16;
    if (c->ps == 4) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_self_sum<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k_self_sum<float><<<1, 128, smem, c->stream>>>(make_store<float>(c), c->d_terms, oid);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_self_sum<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k_self_sum<double><<<1, 128, smem, c->stream>>>(make_store<double>(c), c->d_terms, oid);
    }
    c->cnt.launches++;
}

// dispatch on (precision, nCodes, useMatrix)
#define VFT_DISPATCH(c, CALL)                                                                  \
    do {                                                                                       \
        if ((c)->cfg.precision == 32) {                                                        \
            if ((c)->A == 4) { if ((c)->cfg.useMatrix) { CALL(float, 4, true); } else { CALL(float, 4, false); } } \
            else { if ((c)->cfg.useMatrix) { CALL(float, 20, true); } else { CALL(float, 20, false); } }           \
        } else {                                                                               \
            if ((c)->A == 4) { if ((c)->cfg.useMatrix) { CALL(double, 4, true); } else { CALL(double, 4, false); } } \
            else { if ((c)->cfg.useMatrix) { CALL(double, 20, true); } else { CALL(double, 20, false); } }         \
        }                                                                                      \
    } while (0)

extern "C" const char *vft_last_error(void) { return g_err; }
extern "C" void vftx_set_error(const char *msg) { std::snprintf(g_err, sizeof g_err, "%s", msg); }      // for the host drivers above the kernel-level ABI
extern "C" const char *vft_backend_name(void) { return "cuda-sm100a"; }

static int ensure_lists(vft_ctx *c, int64_t n) {
    if (n <= c->listCap) return VFT_OK;
    int64_t cap = std::max<int64_t>(n, 2 * c->listCap);
    mem_free(c->d_ids); mem_free(c->d_pi); mem_free(c->d_pj); mem_free(c->d_out1); mem_free(c->d_out2);
    c->d_ids = c->d_pi = c->d_pj = nullptr; c->d_out1 = c->d_out2 = nullptr; c->listCap = 0;      // nothing dangling if an allocation below fails
    CK(mem_alloc((void **) &c->d_ids, cap * 8, MEM_DEVICE)); CK(mem_alloc((void **) &c->d_pi, cap * 8, MEM_DEVICE)); CK(mem_alloc((void **) &c->d_pj, cap * 8, MEM_DEVICE));
    CK(mem_alloc((void **) &c->d_out1, cap * 8, MEM_DEVICE)); CK(mem_alloc((void **) &c->d_out2, cap * 8, MEM_DEVICE));
    c->listCap = cap;
    return VFT_OK;
}

static int ensure_pinned(vft_ctx *c, size_t bytes) {
    if (bytes <= c->hCap) return VFT_OK;
    size_t cap = std::max(bytes, 2 * c->hCap);
    mem_free(c->h_in); mem_free(c->h_out);
    c->h_in = c->h_out = nullptr; c->hCap = 0;
    CK(mem_alloc(&c->h_in, cap, MEM_PINNED)); CK(mem_alloc(&c->h_out, cap, MEM_PINNED));
    c->hCap = cap;
    return VFT_OK;
}

static int ctx_create_impl(const vft_config *cfg, vft_ctx **out, vft_ctx **partial) {
    if (!cfg || !out) return fail(VFT_EINVAL, "null argument");
    if (cfg->nSeqs < 1 || cfg->nPos < 1) return fail(VFT_EINVAL, "nSeqs and nPos must be positive");
    if (cfg->nCodes != 4 && cfg->nCodes != 20) return fail(VFT_EINVAL, "nCodes must be 4 or 20");
    if (cfg->precision != 32 && cfg->precision != 64) return fail(VFT_EINVAL, "precision must be 32 or 64");
    if (2 * cfg->nSeqs >= 0xFFFF0000ll) return fail(VFT_EINVAL, "too many sequences for 32-bit sort indices");
    if (cfg->nScratch < 0 || 2 * cfg->nSeqs + cfg->nScratch >= 0x7FFF0000ll) return fail(VFT_EINVAL, "bad nScratch");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        std::snprintf(g_err, sizeof g_err, "no CUDA device (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
        return VFT_ENODEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(VFT_EINVAL, "bad device ordinal");
    CK(cudaSetDevice(cfg->device));
    vft_ctx *c = new vft_ctx();      // value-initialised: every pointer starts out null
    *partial = c;                    // released by the caller if anything below fails
    c->cfg = *cfg; c->A = cfg->nCodes; c->N = cfg->nSeqs; c->M = 2 * cfg->nSeqs; c->L = cfg->nPos;
    c->Lp = (cfg->nPos + 31) / 32 * 32; c->ps = cfg->precision / 8; c->maxnode = 0;
    std::memset(&c->cnt, 0, sizeof c->cnt);
    c->profile = (cfg->reserved & VFT_CFG_PROFILE) != 0;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->tmr0)); CK(cudaEventCreate(&c->tmr1));
    c->S = cfg->nScratch;
    const size_t ps = c->ps, Lp = (size_t) c->Lp, A = (size_t) c->A, N = (size_t) c->N, M = (size_t) c->M, S = (size_t) c->S;
    // the scratch rows (ids M .. M+S-1) extend the three profile arrays; the per-node NJ arrays stay [M]
    CK(mem_alloc((void **) &c->codes, (M + S) * Lp, MEM_DEVICE));
    CK(mem_alloc((void **) &c->weights, (N + S) * Lp * ps, MEM_DEVICE));
    CK(mem_alloc((void **) &c->vecs, (N + S) * Lp * A * ps, MEM_DEVICE));
    CK(mem_alloc((void **) &c->ow, Lp * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->ov, Lp * A * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->ocd, Lp * A * ps, MEM_DEVICE));
    CK(mem_alloc((void **) &c->diameter, M * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->selfdist, M * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->selfweight, M * ps, MEM_DEVICE));
    CK(mem_alloc((void **) &c->outDist, M * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->active, M, MEM_DEVICE));
    CK(mem_alloc((void **) &c->tables, 840 * ps, MEM_DEVICE));
    CK(mem_alloc((void **) &c->d_dist, M * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->d_weight, M * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->d_crit, M * ps, MEM_DEVICE));
    CK(mem_alloc((void **) &c->d_keys, M * 8, MEM_DEVICE));
    CK(mem_alloc((void **) &c->d_candK, (size_t) SEL_CTAS * SEL_MAXK * 8, MEM_DEVICE)); CK(mem_alloc((void **) &c->d_candI, (size_t) SEL_CTAS * SEL_MAXK * 4, MEM_DEVICE));
    CK(cudaMemsetAsync(c->codes, VFT_NOCODE, (M + S) * Lp, c->stream));
    CK(cudaMemsetAsync(c->weights, 0, (N + S) * Lp * ps, c->stream));
    CK(cudaMemsetAsync(c->ow, 0, Lp * ps, c->stream));
    CK(cudaMemsetAsync(c->ov, 0, Lp * A * ps, c->stream));
    CK(cudaMemsetAsync(c->ocd, 0, Lp * A * ps, c->stream));
    CK(cudaMemsetAsync(c->diameter, 0, M * ps, c->stream)); CK(cudaMemsetAsync(c->selfdist, 0, M * ps, c->stream));
    CK(cudaMemsetAsync(c->selfweight, 0, M * ps, c->stream)); CK(cudaMemsetAsync(c->outDist, 0, M * ps, c->stream));
    CK(cudaMemsetAsync(c->active, 0, M, c->stream));
    CK(cudaMemsetAsync(c->tables, 0, 840 * ps, c->stream));
    c->activeHost.assign(M, 0);
    CK(mem_alloc((void **) &c->d_act, (M + 64) * 4, MEM_DEVICE)); CK(mem_alloc((void **) &c->h_act, (M + 64) * 4, MEM_PINNED));
    c->sharded = g_dist.ready && g_dist.world > 1 && g_dist.device == cfg->device;
    if (c->sharded) g_dist.liveContexts++;
    CK(mem_alloc((void **) &c->mlTables, 1300 * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->mlRates, 64 * ps, MEM_DEVICE)); CK(mem_alloc((void **) &c->mlRatecat, Lp * 4, MEM_DEVICE));
    CK(cudaMemsetAsync(c->mlRatecat, 0, Lp * 4, c->stream));
    c->hasTransmat = false; c->hasRates = false;
    CK(mem_alloc((void **) &c->d_doneCount, 4, MEM_DEVICE));
    CK(mem_alloc(&c->ow2, Lp * ps, MEM_DEVICE)); CK(mem_alloc(&c->ov2, Lp * A * ps, MEM_DEVICE)); CK(mem_alloc(&c->ocd2, Lp * A * ps, MEM_DEVICE));
    CK(cudaMemsetAsync(c->ow2, 0, Lp * ps, c->stream)); CK(cudaMemsetAsync(c->ov2, 0, Lp * A * ps, c->stream)); CK(cudaMemsetAsync(c->ocd2, 0, Lp * A * ps, c->stream));
    CK(mem_alloc(&c->d_specR0, SPEC_MAX * ps, MEM_DEVICE)); CK(mem_alloc(&c->d_specR1, SPEC_MAX * ps, MEM_DEVICE)); CK(mem_alloc(&c->d_specSelf, 2 * ps, MEM_DEVICE));
    CK(mem_alloc(&c->h_specIn, SPEC_MAX * 8, MEM_PINNED)); CK(mem_alloc(&c->h_specOut, 2 * SPEC_MAX * ps + 16, MEM_PINNED));
    CK(cudaEventCreateWithFlags(&c->specDone, cudaEventDisableTiming));
    { void *f = nullptr; CK(mem_alloc(&f, 64, MEM_PINNED)); c->h_flag = (volatile unsigned int *) f; *c->h_flag = 0; }
    CK(mem_alloc((void **) &c->d_terms, 2 * Lp * 8, MEM_DEVICE));
    CK(mem_alloc((void **) &c->d_selfFlag, 4, MEM_DEVICE)); CK(cudaMemsetAsync(c->d_selfFlag, 0, 4, c->stream));
    if (const char *e = std::getenv("VFT_SELF_DEFER")) c->deferSelf = e[0] != '0';
    CK(cudaMemsetAsync(c->d_doneCount, 0, 4, c->stream));
    int rc = ensure_lists(c, std::max<int64_t>(4096, c->M));
    // pinned request/response buffers sized once for the largest list the NJ driver produces
    // (m lists of 2m pairs at a refresh, m = sqrt(N); every active node in the all-node sweeps)
    if (rc == VFT_OK) rc = ensure_pinned(c, std::max<size_t>((size_t) 1 << 20, (size_t) 48 * (size_t) c->N + (size_t) 16 * (size_t) c->M + 4096));
    if (rc != VFT_OK) return rc;
    {
        // the grouped distance kernels use up to 4 warps x R rows of the [R][C] term tile
#define SET_SMEM(P, A_, MX) do { const int need = (int) (4 * group_smem_bytes<P, A_, MX>(TileShape<A_, MX>::R)); \
        cudaFuncSetAttribute(k_eval<P, A_, MX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need); \
        cudaFuncSetAttribute(k_eval<P, A_, MX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need); \
        cudaFuncSetAttribute(k_one_vs_all_warp<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, need); \
        cudaFuncSetAttribute(k_out_distance_all<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, need); } while (0)
        VFT_DISPATCH(c, SET_SMEM);
#define SET_SMEM_WIDE(P, A_, MX) do { const size_t w = std::max(wide_smem_bytes<P, A_, MX>(c->Lp), (size_t) c->Lp * 16);   /* (the deferred self distance's CTA: two term rows) */ \
        c->wideOk = w <= 200 * 1024; \
        if (c->wideOk && w > 48 * 1024) cudaFuncSetAttribute(k_eval_wide<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) w); \
        int nSm = 148; cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, cfg->device); \
        for (int k = 0; c->wideOk && k < 6; k++) { int nb = 0; \
            if (WIDE_THREADS[k] <= (sizeof(P) == 4 ? 384 : 256) && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_eval_wide<P, A_, MX>, WIDE_THREADS[k], w) == cudaSuccess) c->wideSlots[k] = nb * nSm; } } while (0)
        VFT_DISPATCH(c, SET_SMEM_WIDE);
        // TMA-staged sweeps: tiles of 16 warps + the staged rows (vectors + weights + codes of one node, or codeDist + weights)
        const size_t rows = (size_t) c->Lp * (20 * 4 + 4 + 1) + 64;
        c->stagedSmem = staged_tile_bytes<float, 20, true>() + ((rows + 15) & ~(size_t) 15);
        c->stagedOk = c->ps == 4 && c->A == 20 && c->cfg.useMatrix && c->stagedSmem <= 225 * 1024 && std::getenv("VFT_STAGING") != nullptr && std::getenv("VFT_STAGING")[0] == '1';      // opt-in: measured 15 % SLOWER than the plain sweeps
        if (c->stagedOk) {
            cudaFuncSetAttribute(k_out_distance_all_staged<float, 20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->stagedSmem);
            cudaFuncSetAttribute(k_one_vs_all_staged<float, 20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->stagedSmem);
        }
    }
    {
        // the table-driven sweeps (vft_sweep.cuh) replace the generic grouped kernels in the 20-state matrix mode; VFT_SWEEP=0 keeps the old ones
        const char *e = std::getenv("VFT_SWEEP");
        c->sweepOk = c->A == 20 && c->cfg.useMatrix && !(e && e[0] == '0');
        if (const char *m = std::getenv("VFT_SWEEP_MODES")) c->sweepModes = std::atoi(m) & 7;
        if (c->sweepOk) {
#define SET_SWEEP(P) do { cudaFuncSetAttribute(k_sweep20<P, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SweepCfg<P>::bytes); \
            cudaFuncSetAttribute(k_sweep20<P, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SweepCfg<P>::bytes); \
            cudaFuncSetAttribute(k_sweep20<P, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SweepCfg<P>::bytes); } while (0)
            if (c->ps == 4) SET_SWEEP(float); else SET_SWEEP(double);
        }
    }
    cudaFuncSetAttribute(k_topk_select<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 256 * 4 + SEL_MAXK * 12);
    cudaFuncSetAttribute(k_topk_select<double, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 256 * 4 + SEL_MAXK * 12);
    CK(sync_stream(c));
    *out = c;
    return VFT_OK;
}

extern "C" int vft_ctx_destroy(vft_ctx *c);
extern "C" int vft_ctx_create(const vft_config *cfg, vft_ctx **out) {
    vft_ctx *partial = nullptr;
    const int rc = ctx_create_impl(cfg, out, &partial);
    if (rc != VFT_OK && partial != nullptr) {        // a failed allocation / CUDA call half way: give everything back
        char keep[sizeof g_err];
        std::memcpy(keep, g_err, sizeof keep);
        vft_ctx_destroy(partial);
        std::memcpy(g_err, keep, sizeof keep);
        if (out) *out = nullptr;
    }
    return rc;
}

extern "C" int vft_ctx_destroy(vft_ctx *c) {
    if (!c) return VFT_OK;
    bind_device(c);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->sharded && g_dist.liveContexts > 0) g_dist.liveContexts--;
    void *ptrs[] = {c->codes, c->weights, c->vecs, c->ow, c->ov, c->ocd, c->diameter, c->selfdist, c->selfweight,
                    c->outDist, c->active, c->tables, c->d_dist, c->d_weight, c->d_crit, c->d_keys, c->d_ids, c->d_pi, c->d_pj, c->d_out1, c->d_out2, c->mlTables, c->mlRates, c->mlRatecat};
    for (void *p : ptrs) mem_free(p);
    mem_free(c->h_in); mem_free(c->h_out); mem_free(c->d_act); mem_free(c->h_act); mem_free(c->d_selfFlag); mem_free(c->d_qcd); mem_free(c->d_qv); mem_free(c->d_qw); mem_free(c->d_qnodes);
    mem_free(c->d_doneCount); mem_free(c->d_mrg); mem_free(c->d_acct); mem_free(c->d_terms); mem_free(c->d_candK); mem_free(c->d_candI);
    for (void *q : {c->ow2, c->ov2, c->ocd2, c->d_specR0, c->d_specR1, c->d_specSelf, c->h_specIn, c->h_specOut}) mem_free(q);
    if (c->specDone) cudaEventDestroy(c->specDone);
    mem_free((void *) c->h_flag);
    for (auto &p : c->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : c->pool) cudaEventDestroy(e);
    if (c->tmr0) cudaEventDestroy(c->tmr0);
    if (c->tmr1) cudaEventDestroy(c->tmr1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return VFT_OK;
}

extern "C" int vft_upload_tables(vft_ctx *c, const void *distances, const void *eigenval, const void *eigentot,
                                 const void *codeFreq) {
    if (!c || !distances || !eigenval || !eigentot || !codeFreq) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    const size_t ps = c->ps;
    char *h = (char *) c->h_in;
    std::memcpy(h, distances, 400 * ps); std::memcpy(h + 400 * ps, eigenval, 20 * ps);
    std::memcpy(h + 420 * ps, eigentot, 20 * ps); std::memcpy(h + 440 * ps, codeFreq, 400 * ps);
    CK(cudaMemcpyAsync(c->tables, h, 840 * ps, cudaMemcpyHostToDevice, c->stream));
    CK(sync_stream(c));
    return VFT_OK;
}

extern "C" int vft_upload_leaves(vft_ctx *c, const uint8_t *codes) {
    if (!c || !codes) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    // clamp unknowns to NOCODE on the way (NJ.tcc:449-452) into the padded rows
    std::vector<uint8_t> padded((size_t) c->N * c->Lp, VFT_NOCODE);
    for (int64_t i = 0; i < c->N; i++)
        for (int64_t p = 0; p < c->L; p++) {
            uint8_t cd = codes[i * c->L + p];
            padded[(size_t) i * c->Lp + p] = cd >= c->A ? VFT_NOCODE : cd;
        }
    CK(cudaMemcpyAsync(c->codes, padded.data(), padded.size(), cudaMemcpyHostToDevice, c->stream));
    c->cnt.h2dBytes += (int64_t) padded.size();
    CK(cudaMemsetAsync(c->active, 0, c->M, c->stream));
#define CALL_INIT(P, A_, MX) k_init_leaves<P><<<(unsigned) ((c->N + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c))
    if (c->cfg.precision == 32) { CALL_INIT(float, 0, 0); } else { CALL_INIT(double, 0, 0); }
    CK(cudaGetLastError());
    CK(sync_stream(c));
    c->cnt.launches++;
    std::fill(c->activeHost.begin(), c->activeHost.end(), 0);
    std::fill(c->activeHost.begin(), c->activeHost.begin() + c->N, 1);
    c->maxnode = c->N;
    c->nActLeaf = c->N; c->nActInternal = 0; c->actDirty = true;
    return VFT_OK;
}

extern "C" int vft_outprofile_rebuild(vft_ctx *c, const int64_t *ids, int64_t n) {
    if (!c) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    std::vector<int64_t> own;
    if (!ids) {
        for (int64_t i = 0; i < c->maxnode; i++) if (c->activeHost[i]) own.push_back(i);
        ids = own.data(); n = (int64_t) own.size();
    }
    if (n < 1) return fail(VFT_EINVAL, "no profiles to average");
    int rc = ensure_lists(c, n); if (rc) return rc;
    rc = ensure_pinned(c, (size_t) n * 8); if (rc) return rc;
    std::memcpy(c->h_in, ids, (size_t) n * 8);
    CK(cudaMemcpyAsync(c->d_ids, c->h_in, (size_t) n * 8, cudaMemcpyHostToDevice, c->stream));
    c->cnt.h2dBytes += n * 8;
#define CALL_REB(P, A_, MX)                                                                                          \
    do {                                                                                                             \
        cudaFuncSetAttribute(k_outprofile_rebuild<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) rebuild_smem_bytes<P, A_>()); \
        k_outprofile_rebuild<P, A_, MX><<<(unsigned) ((c->L + RebShape<A_>::PP - 1) / RebShape<A_>::PP), REB_T, rebuild_smem_bytes<P, A_>(), c->stream>>>(make_store<P>(c), c->d_ids, n); \
    } while (0)
    prof_begin(c, CLS_PROFILE, K_REBUILD);
    VFT_DISPATCH(c, CALL_REB);
    prof_end(c);
    CK(cudaGetLastError());
    CK(sync_stream(c));     // h_in is reused by the next call
    c->cnt.launches++;
    return VFT_OK;
}

extern "C" int vft_outprofile_update(vft_ctx *c, int64_t old1, int64_t old2, int64_t newnode, int64_t nActiveOld) {
    if (!c || nActiveOld < 2 || newnode < c->N || newnode >= c->maxnode || old1 < 0 || old2 < 0 || old1 >= c->maxnode
        || old2 >= c->maxnode)
        return fail(VFT_EINVAL, "bad argument");
    bind_device(c);
#define CALL_UPD(P, A_, MX) k_outprofile_update<P, A_, MX><<<(unsigned) ((c->L + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c), old1, old2, newnode, nActiveOld)
    prof_begin(c, CLS_PROFILE, K_OUTPROFILE_UPDATE);
    VFT_DISPATCH(c, CALL_UPD);
    prof_end(c);
    CK(cudaGetLastError());
    c->cnt.launches++;
    return VFT_OK;            // asynchronous: ordered on the context's stream
}

static int launch_average(vft_ctx *c, int64_t out_id, int64_t id1, int64_t id2, double bionjWeight, double diameter_out,
                          int64_t nActiveOld, bool update) {
    if (!c || out_id < c->N || out_id >= c->M || id1 < 0 || id2 < 0 || id1 >= c->maxnode || id2 >= c->maxnode)
        return fail(VFT_EINVAL, "bad node id");
    bind_device(c);
    if (update && nActiveOld < 2) return fail(VFT_EINVAL, "bad nActiveOld");
    if (bionjWeight < 0) bionjWeight = 0.5;
    const size_t smem = (size_t) c->Lp * 16;
    if (smem > 200 * 1024) return fail(VFT_EINVAL, "alignment too long for the average kernel's term buffer");
    // short alignments: one CTA (no cross-CTA hand-off); long ones: 64 positions per CTA, the last CTA adds the terms
    int AVG_T = c->Lp <= 512 ? 256 : 64, ppc = 0;
    unsigned avgBlocks = c->Lp <= 512 ? 1u : (unsigned) ((c->Lp + AVG_T - 1) / AVG_T);
    if (update && c->cfg.useMatrix && c->Lp > 512 && !(std::getenv("VFT_AVG_SPLIT") && std::getenv("VFT_AVG_SPLIT")[0] == '0')) {
        AVG_T = 256; ppc = 16; avgBlocks = (unsigned) ((c->Lp + ppc - 1) / ppc);      // codeDist shared by all threads of the CTA
    }
    // joins of long alignments: the self distance leaves this kernel's tail (see vft_ctx::pendingSelf)
    const int noSelf = (update && c->deferSelf && c->wideOk && c->Lp >= 512) ? 1 : 0;
#define CALL_AVG(P, A_, MX)                                                                                   \
    do {                                                                                                      \
        if (update) {                                                                                         \
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_average<P, A_, MX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
            k_average<P, A_, MX, true><<<avgBlocks, AVG_T, smem, c->stream>>>(make_store<P>(c), out_id, id1, id2, bionjWeight, (P) diameter_out, nActiveOld, c->d_terms, c->d_doneCount, (P *) c->ow, (P *) c->ov, (P *) c->ocd, (P *) nullptr, nullptr, ppc, noSelf); \
        } else {                                                                                              \
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_average<P, A_, MX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
            k_average<P, A_, MX, false><<<avgBlocks, AVG_T, smem, c->stream>>>(make_store<P>(c), out_id, id1, id2, bionjWeight, (P) diameter_out, nActiveOld, c->d_terms, c->d_doneCount, (P *) c->ow, (P *) c->ov, (P *) c->ocd, (P *) nullptr); \
        }                                                                                                     \
    } while (0)
    prof_begin(c, CLS_PROFILE, K_AVERAGE);
    VFT_DISPATCH(c, CALL_AVG);
    prof_end(c);
    CK(cudaGetLastError());
    c->cnt.launches++; c->cnt.profileAvgOps++; c->cnt.profileOps++;
    for (int64_t ch : {id1, id2})
        if (c->activeHost[ch]) { c->activeHost[ch] = 0; if (ch < c->N) c->nActLeaf--; else c->nActInternal--; }
    if (!c->activeHost[out_id]) { c->activeHost[out_id] = 1; c->nActInternal++; }
    if (out_id >= c->maxnode) c->maxnode = out_id + 1;
    c->actDirty = true;
    if (noSelf) c->pendingSelf = out_id;
    return VFT_OK;            // asynchronous: ordered on the context's stream
}

extern "C" int vft_profile_average(vft_ctx *c, int64_t out_id, int64_t id1, int64_t id2, double bionjWeight,
                                   double diameter_out) {
    return launch_average(c, out_id, id1, id2, bionjWeight, diameter_out, 0, false);
}

extern "C" int vft_profile_average_update(vft_ctx *c, int64_t out_id, int64_t id1, int64_t id2, double bionjWeight,
                                          double diameter_out, int64_t nActiveOld) {
    return launch_average(c, out_id, id1, id2, bionjWeight, diameter_out, nActiveOld, true);
}

extern "C" int vft_profile_average_batch(vft_ctx *c, int64_t n, const int64_t *out_id, const int64_t *id1, const int64_t *id2) {
    if (!c || n < 0 || (n > 0 && (!out_id || !id1 || !id2))) return fail(VFT_EINVAL, "null argument");
    if (n == 0) return VFT_OK;
    bind_device(c);
    int rc = ensure_pinned(c, (size_t) n * 12); if (rc) return rc;
    rc = ensure_lists(c, (3 * n + 1) / 2 + 1); if (rc) return rc;      // three int32 per item in the 8-byte-per-entry list buffer
    int32_t *h = (int32_t *) c->h_in;
    for (int64_t k = 0; k < n; k++) {
        // items of one call must be independent: every input row is a leaf or an internal row that is not an output of the call
        if (out_id[k] < c->N || out_id[k] >= c->M || id1[k] < 0 || id2[k] < 0 || id1[k] >= c->M || id2[k] >= c->M) return fail(VFT_EINVAL, "bad node id");
        h[k] = (int32_t) out_id[k]; h[n + k] = (int32_t) id1[k]; h[2 * n + k] = (int32_t) id2[k];
    }
    CK(cudaMemcpyAsync(c->d_pi, h, (size_t) n * 12, cudaMemcpyHostToDevice, c->stream));
    c->cnt.h2dBytes += n * 12;
    const int32_t *d = (const int32_t *) c->d_pi;
    prof_begin(c, CLS_PROFILE, K_AVERAGE);
    for (int64_t base = 0; base < n; base += 32768) {
        const unsigned ny = (unsigned) std::min<int64_t>(32768, n - base);
#define CALL_AVGB(P, A_, MX) k_average_batch<P, A_, MX><<<dim3((unsigned) ((c->Lp + 127) / 128), ny), 128, 0, c->stream>>>(make_store<P>(c), d, d + n, d + 2 * n, base)
        VFT_DISPATCH(c, CALL_AVGB);
        c->cnt.launches++;
    }
    prof_end(c);
    CK(cudaGetLastError());
    CK(sync_stream(c));                    // h_in is reused by the next call
    c->cnt.profileAvgOps += n;
    for (int64_t k = 0; k < n; k++) if (out_id[k] >= c->maxnode) c->maxnode = out_id[k] + 1;
    return VFT_OK;
}

extern "C" int vft_get_self(vft_ctx *c, int64_t id, double *selfdist, double *selfweight) {
    if (!c || id < 0 || id >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
    bind_device(c);
    char buf[16];
    CK(cudaMemcpyAsync(buf, (char *) c->selfdist + id * c->ps, c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(buf + 8, (char *) c->selfweight + id * c->ps, c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(sync_stream(c));
    if (c->ps == 4) { *selfdist = *(float *) buf; *selfweight = *(float *) (buf + 8); }
    else { *selfdist = *(double *) buf; *selfweight = *(double *) (buf + 8); }
    return VFT_OK;
}

static int64_t profile_bytes(vft_ctx *c, int64_t id) {      // algorithmic bytes of one node's profile
    return id >= 0 && id < c->N ? c->L : c->L * ((int64_t) c->A * c->ps + c->ps + 1);
}

static int pick_group(vft_ctx *c, int64_t nItems) {
    // items per warp: 1 while the batch cannot fill the machine, more once it can (148 SMs x 16 warps),
    // up to the rows of the term tile (TileShape<A, MATRIX>::R)
    const int R = (c->A == 4 && !c->cfg.useMatrix) ? 8 : 32;
    int64_t g = (nItems + 2367) / 2368;
    return (int) std::min<int64_t>(R, std::max<int64_t>(1, g));
}

// One request = one launch + one synchronisation.  Inputs are packed as int32 into pinned host
// memory that the kernel reads directly (zero-copy); results come back the same way.
extern "C" int vft_eval_batch(vft_ctx *c, const int64_t *out_ids, int64_t nOut, int64_t nActive, double totdiam,
                              void *outDist, const int64_t *pi, const int64_t *pj, int64_t nPairs, int32_t flags,
                              void *dist, void *weight) {
    if (!c) return fail(VFT_EINVAL, "null argument");
    bind_device(c, /*keepPendingSelf=*/true);
    if ((nOut > 0 && (!out_ids || !outDist)) || (nPairs > 0 && (!pi || !pj || !dist || !weight)) || nOut < 0 || nPairs < 0)
        return fail(VFT_EINVAL, "null argument");
    const int64_t n = nOut + nPairs;
    if (n == 0) return VFT_OK;
    BytesScope bytesScope(c, n <= INLINE_ITEMS ? K_EVAL_SMALL : K_EVAL_LARGE);
    int rc = ensure_pinned(c, (size_t) n * 16); if (rc) return rc;
    int32_t *ha = (int32_t *) c->h_in, *hb = ha + n;
    const bool raw = (flags & VFT_PAIRS_PROFILE_RAW) != 0;
    for (int64_t k = 0; k < nOut; k++) {
        if (out_ids[k] < 0 || out_ids[k] >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
        ha[k] = (int32_t) out_ids[k]; hb[k] = -1;
        c->cnt.algoBytes += profile_bytes(c, out_ids[k]);
    }
    if (nOut) { c->cnt.algoBytes += profile_bytes(c, -1); c->cnt.profileOps += nOut; c->cnt.outprofileOps += nOut; }
    int64_t lastQuery = -2;
    for (int64_t k = 0; k < nPairs; k++) {
        // j == -1 with VFT_PAIRS_PROFILE_RAW: the out-profile (bare profileDist(i, outprofile), NJ.tcc:2945-2946)
        if (pi[k] < 0 || (pj[k] < 0 && !(raw && pj[k] == -1)) || pi[k] >= c->maxnode || pj[k] >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
        ha[nOut + k] = (int32_t) pi[k]; hb[nOut + k] = (int32_t) pj[k];
        if (!raw && pi[k] < c->N && pj[k] < c->N) { c->cnt.seqOps++; c->cnt.algoBytes += c->L; }
        else { c->cnt.profileOps++; c->cnt.algoBytes += profile_bytes(c, pj[k]); }
        if (pi[k] != lastQuery) { c->cnt.algoBytes += profile_bytes(c, pi[k]); lastQuery = pi[k]; }   // a list shares its query
    }
    const int G = pick_group(c, n);
    const int64_t warps = (n + G - 1) / G;
    const unsigned blocks = (unsigned) ((warps + 3) / 4);
    rc = ensure_lists(c, n); if (rc) return rc;
    void *r0 = c->d_out1, *r1 = c->d_out2;                    // device-side results; the kernel's last CTA copies them out
    void *hr0 = c->h_out, *hr1 = (char *) c->h_out + (size_t) n * c->ps;
    const bool inlineItems = n <= INLINE_ITEMS;
    InlineItems inl;
    if (inlineItems) { std::memcpy(inl.a, ha, (size_t) n * 4); std::memcpy(inl.b, hb, (size_t) n * 4); }
    const int32_t *qa = ha, *qb = hb;                         // mid-sized lists: read by the kernel from the mapped buffer
    const bool dma = n > 4096;                                // big batches: requests and results travel by DMA
    if (dma) {
        CK(cudaMemcpyAsync(c->d_pi, ha, (size_t) n * 8, cudaMemcpyHostToDevice, c->stream));
        qa = (const int32_t *) c->d_pi; qb = qa + n;
    }
    void *hostOut = dma ? nullptr : hr0;
    // small and mid-sized lists: the kernel's last CTA writes a sequence number behind the results and the host spins on
    // it -- ~2.5 us sooner than cudaStreamSynchronize notices the same completion (not while events are being resolved)
    const bool spin = !dma && !c->profile;
    const unsigned int seq = ++c->flagSeq;
    volatile unsigned int *flag = spin ? c->h_flag : nullptr;
#define EVAL_ARGS(P) make_store<P>(c), inl, inlineItems ? nullptr : qa, inlineItems ? nullptr : qb, n, nOut, G, raw ? 1 : 0, nActive, totdiam, (P *) r0, (P *) r1, c->d_doneCount, (P *) hostOut, flag, seq
#define CALL_EVAL(P, A_, MX) do { if (G > 1) k_eval<P, A_, MX, true><<<blocks, 128, 4 * group_smem_bytes<P, A_, MX>(G), c->stream>>>(EVAL_ARGS(P)); \
        else k_eval<P, A_, MX, false><<<blocks, 128, 4 * group_smem_bytes<P, A_, MX>(G), c->stream>>>(EVAL_ARGS(P)); } while (0)
    // long alignments, lists that cannot fill the machine with a warp per pair: a CTA per pair
    const bool wide = c->wideOk && c->Lp >= 512 && n <= 2048;
    // threads per pair: the largest CTA for which ALL n CTAs are co-resident (registers decide: 154 per thread in the fp32
    // 20-state build = 12 warps per SM).  One wave matters more than warps per pair: the list's latency is one CTA's
    // (chunks per warp + the ordered sum), whereas a second wave doubles it.
    // a pending self distance rides along as one more CTA -- only when every CTA (items + 1) is co-resident, because an item may
    // wait for it; otherwise it is resolved by its own kernel first
    int64_t selfNode = -1;
    if (c->pendingSelf >= 0) {
        if (wide && c->wideSlots[5] >= n + 1) { selfNode = c->pendingSelf; c->pendingSelf = -1; c->selfSeq++; }
        else resolve_pending_self(c);
    }
    int wideThreads = 32;
    for (int k = 0; k < 6; k++) if (c->wideSlots[k] >= n + (selfNode >= 0 ? 1 : 0)) { wideThreads = WIDE_THREADS[k]; break; }
    if (const char *e = std::getenv("VFT_WIDE_THREADS")) { const int t = std::atoi(e); if (t >= 32 && t <= (c->ps == 4 ? 384 : 256) && t % 32 == 0) wideThreads = t; }
#define CALL_EVAL_WIDE(P, A_, MX) k_eval_wide<P, A_, MX><<<(unsigned) (n + (selfNode >= 0 ? 1 : 0)), wideThreads, std::max(wide_smem_bytes<P, A_, MX>(c->Lp), (size_t) c->Lp * 16), c->stream>>>(make_store<P>(c), inl, inlineItems ? nullptr : qa, inlineItems ? nullptr : qb, n, nOut, raw ? 1 : 0, nActive, totdiam, (P *) r0, (P *) r1, c->d_doneCount, (P *) hostOut, flag, seq, selfNode, c->d_terms, c->d_selfFlag, c->selfSeq)
    prof_begin(c, CLS_DIST, n <= INLINE_ITEMS ? K_EVAL_SMALL : K_EVAL_LARGE);
    if (wide) { VFT_DISPATCH(c, CALL_EVAL_WIDE); } else { VFT_DISPATCH(c, CALL_EVAL); }
    prof_end(c);
    CK(cudaGetLastError());
    if (dma) {
        CK(cudaMemcpyAsync(hr0, r0, (size_t) n * c->ps, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(hr1, r1, (size_t) n * c->ps, cudaMemcpyDeviceToHost, c->stream));
    }
    if (spin) {
        // bounded: a kernel that faulted never writes the word; the stream synchronisation then reports the error
        bool seen = false;
        for (long it = 0; it < 400000000L; it++) { if (*flag == seq) { seen = true; break; } cpu_relax(); }
        if (!seen) CK(sync_stream(c));
    } else CK(sync_stream(c));
    c->cnt.launches++;
    c->cnt.h2dBytes += n * 8; c->cnt.d2hBytes += n * 2 * (int64_t) c->ps;
    if (nOut) std::memcpy(outDist, hr0, (size_t) nOut * c->ps);
    if (nPairs) {
        std::memcpy(dist, (char *) hr0 + (size_t) nOut * c->ps, (size_t) nPairs * c->ps);
        std::memcpy(weight, (char *) hr1 + (size_t) nOut * c->ps, (size_t) nPairs * c->ps);
    }
    return VFT_OK;
}

// ---- speculative join ---------------------------------------------------------------------------------------------
// The join loop is a chain: search -> average the pair -> distances of the new node to its candidates -> bookkeeping ->
// search ...; the device idles while the host decides and the host idles while the device computes.  The caller's guess
// of the NEXT join is right ~95 % of the time (nj_host.cpp), so the next join's device work is launched ahead,
// asynchronously and WITHOUT committing anything: the new profile goes into the next free row, the updated out-profile
// into a shadow copy, every distance comes back raw (bare profileDist: the diameter / out-distance algebra needs host
// scalars that are not known yet and is finished by the caller in the same arithmetic).  take() makes it real (state
// kernel + pointer swap), discard() forgets it; a wrong guess costs nothing but the idle device time it used.
extern "C" int vft_spec_join_launch(vft_ctx *c, int64_t out_id, int64_t id1, int64_t id2, double bionjWeight, int64_t nActiveOld,
                                    const int64_t *pair_j, int64_t nPairs, const int64_t *out_ids, int64_t nOut) {
    if (!c || nPairs < 0 || nOut < 0 || (nPairs > 0 && !pair_j) || (nOut > 0 && !out_ids)) return fail(VFT_EINVAL, "null argument");
    const int64_t n = nOut + nPairs;
    if (n < 1 || (size_t) n > SPEC_MAX) return fail(VFT_EINVAL, "speculative request too large");
    if (out_id != c->maxnode || out_id >= c->M || id1 < 0 || id2 < 0 || id1 >= c->maxnode || id2 >= c->maxnode || id1 == id2
        || !c->activeHost[id1] || !c->activeHost[id2] || nActiveOld < 3)
        return fail(VFT_EINVAL, "bad speculative join");
    bind_device(c);
    c->specPending = false;
    if (bionjWeight < 0) bionjWeight = 0.5;
    const size_t smemAvg = (size_t) c->Lp * 16;
    if (smemAvg > 200 * 1024) return fail(VFT_EINVAL, "alignment too long for the average kernel's term buffer");
    int32_t *ha = (int32_t *) c->h_specIn, *hb = ha + n;
    int64_t bytes = profile_bytes(c, out_id);
    for (int64_t k = 0; k < nOut; k++) {
        if (out_ids[k] < 0 || (out_ids[k] >= c->maxnode && out_ids[k] != out_id)) return fail(VFT_EINVAL, "bad node id");
        ha[k] = (int32_t) out_ids[k]; hb[k] = -1;
        bytes += profile_bytes(c, out_ids[k]);
    }
    if (nOut) bytes += profile_bytes(c, -1);
    for (int64_t k = 0; k < nPairs; k++) {
        if (pair_j[k] < 0 || pair_j[k] >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
        ha[nOut + k] = (int32_t) out_id; hb[nOut + k] = (int32_t) pair_j[k];
        bytes += profile_bytes(c, pair_j[k]);
    }
    const int AVG_T = c->Lp <= 512 ? 256 : 64;
    const unsigned avgBlocks = c->Lp <= 512 ? 1u : (unsigned) ((c->Lp + AVG_T - 1) / AVG_T);
#define CALL_SPEC_AVG(P, A_, MX)                                                                              \
    do {                                                                                                      \
        if (smemAvg > 48 * 1024) cudaFuncSetAttribute(k_average<P, A_, MX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemAvg); \
        k_average<P, A_, MX, true><<<avgBlocks, AVG_T, smemAvg, c->stream>>>(make_store<P>(c), out_id, id1, id2, bionjWeight, (P) 0, nActiveOld, c->d_terms, c->d_doneCount, (P *) c->ow2, (P *) c->ov2, (P *) c->ocd2, (P *) c->d_specSelf); \
    } while (0)
    prof_begin(c, CLS_PROFILE, K_AVERAGE);
    VFT_DISPATCH(c, CALL_SPEC_AVG);
    prof_end(c);
    CK(cudaGetLastError());
    // the distances, against the SHADOW out-profile, all raw
    const int G = pick_group(c, n);
    const int64_t warps = (n + G - 1) / G;
    const unsigned blocks = (unsigned) ((warps + 3) / 4);
    const bool inlineItems = n <= INLINE_ITEMS;
    InlineItems inl;
    if (inlineItems) { std::memcpy(inl.a, ha, (size_t) n * 4); std::memcpy(inl.b, hb, (size_t) n * 4); }
    const bool wide = c->wideOk && c->Lp >= 512 && n <= 2048;
    const int wideThreads = n <= 160 ? 256 : 128;
#define SPEC_STORE(P) ([&] { Store<P> st = make_store<P>(c); st.ow = (P *) c->ow2; st.ov = (P *) c->ov2; st.ocd = c->cfg.useMatrix ? (P *) c->ocd2 : nullptr; return st; }())
#define SPEC_ARGS(P) SPEC_STORE(P), inl, inlineItems ? nullptr : ha, inlineItems ? nullptr : hb, n, nOut
#define CALL_SPEC_EVAL(P, A_, MX) do { if (wide) k_eval_wide<P, A_, MX><<<(unsigned) n, wideThreads, wide_smem_bytes<P, A_, MX>(c->Lp), c->stream>>>(SPEC_ARGS(P), 1, nActiveOld - 1, 0.0, (P *) c->d_specR0, (P *) c->d_specR1, c->d_doneCount, (P *) c->h_specOut); \
        else if (G > 1) k_eval<P, A_, MX, true><<<blocks, 128, 4 * group_smem_bytes<P, A_, MX>(G), c->stream>>>(SPEC_ARGS(P), G, 1, nActiveOld - 1, 0.0, (P *) c->d_specR0, (P *) c->d_specR1, c->d_doneCount, (P *) c->h_specOut); \
        else k_eval<P, A_, MX, false><<<blocks, 128, 4 * group_smem_bytes<P, A_, MX>(G), c->stream>>>(SPEC_ARGS(P), G, 1, nActiveOld - 1, 0.0, (P *) c->d_specR0, (P *) c->d_specR1, c->d_doneCount, (P *) c->h_specOut); } while (0)
    prof_begin(c, CLS_DIST, n <= INLINE_ITEMS ? K_EVAL_SMALL : K_EVAL_LARGE);
    VFT_DISPATCH(c, CALL_SPEC_EVAL);
    prof_end(c);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync((char *) c->h_specOut + (size_t) 2 * n * c->ps, c->d_specSelf, 2 * c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->specDone, c->stream));
    c->cnt.launches += 2;
    c->specPending = true; c->specOut = out_id; c->specId1 = id1; c->specId2 = id2; c->specNPairs = nPairs; c->specNOut = nOut; c->specBytes = bytes;
    return VFT_OK;            // asynchronous
}

// The guess was right: (id1, id2) are joined into specOut.  pairDist/pairWeight[nPairs] = bare profileDist(out, pair_j[k]);
// outDist/outWeight[nOut] = bare profileDist(out_ids[k], NEW out-profile); self2 = {selfdist, selfweight} of the new node.
extern "C" int vft_spec_join_take(vft_ctx *c, double diameter_out, void *pairDist, void *pairWeight, void *outDist, void *outWeight,
                                  void *self2) {
    if (!c || !c->specPending) return fail(VFT_EINVAL, "no speculative join pending");
    bind_device(c);
    CK(cudaEventSynchronize(c->specDone));
    c->specPending = false;
    const int64_t out_id = c->specOut, id1 = c->specId1, id2 = c->specId2, nOut = c->specNOut, nPairs = c->specNPairs, n = nOut + nPairs;
    if (c->cfg.precision == 32) k_spec_commit<float><<<1, 1, 0, c->stream>>>(make_store<float>(c), out_id, id1, id2, (float) diameter_out, (const float *) c->d_specSelf);
    else k_spec_commit<double><<<1, 1, 0, c->stream>>>(make_store<double>(c), out_id, id1, id2, diameter_out, (const double *) c->d_specSelf);
    CK(cudaGetLastError());
    std::swap(c->ow, c->ow2); std::swap(c->ov, c->ov2); std::swap(c->ocd, c->ocd2);
    c->cnt.launches++; c->cnt.profileAvgOps++; c->cnt.profileOps += 1 + n; c->cnt.outprofileOps += nOut;
    c->cnt.algoBytes += c->specBytes; c->cnt.bytesKernel[n <= INLINE_ITEMS ? K_EVAL_SMALL : K_EVAL_LARGE] += c->specBytes;
    c->cnt.h2dBytes += n * 8; c->cnt.d2hBytes += (n * 2 + 2) * (int64_t) c->ps;
    for (int64_t ch : {id1, id2})
        if (c->activeHost[ch]) { c->activeHost[ch] = 0; if (ch < c->N) c->nActLeaf--; else c->nActInternal--; }
    if (!c->activeHost[out_id]) { c->activeHost[out_id] = 1; c->nActInternal++; }
    if (out_id >= c->maxnode) c->maxnode = out_id + 1;
    c->actDirty = true;
    const char *h0 = (const char *) c->h_specOut, *h1 = h0 + (size_t) n * c->ps;
    if (nOut && outDist) std::memcpy(outDist, h0, (size_t) nOut * c->ps);
    if (nOut && outWeight) std::memcpy(outWeight, h1, (size_t) nOut * c->ps);
    if (nPairs && pairDist) std::memcpy(pairDist, h0 + (size_t) nOut * c->ps, (size_t) nPairs * c->ps);
    if (nPairs && pairWeight) std::memcpy(pairWeight, h1 + (size_t) nOut * c->ps, (size_t) nPairs * c->ps);
    if (self2) std::memcpy(self2, h0 + (size_t) 2 * n * c->ps, 2 * c->ps);
    return VFT_OK;
}

extern "C" int vft_spec_join_discard(vft_ctx *c) {
    if (!c) return fail(VFT_EINVAL, "null argument");
    c->specPending = false;      // the kernels may still run; nothing they wrote is live
    return VFT_OK;
}

template<typename P, int KEYBYTES>
static void launch_topk(vft_ctx *c, const uint64_t *keys, int64_t n, int K, Rec<P> *out, const int32_t *list, int lstride, int loffset) {
    const size_t selSmem = 32 * 256 * 4 + SEL_MAXK * 12;
    const int nCta = (int) std::min<int64_t>(SEL_CTAS, n / 8192);
    if (nCta < 2 || (int64_t) nCta * K > n) {
        k_topk_select<P, KEYBYTES><<<1, SEL_T, selSmem, c->stream>>>(keys, n, K, (const P *) c->d_dist, (const P *) c->d_weight, (const P *) c->d_crit, out,
                                                                     (const uint32_t *) nullptr, (uint64_t *) nullptr, (uint32_t *) nullptr, list, lstride, loffset);
        c->cnt.launches++;
        return;
    }
    k_topk_select<P, KEYBYTES><<<nCta, SEL_T, selSmem, c->stream>>>(keys, n, K, (const P *) c->d_dist, (const P *) c->d_weight, (const P *) c->d_crit, (Rec<P> *) nullptr,
                                                                    (const uint32_t *) nullptr, c->d_candK, c->d_candI);
    k_topk_select<P, KEYBYTES><<<1, SEL_T, selSmem, c->stream>>>(c->d_candK, (int64_t) nCta * K, K, (const P *) c->d_dist, (const P *) c->d_weight, (const P *) c->d_crit, out,
                                                                 c->d_candI, (uint64_t *) nullptr, (uint32_t *) nullptr, list, lstride, loffset);
    c->cnt.launches += 2;
}

extern "C" int vft_out_distance_batch(vft_ctx *c, const int64_t *ids, int64_t n, int64_t nActive, double totdiam,
                                      void *outDist) {
    return vft_eval_batch(c, ids, n, nActive, totdiam, outDist, nullptr, nullptr, 0, 0, nullptr, nullptr);
}

extern "C" int vft_dist_pairs(vft_ctx *c, const int64_t *pi, const int64_t *pj, int64_t n, int32_t flags, void *dist,
                              void *weight) {
    return vft_eval_batch(c, nullptr, 0, 0, 0.0, nullptr, pi, pj, n, flags, dist, weight);
}

// room for the query tables of nTab queries
static int ensure_qtabs(vft_ctx *c, int64_t nTab) {
    if (nTab <= c->qtabCap) return VFT_OK;
    const int64_t cap = std::max<int64_t>(nTab, std::min<int64_t>(2 * c->qtabCap, nTab + 64));
    mem_free(c->d_qcd); mem_free(c->d_qv); mem_free(c->d_qw); mem_free(c->d_qnodes);
    c->d_qcd = c->d_qv = c->d_qw = nullptr; c->d_qnodes = nullptr; c->qtabCap = 0;
    const size_t row = (size_t) c->Lp * c->ps;
    CK(mem_alloc(&c->d_qcd, (size_t) cap * row * 20, MEM_DEVICE)); CK(mem_alloc(&c->d_qv, (size_t) cap * row * 20, MEM_DEVICE));
    CK(mem_alloc(&c->d_qw, (size_t) cap * row, MEM_DEVICE)); CK(mem_alloc((void **) &c->d_qnodes, (size_t) cap * 4, MEM_DEVICE));
    c->qtabCap = cap;
    return VFT_OK;
}
template<typename P>
static QTab<P> make_qtab(vft_ctx *c, bool outProfile) {
    QTab<P> q;
    if (outProfile) { q.cd = (const P *) c->ocd; q.v = (const P *) c->ov; q.w = (const P *) c->ow; }
    else { q.cd = (const P *) c->d_qcd; q.v = (const P *) c->d_qv; q.w = (const P *) c->d_qw; }
    q.stride = (size_t) c->Lp * 20;
    return q;
}
#define LAUNCH_SWEEP(P, M_, nSl, ...) k_sweep20<P, M_><<<sweep_grid(nSl), SweepCfg<P>::threads, SweepCfg<P>::bytes, c->stream>>>(__VA_ARGS__)

// the ascending active list on the device (the compact sweeps index it; a sharded context takes every W-th entry)
static int ensure_active(vft_ctx *c) {
    if (!c->actDirty) return VFT_OK;
    int64_t n = 0;
    const uint8_t *act = c->activeHost.data();
    for (int64_t i = 0; i < c->maxnode; i++) if (act[i]) c->h_act[n++] = (int32_t) i;
    CK(cudaMemcpyAsync(c->d_act, c->h_act, (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
    c->cnt.h2dBytes += n * 4;
    c->nAct = n; c->actDirty = false;
    return VFT_OK;
}
// A sweep is sharded only when every rank's share is big enough to pay for the exchange (its kernels are latency bound below
// that: half the candidates are not half the time) -- otherwise every rank evaluates all of it, with no communication.  The
// decision depends only on replicated state (the number of active nodes / request slots), so every rank takes the same one.
static inline int64_t shard_min() {
    static const int64_t v = [] { const char *e = std::getenv("VFT_SHARD_MIN"); const long x = e ? std::atol(e) : 4096; return (int64_t) (x > 0 ? x : 1); }();
    return v;
}
static inline int shard_world(const vft_ctx *c, int64_t units = -1) {
    if (!c->sharded) return 1;
    return (units >= 0 && units < (int64_t) g_dist.world * shard_min()) ? 1 : g_dist.world;
}
static inline int shard_rank(const vft_ctx *c, int W) { return (c->sharded && W > 1) ? g_dist.rank : 0; }
static inline int64_t shard_len(int64_t nAct, int W, int r) { return nAct > r ? (nAct - r + W - 1) / W : 0; }

extern "C" int vft_out_distance_all(vft_ctx *c, int64_t nActive, double totdiam, void *outDist, int64_t maxnode) {
    if (!c || !outDist || maxnode < c->maxnode) return fail(VFT_EINVAL, "bad argument");
    bind_device(c);
    const int64_t n = c->maxnode;
    BytesScope bytesScope(c, K_OUT_DIST_ALL);
    int rc = ensure_pinned(c, (size_t) n * 8); if (rc) return rc;
    rc = ensure_active(c); if (rc) return rc;
    const int W = shard_world(c, c->nAct), r = shard_rank(c, W);
    const int64_t len = shard_len(c->nAct, W, r), chunk = (c->nAct + W - 1) / W;
    const int G = pick_group(c, len);
    const int64_t warps = (len + G - 1) / G;
    void *res = nullptr;
    if (W > 1) { rc = dist_reserve(c->stream, (size_t) chunk * c->ps); if (rc) return rc; res = g_dist.send; }
#define CALL_ODA(P, A_, MX) k_out_distance_all<P, A_, MX><<<(unsigned) ((warps + 3) / 4), 128, 4 * group_smem_bytes<P, A_, MX>(G), c->stream>>>(make_store<P>(c), len, G, nActive, totdiam, c->d_act, W, r, (P *) res)
    prof_begin(c, CLS_DIST, K_OUT_DIST_ALL);
#define CALL_ODA_SWEEP(P) LAUNCH_SWEEP(P, 1, len, make_store<P>(c), make_qtab<P>(c, true), c->d_act, W, r, nullptr, nullptr, 32, len, -1, nActive, totdiam, (P *) nullptr, (P *) nullptr, (P *) nullptr, (uint64_t *) nullptr, (P *) res)
    if (len > 0) {
        if (c->sweepOk && (c->sweepModes & 2)) { if (c->ps == 4) CALL_ODA_SWEEP(float); else CALL_ODA_SWEEP(double); }
        else if (c->stagedOk) k_out_distance_all_staged<float, 20, true><<<148, STG_T, c->stagedSmem, c->stream>>>(make_store<float>(c), len, std::min(G, STG_G), nActive, totdiam, c->d_act, W, r, (float *) res);
        else { VFT_DISPATCH(c, CALL_ODA); }
    }
    prof_end(c);
    CK(cudaGetLastError());
    if (W > 1) {
        // every rank's share -> every rank's table
        const char *base; size_t stride;
        rc = dist_allgather(c->stream, (size_t) chunk * c->ps, &base, &stride); if (rc) return rc;
        if (c->ps == 4) k_scatter_outdist<float><<<(unsigned) ((c->nAct + 255) / 256), 256, 0, c->stream>>>(base, stride, c->d_act, c->nAct, W, (float *) c->outDist);
        else k_scatter_outdist<double><<<(unsigned) ((c->nAct + 255) / 256), 256, 0, c->stream>>>(base, stride, c->d_act, c->nAct, W, (double *) c->outDist);
        CK(cudaGetLastError());
        c->cnt.launches += g_dist.mode == DIST_PEER ? 2 : 1;
    }
    CK(cudaMemcpyAsync(c->h_out, c->outDist, (size_t) n * c->ps, cudaMemcpyDeviceToHost, c->stream));   // one DMA instead of a PCIe write per lane
    CK(sync_stream(c));
    for (int64_t k = 0; k < c->nAct; k++) {
        const int64_t i = c->h_act[k];
        std::memcpy((char *) outDist + i * c->ps, (char *) c->h_out + i * c->ps, c->ps);
        if (k % W == r) c->cnt.algoBytes += profile_bytes(c, i);                        // (a sharded context accounts its share)
    }
    c->cnt.launches++; c->cnt.profileOps += len; c->cnt.outprofileOps += len;
    c->cnt.algoBytes += profile_bytes(c, -1);
    c->cnt.d2hBytes += c->nAct * (int64_t) c->ps;
    return VFT_OK;
}

extern "C" int vft_dist_one_vs_all(vft_ctx *c, int64_t query, int64_t nActive, int64_t K, int64_t *j_out, void *dist,
                                   void *weight, void *criterion, int64_t *nOut) {
    if (!c) return fail(VFT_EINVAL, "null argument");
    return vft_dist_one_vs_all_range(c, query, nActive, K, 0, c->maxnode, j_out, dist, weight, criterion, nOut);
}

extern "C" int vft_dist_one_vs_all_range(vft_ctx *c, int64_t query, int64_t nActive, int64_t K, int64_t jBegin, int64_t jEnd,
                                         int64_t *j_out, void *dist, void *weight, void *criterion, int64_t *nOut) {
    if (!c || !j_out || !dist || !weight || !criterion || !nOut) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    if (query < 0 || query >= c->maxnode || !c->activeHost[query]) return fail(VFT_EINVAL, "query must be an active node");
    if (K < 1) return fail(VFT_EINVAL, "K must be positive");
    BytesScope bytesScope(c, K_ONE_VS_ALL);
    const int64_t nAll = c->maxnode;
    if (K > SEL_MAXK) return fail(VFT_EINVAL, "K larger than 4096 is not supported");
    // The whole candidate range goes through the COMPACT form (slot k = k-th active node; a sharded context takes slots
    // r, r+W, ...); an explicit sub-range keeps the node-indexed form.
    const bool compact = jBegin <= 0 && jEnd >= nAll;
    int W = 1, r = 0;
    int64_t n = nAll, inBlock = nActive;
    const int32_t *list = nullptr;
    if (compact) {
        int rc = ensure_active(c); if (rc) return rc;
        W = shard_world(c, c->nAct); r = shard_rank(c, W);
        n = shard_len(c->nAct, W, r); inBlock = n; list = c->d_act;
    } else {
        inBlock = 0;
        for (int64_t j = std::max<int64_t>(0, jBegin); j < std::min(jEnd, nAll); j++) inBlock += c->activeHost[j];
    }
    const int Gq = pick_group(c, n);
    const int64_t warpsQ = (n + Gq - 1) / Gq;
#define CALL_OVA_LEAF(P, A_, MX) k_one_vs_all_leaf<P, A_, MX><<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c), query, nActive, n, jBegin, jEnd, (P *) c->d_dist, (P *) c->d_weight, (P *) c->d_crit, c->d_keys, list, W, r)
#define CALL_OVA_WARP(P, A_, MX) k_one_vs_all_warp<P, A_, MX><<<(unsigned) ((warpsQ + 3) / 4), 128, 4 * group_smem_bytes<P, A_, MX>(Gq), c->stream>>>(make_store<P>(c), query, nActive, n, jBegin, jEnd, Gq, (P *) c->d_dist, (P *) c->d_weight, (P *) c->d_crit, c->d_keys, list, W, r)
    prof_begin(c, CLS_DIST, K_ONE_VS_ALL);
#define CALL_OVA_SWEEP(P) do { \
        k_query_tables<P><<<dim3((unsigned) ((c->Lp + 127) / 128), 1), 128, 0, c->stream>>>(make_store<P>(c), nullptr, query, (P *) c->d_qcd, (P *) c->d_qv, (P *) c->d_qw); \
        LAUNCH_SWEEP(P, 0, n, make_store<P>(c), make_qtab<P>(c, false), list, W, r, nullptr, nullptr, 32, n, query, nActive, 0.0, (P *) c->d_dist, (P *) c->d_weight, (P *) c->d_crit, c->d_keys, (P *) nullptr); } while (0)
    if (n > 0 && compact && c->sweepOk && (c->sweepModes & 1)) {
        int rq = ensure_qtabs(c, 1); if (rq) return rq;
        if (c->ps == 4) CALL_OVA_SWEEP(float); else CALL_OVA_SWEEP(double);
        c->cnt.launches++;
    } else if (n > 0) {
        if (query < c->N) { VFT_DISPATCH(c, CALL_OVA_LEAF); }
        else if (c->stagedOk) k_one_vs_all_staged<float, 20, true><<<148, STG_T, c->stagedSmem, c->stream>>>(make_store<float>(c), query, nActive, n, jBegin, jEnd, std::min(Gq, STG_G), (float *) c->d_dist, (float *) c->d_weight, (float *) c->d_crit, c->d_keys, list, W, r);
        else { VFT_DISPATCH(c, CALL_OVA_WARP); }
    }
    prof_end(c);
    CK(cudaGetLastError());
    c->cnt.launches++;
    const int64_t nLocal = std::min<int64_t>(K, inBlock);                        // this rank's records
    const int64_t nRet = W > 1 ? std::min<int64_t>(K, c->nAct) : nLocal;          // records returned
    const size_t recSz = c->ps == 4 ? sizeof(Rec<float>) : sizeof(Rec<double>);
    int rc = ensure_pinned(c, (size_t) std::max<int64_t>(nRet, 1) * recSz); if (rc) return rc;
    if (W > 1) { rc = dist_reserve(c->stream, (size_t) K * recSz); if (rc) return rc; }
    if (nLocal > 0) {
        prof_begin(c, CLS_SELECT, K_SELECT);
        void *dst = W > 1 ? (void *) g_dist.send : c->h_out;
        if (c->ps == 4) launch_topk<float, 4>(c, c->d_keys, n, (int) nLocal, (Rec<float> *) dst, list, W, r);
        else launch_topk<double, 8>(c, c->d_keys, n, (int) nLocal, (Rec<double> *) dst, list, W, r);
        prof_end(c);
        CK(cudaGetLastError());
    }
    if (W > 1 && nRet > 0) {
        // the exchange step of SURVEY 8e: W x K fixed-size records, then every rank merges them to the same K best
        const char *base; size_t stride;
        rc = dist_allgather(c->stream, (size_t) K * recSz, &base, &stride); if (rc) return rc;
        prof_begin(c, CLS_SELECT, K_SELECT);
        const unsigned blocks = (unsigned) (((int64_t) W * K + 255) / 256);
        if (c->ps == 4) k_rank_merge<float><<<blocks, 256, 0, c->stream>>>(base, stride, W, c->nAct, (int) K, (Rec<float> *) c->h_out);
        else k_rank_merge<double><<<blocks, 256, 0, c->stream>>>(base, stride, W, c->nAct, (int) K, (Rec<double> *) c->h_out);
        prof_end(c);
        CK(cudaGetLastError());
        c->cnt.launches += g_dist.mode == DIST_PEER ? 2 : 1;
    }
    c->cnt.d2hBytes += (int64_t) (nRet * recSz);
    CK(sync_stream(c));
    if (c->ps == 4) {
        const Rec<float> *rr = (const Rec<float> *) c->h_out;
        for (int64_t k = 0; k < nRet; k++) { j_out[k] = rr[k].j; ((float *) dist)[k] = rr[k].dist; ((float *) weight)[k] = rr[k].weight; ((float *) criterion)[k] = rr[k].crit; }
    } else {
        const Rec<double> *rr = (const Rec<double> *) c->h_out;
        for (int64_t k = 0; k < nRet; k++) { j_out[k] = rr[k].j; ((double *) dist)[k] = rr[k].dist; ((double *) weight)[k] = rr[k].weight; ((double *) criterion)[k] = rr[k].crit; }
    }
    *nOut = nRet;
    // accounting: one distance per active node (a sharded context accounts its share)
    const int64_t nLeaf = W > 1 ? c->nActLeaf / W : c->nActLeaf, nInt = W > 1 ? c->nActInternal / W : c->nActInternal;
    if (query < c->N) { c->cnt.seqOps += nLeaf; c->cnt.algoBytes += nLeaf * c->L; }
    else { c->cnt.profileOps += nLeaf; c->cnt.algoBytes += nLeaf * profile_bytes(c, 0); }
    c->cnt.profileOps += nInt;
    c->cnt.algoBytes += nInt * profile_bytes(c, c->N);
    c->cnt.algoBytes += profile_bytes(c, query);
    return VFT_OK;
}

extern "C" int vft_tophits_merge(vft_ctx *c, int64_t newnode, int64_t nActive, int64_t m, int64_t nLists, const int64_t *iNode,
                                 const int64_t *ownOffset, const int64_t *ownJ, const void *ownDist, int64_t nAvail,
                                 const int64_t *allJ, const void *allDist, int64_t *outCount, int64_t *outJ, void *outDist) {
    if (!c || !iNode || !ownOffset || !allJ || !allDist || !outCount || !outJ || !outDist || nLists < 0 || m < 1 || nAvail < 0 || nActive < 3)
        return fail(VFT_EINVAL, "bad argument");
    bind_device(c);
    if (nLists == 0) return VFT_OK;
    BytesScope bytesScope(c, K_EVAL_LARGE);
    int64_t maxOwn = 0;
    for (int64_t l = 0; l < nLists; l++) {
        if (iNode[l] < 0 || iNode[l] >= c->maxnode || !c->activeHost[iNode[l]] || ownOffset[l + 1] < ownOffset[l]) return fail(VFT_EINVAL, "bad list");
        maxOwn = std::max(maxOwn, ownOffset[l + 1] - ownOffset[l]);
    }
    const int64_t total = ownOffset[nLists], cap = (maxOwn + nAvail + 31) / 32 * 32;      // slot stride of a list: whole warps (k_sweep20)
    if (total > 0 && (!ownJ || !ownDist)) return fail(VFT_EINVAL, "null argument");
    int np2 = 32;
    while (np2 < cap) np2 <<= 1;
    if (np2 > MRG_MAX) return fail(VFT_EINVAL, "candidate lists longer than 4096 entries are not supported");
    const size_t ps = c->ps;
    // request: int32 iNode[nLists] | ownOffset[nLists+1] | ownJ[total] | allJ[nAvail] | (8-aligned) P ownDist[total] | allDist[nAvail]
    const size_t nInts = (size_t) nLists + (size_t) nLists + 1 + (size_t) total + (size_t) nAvail;
    const size_t offP = (nInts * 4 + 7) & ~(size_t) 7;
    const size_t inBytes = offP + ((size_t) total + (size_t) nAvail) * ps;
    // response: int32 count[nLists] | j[nLists*m] | (8-aligned) P dist[nLists*m]
    const size_t offOD = (((size_t) nLists + (size_t) nLists * m) * 4 + 7) & ~(size_t) 7;
    const size_t outBytes = offOD + (size_t) nLists * m * ps + 32;
    // a sharded context takes the lists [l0, l1) of W contiguous chunks; its saved lists travel in the exchange buffer
    const int W = shard_world(c, nLists * cap), rk = shard_rank(c, W);
    const int64_t chunkLists = (nLists + W - 1) / W;
    const int64_t l0 = std::min(nLists, rk * chunkLists), l1 = std::min(nLists, l0 + chunkLists), myLists = l1 - l0;
    const size_t xOffD = (((size_t) chunkLists + (size_t) chunkLists * m) * 4 + 7) & ~(size_t) 7;      // chunk image: count | j | (8-aligned) dist
    const size_t xBytes = (xOffD + (size_t) chunkLists * m * ps + 15) & ~(size_t) 15;
    int rc = ensure_pinned(c, std::max(std::max(inBytes, outBytes), W > 1 ? (size_t) W * xBytes : (size_t) 0)); if (rc) return rc;
    if (W > 1) { rc = dist_reserve(c->stream, xBytes); if (rc) return rc; }
    int32_t *hi = (int32_t *) c->h_in;
    int32_t *hNode = hi, *hOff = hNode + nLists, *hOwnJ = hOff + nLists + 1, *hAllJ = hOwnJ + total;
    for (int64_t l = 0; l < nLists; l++) hNode[l] = (int32_t) iNode[l];
    for (int64_t l = 0; l <= nLists; l++) hOff[l] = (int32_t) ownOffset[l];
    for (int64_t k = 0; k < total; k++) {
        if (ownJ[k] >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
        hOwnJ[k] = ownJ[k] < 0 ? -1 : (int32_t) ownJ[k];
    }
    for (int64_t k = 0; k < nAvail; k++) {
        if (allJ[k] < 0 || allJ[k] >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
        hAllJ[k] = (int32_t) allJ[k];
    }
    char *hP = (char *) c->h_in + offP;
    std::memcpy(hP, ownDist ? ownDist : allDist, (size_t) total * ps);
    std::memcpy(hP + (size_t) total * ps, allDist, (size_t) nAvail * ps);
    // device scratch per slot: uJ, reqA, reqB (int32), uD, r0, r1 (P); + count[nLists]
    const size_t slots = (size_t) nLists * cap;
    const size_t need = slots * (12 + 3 * ps) + (size_t) nLists * 4 + 64;
    if (need > c->mrgCap) {
        mem_free(c->d_mrg);
        c->d_mrg = nullptr; c->mrgCap = 0;
        CK(mem_alloc((void **) &c->d_mrg, need * 2, MEM_DEVICE));
        c->mrgCap = need * 2;
    }
    if (!c->d_acct) { CK(mem_alloc((void **) &c->d_acct, 32, MEM_DEVICE)); }
    CK(cudaMemsetAsync(c->d_acct, 0, 32, c->stream));
    char *dm = (char *) c->d_mrg;
    void *uD = dm, *r0 = dm + slots * ps, *r1 = dm + 2 * slots * ps;                      // P arrays first (8-byte aligned)
    int32_t *uJ = (int32_t *) (dm + 3 * slots * ps), *reqA = uJ + slots, *reqB = reqA + slots, *cnt = reqB + slots;
    const size_t smemSort = (size_t) np2 * 12;
    // results: straight into mapped host memory, or (sharded) into this rank's chunk image of the exchange buffer
    int32_t *hoCount = (int32_t *) c->h_out, *hoJ = hoCount + nLists;
    void *hoD = (char *) c->h_out + offOD;
    if (W > 1) { hoCount = (int32_t *) g_dist.send; hoJ = hoCount + chunkLists; hoD = g_dist.send + xOffD; }
    const size_t mySlots = (size_t) myLists * cap, so = (size_t) l0 * cap;                 // this rank's slots start at `so`
    const int G = pick_group(c, (int64_t) mySlots);
    const int64_t warps = ((int64_t) mySlots + G - 1) / G;
    const unsigned evalBlocks = (unsigned) ((warps + 3) / 4);
    InlineItems inl;
    inl.a[0] = 0;
    // (kernels index lists from 0: the request arrays are passed at list l0, the absolute ownOffset values still address
    //  the whole ownJ / ownDist arrays)
#define CALL_MERGE(P, A_, MX)                                                                                     \
    do {                                                                                                          \
        cudaFuncSetAttribute(k_merge_prep<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, MRG_MAX * 12);         \
        cudaFuncSetAttribute(k_merge_finish<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, MRG_MAX * 12);       \
        prof_begin(c, CLS_SELECT, K_MERGE);                                                                                \
        k_merge_prep<P><<<(unsigned) myLists, MRG_T, smemSort, c->stream>>>(hNode + l0, hOff + l0, hOwnJ, (const P *) hP, (int) nAvail, hAllJ, \
            (const P *) (hP + (size_t) total * ps), (int) newnode, (int) cap, np2, (int) c->N, uJ + so, (P *) uD + so, reqA + so, reqB + so, cnt + l0, c->d_acct); \
        prof_end(c);                                                                                              \
        prof_begin(c, CLS_DIST, K_EVAL_LARGE);                                                                                  \
        if (c->sweepOk && (c->sweepModes & 4)) {                                                                  \
            k_query_tables<P><<<dim3((unsigned) ((c->Lp + 127) / 128), (unsigned) myLists), 128, 0, c->stream>>>(make_store<P>(c), hNode + l0, -1, (P *) c->d_qcd, (P *) c->d_qv, (P *) c->d_qw); \
            LAUNCH_SWEEP(P, 2, (int64_t) mySlots, make_store<P>(c), make_qtab<P>(c, false), nullptr, 1, 0, reqA + so, reqB + so, (int) cap, (int64_t) mySlots, -1, nActive, 0.0, (P *) r0 + so, (P *) r1 + so, (P *) nullptr, (uint64_t *) nullptr, (P *) nullptr); \
        } else                                                                                                    \
        k_eval<P, A_, MX, true><<<evalBlocks, 128, 4 * group_smem_bytes<P, A_, MX>(G), c->stream>>>(make_store<P>(c), inl, reqA + so, reqB + so, (int64_t) mySlots, 0, G, 0, nActive, 0.0, (P *) r0 + so, (P *) r1 + so, c->d_doneCount, (P *) nullptr); \
        prof_end(c);                                                                                              \
        prof_begin(c, CLS_SELECT, K_MERGE);                                                                                \
        k_merge_finish<P><<<(unsigned) myLists, MRG_T, smemSort, c->stream>>>(make_store<P>(c), hNode + l0, nActive, (int) m, (int) cap, np2, uJ + so, (P *) uD + so, reqA + so, (const P *) r0 + so, cnt + l0, hoCount, hoJ, (P *) hoD); \
        prof_end(c);                                                                                              \
    } while (0)
    if (myLists > 0 && c->sweepOk && (c->sweepModes & 4)) { rc = ensure_qtabs(c, myLists); if (rc) return rc; }
    if (myLists > 0) { VFT_DISPATCH(c, CALL_MERGE); }
    CK(cudaGetLastError());
    unsigned long long acct[4];
    CK(cudaMemcpyAsync(acct, c->d_acct, 32, cudaMemcpyDeviceToHost, c->stream));
    if (W > 1) {
        const char *base; size_t stride;
        rc = dist_allgather(c->stream, xBytes, &base, &stride); if (rc) return rc;
        for (int w = 0; w < W; w++) CK(cudaMemcpyAsync((char *) c->h_out + (size_t) w * xBytes, base + (size_t) w * stride, xBytes, cudaMemcpyDeviceToHost, c->stream));
        c->cnt.launches += g_dist.mode == DIST_PEER ? 1 : 0;
    }
    CK(sync_stream(c));
    c->cnt.launches += 3;
    for (int64_t l = 0; l < nLists; l++) {
        const int32_t *count = (const int32_t *) c->h_out, *jj = count + nLists;
        const char *dd = (const char *) c->h_out + offOD;
        int64_t li = l;
        if (W > 1) {                                      // list l sits in the chunk image of rank l / chunkLists
            const char *img = (const char *) c->h_out + (size_t) (l / chunkLists) * xBytes;
            count = (const int32_t *) img; jj = count + chunkLists; dd = img + xOffD; li = l % chunkLists;
        }
        outCount[l] = count[li];
        for (int64_t k = 0; k < count[li]; k++) outJ[l * m + k] = jj[li * m + k];
        std::memcpy((char *) outDist + (size_t) l * m * ps, dd + (size_t) li * m * ps, (size_t) count[li] * ps);
        if (l >= l0 && l < l1) c->cnt.algoBytes += profile_bytes(c, iNode[l]);           // a list shares its query
    }
    c->cnt.seqOps += (int64_t) acct[0]; c->cnt.profileOps += (int64_t) acct[1];
    c->cnt.algoBytes += (int64_t) acct[2] * c->L + (int64_t) acct[3] * profile_bytes(c, c->N);
    c->cnt.h2dBytes += (int64_t) inBytes; c->cnt.d2hBytes += (int64_t) (W > 1 ? (size_t) W * xBytes : outBytes);
    return VFT_OK;
}

template<typename P>
static MLModel<P> make_model(vft_ctx *c) {
    MLModel<P> m;
    const P *t = (const P *) c->mlTables;
    const int A = c->A;
    m.codeFreq = c->hasTransmat ? t : nullptr;
    m.eigenval = t + (A + 1) * A; m.eigeninv = m.eigenval + A; m.eigeninvT = m.eigeninv + A * A; m.statinv = m.eigeninvT + A * A;
    m.rates = (const P *) c->mlRates; m.ratecat = c->mlRatecat;
    m.nRateCats = c->nRateCats; m.fastexp = c->fastexp;
    m.MLMinRelBranchLength = c->MLMinRel; m.MLMinBranchLength = c->MLMinBr;
    return m;
}

extern "C" int vft_upload_transmat(vft_ctx *c, const void *codeFreq, const void *eigenval, const void *eigeninv,
                                   const void *eigeninvT, const void *statinv) {
    if (!c) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    c->hasTransmat = codeFreq != nullptr;
    if (!codeFreq) return VFT_OK;
    if (!eigenval || !eigeninv || !statinv) return fail(VFT_EINVAL, "null argument");
    const size_t ps = c->ps, A = (size_t) c->A;
    std::vector<char> h(1300 * ps, 0);
    char *q = h.data();
    std::memcpy(q, codeFreq, (A + 1) * A * ps); q += (A + 1) * A * ps;
    std::memcpy(q, eigenval, A * ps); q += A * ps;
    std::memcpy(q, eigeninv, A * A * ps); q += A * A * ps;
    if (eigeninvT) std::memcpy(q, eigeninvT, A * A * ps);
    q += A * A * ps;
    std::memcpy(q, statinv, A * ps);
    CK(cudaMemcpyAsync(c->mlTables, h.data(), h.size(), cudaMemcpyHostToDevice, c->stream));
    CK(sync_stream(c));
    return VFT_OK;
}

extern "C" int vft_sync_rates(vft_ctx *c, const void *rates, int64_t nRateCats, const int64_t *ratecat,
                              double MLMinRelBranchLength, double MLMinBranchLength, int32_t fastexpLevel) {
    if (!c || !rates || !ratecat || nRateCats < 1 || nRateCats > 64 || fastexpLevel < 0 || fastexpLevel > 3)
        return fail(VFT_EINVAL, "bad argument");
    bind_device(c);
    std::vector<int32_t> rc((size_t) c->Lp, 0);
    for (int64_t i = 0; i < c->L; i++) {
        if (ratecat[i] < 0 || ratecat[i] >= nRateCats) return fail(VFT_EINVAL, "bad rate category");
        rc[(size_t) i] = (int32_t) ratecat[i];
    }
    CK(cudaMemcpyAsync(c->mlRates, rates, (size_t) nRateCats * c->ps, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->mlRatecat, rc.data(), rc.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(sync_stream(c));
    c->nRateCats = (int) nRateCats; c->fastexp = fastexpLevel; c->MLMinRel = MLMinRelBranchLength; c->MLMinBr = MLMinBranchLength;
    c->hasRates = true;
    return VFT_OK;
}

// ids the likelihood entry points may read / write: tree nodes, plus the scratch rows of cfg.nScratch
static inline bool ml_readable(const vft_ctx *c, int64_t id) { return id >= 0 && (id < c->maxnode || (id >= c->M && id < c->M + c->S)); }
static inline bool ml_writable(const vft_ctx *c, int64_t id) { return id >= c->N && id < c->M + c->S; }
static inline void ml_written(vft_ctx *c, int64_t id) {
    if (id >= c->M) return;
    if (!c->activeHost[id]) { c->activeHost[id] = 1; c->nActInternal++; c->actDirty = true; }
    if (id >= c->maxnode) c->maxnode = id + 1;
}

extern "C" int vft_pair_loglk_batch(vft_ctx *c, const int64_t *pi, const int64_t *pj, const double *length, int64_t n,
                                    double *loglk, double *siteLk) {
    if (!c || (n > 0 && (!pi || !pj || !length || !loglk))) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    if (!c->hasRates) return fail(VFT_EINVAL, "vft_sync_rates has not been called");
    if (!c->hasTransmat && c->A != 4) return fail(VFT_EINVAL, "Jukes-Cantor needs nCodes == 4");
    if (n == 0) return VFT_OK;
    BytesScope bytesScope(c, K_LOGLK);
    const size_t need = (size_t) n * 24 + (siteLk ? (size_t) n * c->L * 8 : 0);
    int rc = ensure_pinned(c, need); if (rc) return rc;
    int32_t *ha = (int32_t *) c->h_in, *hb = ha + n;
    double *hl = (double *) ((char *) c->h_in + (size_t) n * 8);
    for (int64_t k = 0; k < n; k++) {
        if (!ml_readable(c, pi[k]) || !ml_readable(c, pj[k])) return fail(VFT_EINVAL, "bad node id");
        ha[k] = (int32_t) pi[k]; hb[k] = (int32_t) pj[k]; hl[k] = length[k];
        c->cnt.algoBytes += profile_bytes(c, pi[k]) + profile_bytes(c, pj[k]) + c->L * 4;
    }
    double *outLk = (double *) c->h_out, *outSite = siteLk ? outLk + n : nullptr;
    // warps per item: as many as keep ~16 warps per SM busy, at most 8 (one CTA per item)
    int W = 1;
    while (W < 8 && n * (2 * W) <= 148 * 16) W *= 2;
    const int tableBytes = (int) (((c->hasTransmat ? (size_t) c->nRateCats * c->A * c->ps : (size_t) c->nRateCats * 16) + 15) / 16 * 16);
    const size_t perItem = (size_t) c->Lp * 8 + (size_t) tableBytes;
    while (W < 8 && (size_t) (8 / W) * perItem > 100 * 1024) W *= 2;      // long alignments: fewer items per CTA, two CTAs per SM
    const int perCta = 8 / W;
    const size_t smem = (size_t) perCta * perItem;
    if (smem > 200 * 1024) return fail(VFT_EINVAL, "alignment too long for the per-item likelihood buffers");
    const unsigned blocks = (unsigned) ((n + perCta - 1) / perCta);
#define CALL_LK(P, A_)                                                                                          \
    do {                                                                                                        \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_pair_loglk<P, A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
        k_pair_loglk<P, A_><<<blocks, 256, smem, c->stream>>>(make_store<P>(c), make_model<P>(c), ha, hb, hl, n, outLk, outSite, W, tableBytes); \
    } while (0)
    prof_begin(c, CLS_DIST, K_LOGLK);
    if (c->cfg.precision == 32) { if (c->A == 4) CALL_LK(float, 4); else CALL_LK(float, 20); }
    else { if (c->A == 4) CALL_LK(double, 4); else CALL_LK(double, 20); }
    prof_end(c);
    CK(cudaGetLastError());
    CK(sync_stream(c));
    c->cnt.launches++;
    std::memcpy(loglk, outLk, (size_t) n * 8);
    if (siteLk) std::memcpy(siteLk, outSite, (size_t) n * c->L * 8);
    return VFT_OK;
}

extern "C" int vft_posterior_profile(vft_ctx *c, int64_t out_id, int64_t id1, int64_t id2, double len1, double len2) {
    if (!c || !ml_writable(c, out_id) || !ml_readable(c, id1) || !ml_readable(c, id2))
        return fail(VFT_EINVAL, "bad node id");
    bind_device(c);
    if (!c->hasRates) return fail(VFT_EINVAL, "vft_sync_rates has not been called");
    if (!c->hasTransmat && c->A != 4) return fail(VFT_EINVAL, "Jukes-Cantor needs nCodes == 4");
    const unsigned blocks = (unsigned) ((c->Lp + 127) / 128);
#define CALL_POST(P, A_) k_posterior<P, A_><<<blocks, 128, 0, c->stream>>>(make_store<P>(c), make_model<P>(c), out_id, id1, id2, len1, len2, nullptr, nullptr)
    prof_begin(c, CLS_PROFILE, K_POSTERIOR);
    if (c->cfg.precision == 32) { if (c->A == 4) CALL_POST(float, 4); else CALL_POST(float, 20); }
    else { if (c->A == 4) CALL_POST(double, 4); else CALL_POST(double, 20); }
    prof_end(c);
    CK(cudaGetLastError());
    c->cnt.launches++;
    ml_written(c, out_id);
    return VFT_OK;            // asynchronous
}

extern "C" int vft_posterior_profile_batch(vft_ctx *c, int64_t n, const int64_t *out_id, const int64_t *id1, const int64_t *id2,
                                           const double *len1, const double *len2) {
    if (!c || n < 0 || (n > 0 && (!out_id || !id1 || !id2 || !len1 || !len2))) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    if (!c->hasRates) return fail(VFT_EINVAL, "vft_sync_rates has not been called");
    if (!c->hasTransmat && c->A != 4) return fail(VFT_EINVAL, "Jukes-Cantor needs nCodes == 4");
    const int64_t CH = 32768;                                  // grid.y limit is 65535
    for (int64_t k0 = 0; k0 < n; k0 += CH) {
        const int64_t m = std::min(CH, n - k0);
        int rc = ensure_pinned(c, (size_t) m * 32); if (rc) return rc;
        rc = ensure_lists(c, m * 4); if (rc) return rc;
        int32_t *hi = (int32_t *) c->h_in;
        double *hl = (double *) ((char *) c->h_in + (size_t) m * 16);
        for (int64_t k = 0; k < m; k++) {
            const int64_t o = out_id[k0 + k], a = id1[k0 + k], b = id2[k0 + k];
            if (!ml_writable(c, o) || a < 0 || b < 0 || a >= c->M + c->S || b >= c->M + c->S) return fail(VFT_EINVAL, "bad node id");
            hi[3 * k] = (int32_t) o; hi[3 * k + 1] = (int32_t) a; hi[3 * k + 2] = (int32_t) b;
            hl[2 * k] = len1[k0 + k]; hl[2 * k + 1] = len2[k0 + k];
        }
        // one DMA for the items, then one launch for the whole level
        CK(cudaMemcpyAsync(c->d_pi, c->h_in, (size_t) m * 32, cudaMemcpyHostToDevice, c->stream));
        const int32_t *dItems = (const int32_t *) c->d_pi;
        const double *dLens = (const double *) ((const char *) c->d_pi + (size_t) m * 16);
        const dim3 grid((unsigned) ((c->Lp + 127) / 128), (unsigned) m);
#define CALL_POSTB(P, A_) k_posterior<P, A_><<<grid, 128, 0, c->stream>>>(make_store<P>(c), make_model<P>(c), 0, 0, 0, 0.0, 0.0, dItems, dLens)
        prof_begin(c, CLS_PROFILE, K_POSTERIOR);
        if (c->cfg.precision == 32) { if (c->A == 4) CALL_POSTB(float, 4); else CALL_POSTB(float, 20); }
        else { if (c->A == 4) CALL_POSTB(double, 4); else CALL_POSTB(double, 20); }
        prof_end(c);
        CK(cudaGetLastError());
        CK(sync_stream(c));                                   // h_in is reused
        c->cnt.launches++;
        for (int64_t k = 0; k < m; k++) ml_written(c, out_id[k0 + k]);
    }
    return VFT_OK;
}

extern "C" int vft_get_config(vft_ctx *c, vft_config *out, int32_t *hasTransmat) {
    if (!c || !out) return fail(VFT_EINVAL, "null argument");
    *out = c->cfg;
    if (hasTransmat) *hasTransmat = c->hasTransmat ? 1 : 0;
    return VFT_OK;
}

extern "C" int vft_get_profile(vft_ctx *c, int64_t id, void *weights, uint8_t *codes, void *vectors) {
    if (!c || id < -1 || !(id == -1 || ml_readable(c, id))) return fail(VFT_EINVAL, "bad node id");
    bind_device(c);
    const size_t ps = c->ps, L = (size_t) c->L, Lp = (size_t) c->Lp, A = (size_t) c->A;
    CK(sync_stream(c));
    if (id < 0) {
        if (weights) CK(cudaMemcpy(weights, c->ow, L * ps, cudaMemcpyDeviceToHost));
        if (codes) std::memset(codes, VFT_NOCODE, L);
        if (vectors) CK(cudaMemcpy(vectors, c->ov, L * A * ps, cudaMemcpyDeviceToHost));
        return VFT_OK;
    }
    std::vector<uint8_t> cd(L);
    CK(cudaMemcpy(cd.data(), (char *) c->codes + (size_t) id * Lp, L, cudaMemcpyDeviceToHost));
    if (codes) std::memcpy(codes, cd.data(), L);
    if (id < c->N) {
        if (weights) for (size_t p = 0; p < L; p++) {
            if (ps == 4) ((float *) weights)[p] = cd[p] != VFT_NOCODE; else ((double *) weights)[p] = cd[p] != VFT_NOCODE;
        }
        if (vectors) std::memset(vectors, 0, L * A * ps);
    } else {
        const size_t row = (size_t) (id - c->N);
        if (weights) CK(cudaMemcpy(weights, (char *) c->weights + row * Lp * ps, L * ps, cudaMemcpyDeviceToHost));
        if (vectors) CK(cudaMemcpy(vectors, (char *) c->vecs + row * Lp * A * ps, L * A * ps, cudaMemcpyDeviceToHost));
    }
    return VFT_OK;
}

// a dense profile written from the host into an internal-node or scratch row (the counterpart of vft_get_profile)
extern "C" int vft_put_profile(vft_ctx *c, int64_t id, const void *weights, const uint8_t *codes, const void *vectors) {
    if (!c || !weights || !codes || !vectors) return fail(VFT_EINVAL, "null argument");
    if (!ml_writable(c, id)) return fail(VFT_EINVAL, "bad node id");
    bind_device(c);
    const size_t ps = c->ps, L = (size_t) c->L, Lp = (size_t) c->Lp, A = (size_t) c->A, row = (size_t) (id - c->N);
    CK(sync_stream(c));
    CK(cudaMemcpy((char *) c->codes + (size_t) id * Lp, codes, L, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((char *) c->weights + row * Lp * ps, weights, L * ps, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((char *) c->vecs + row * Lp * A * ps, vectors, L * A * ps, cudaMemcpyHostToDevice));
    ml_written(c, id);
    return VFT_OK;
}

extern "C" int vft_sh_support_batch(vft_ctx *c, int64_t n, int64_t nBootstrap, const int64_t *col, const double *loglk,
                                    const double *siteLk, double *support) {
    if (!c || n < 0 || nBootstrap < 1 || (n > 0 && (!col || !loglk || !siteLk || !support))) return fail(VFT_EINVAL, "bad argument");
    if (n == 0) return VFT_OK;
    bind_device(c);
    const int64_t L = c->L;
    const size_t smem = (size_t) 3 * L * 8;
    if (smem > 200 * 1024) return fail(VFT_EINVAL, "alignment too long for the per-quartet site table");
    // the resampled columns, transposed to [column][resample] and narrowed; the logarithms with the host's libm (NJ.tcc:1134-1136)
    std::vector<int32_t> colT((size_t) (L * nBootstrap));
    for (int64_t b = 0; b < nBootstrap; b++)
        for (int64_t j = 0; j < L; j++) {
            const int64_t p = col[b * L + j];
            if (p < 0 || p >= L) return fail(VFT_EINVAL, "bad column index");
            colT[(size_t) (j * nBootstrap + b)] = (int32_t) p;
        }
    const int64_t CH = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n, 65535), (int64_t) (256 << 20) / (24 * L)));     // quartets per launch (grid.y)
    void *dCol = nullptr, *dSl = nullptr, *dLk = nullptr, *dVotes = nullptr;
    CK(mem_alloc(&dCol, colT.size() * 4, MEM_DEVICE)); CK(mem_alloc(&dSl, (size_t) CH * 3 * L * 8, MEM_DEVICE));
    CK(mem_alloc(&dLk, (size_t) CH * 24, MEM_DEVICE)); CK(mem_alloc(&dVotes, (size_t) CH * 4, MEM_DEVICE));
    CK(cudaMemcpyAsync(dCol, colT.data(), colT.size() * 4, cudaMemcpyHostToDevice, c->stream));
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_sh_support, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    std::vector<double> sl((size_t) CH * 3 * L);
    std::vector<unsigned int> votes((size_t) CH);
    int rc = VFT_OK;
    for (int64_t q0 = 0; q0 < n && rc == VFT_OK; q0 += CH) {
        const int64_t m = std::min(CH, n - q0);
        for (int64_t k = 0; k < m * 3 * L; k++) sl[(size_t) k] = std::log(siteLk[q0 * 3 * L + k]);
        cudaError_t e = cudaMemcpyAsync(dSl, sl.data(), (size_t) m * 3 * L * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dLk, loglk + 3 * q0, (size_t) m * 24, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(dVotes, 0, (size_t) m * 4, c->stream);
        if (e == cudaSuccess) {
            const dim3 grid((unsigned) ((nBootstrap + 255) / 256), (unsigned) m);
            k_sh_support<<<grid, 256, smem, c->stream>>>((const double *) dSl, (const double *) dLk, (const int32_t *) dCol, L, nBootstrap, (unsigned int *) dVotes);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(votes.data(), dVotes, (size_t) m * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = sync_stream(c);
        if (e != cudaSuccess) { std::snprintf(g_err, sizeof g_err, "CUDA error: %s", cudaGetErrorString(e)); rc = VFT_ECUDA; break; }
        c->cnt.launches++;
        for (int64_t k = 0; k < m; k++) support[q0 + k] = (double) votes[(size_t) k] / (double) nBootstrap;         // :1164
    }
    mem_free(dCol); mem_free(dSl); mem_free(dLk); mem_free(dVotes);
    return rc;
}

extern "C" int vft_get_counters(vft_ctx *c, vft_counters *out) {
    if (!c || !out) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    CK(sync_stream(c));
    c->cnt.distBytes = c->cnt.algoBytes;      // only the distance kernels account algorithmic bytes
    *out = c->cnt;
    return VFT_OK;
}

extern "C" int vft_timer_start(vft_ctx *c) {
    if (!c) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    CK(cudaEventRecord(c->tmr0, c->stream));
    return VFT_OK;
}

extern "C" int vft_timer_stop(vft_ctx *c, double *ms) {
    if (!c || !ms) return fail(VFT_EINVAL, "null argument");
    bind_device(c);
    CK(cudaEventRecord(c->tmr1, c->stream));
    CK(cudaEventSynchronize(c->tmr1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->tmr0, c->tmr1));
    *ms = f;
    return VFT_OK;
}

#include "vft_ingest.cuh"
#include "nj_loop_gpu.cuh"
