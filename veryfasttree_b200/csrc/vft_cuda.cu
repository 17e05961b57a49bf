// vft_cuda.cu -- the CUDA (sm_100a) implementation of the C-ABI in include/vft_b200.h.
//
// Data layout in HBM (one context = one GPU):
//   codes    uint8 [2N][Lp]        every node; leaves are nothing else (1 B/position)
//   weights  P     [N][Lp]         internal nodes only (row = id - N)
//   vecs     P     [N][Lp][A]      internal nodes only, dense (unused where the code is known)
//   out-profile ow[Lp], ov[Lp][A], ocd[Lp][A]; per-node scalars diameter/selfdist/selfweight/outDist/active
// Lp = nPos rounded up to 16 so that code rows are read with 128-bit loads.
//
// Kernels (one thread accumulates one pair, see vft_device.cuh for why):
//   k_dist_pairs      candidate lists           (transferBestHits / uniqueBestHits / getBestFromTopHits)
//   k_one_vs_all      query vs every active node + criterion -> 64-bit sort keys   (setBestHit)
//   k_topk_*          chunked bitonic sort + merge tree: the K best in the reference's psort order
//   k_out_distance    profileDist(node, out-profile) + the setOutDistance algebra
//   k_average         averageProfile + self distance of the new node
//   k_outprofile_*    updateOutProfile / outProfile + setCodeDist
// There is no CPU fallback: without a usable device vft_ctx_create returns VFT_ENODEVICE.
#include "../../include/vft_b200.h"
#include "vft_device.cuh"

#include <algorithm>
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

// =================================================================================================
// kernels
// =================================================================================================

template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128)
k_dist_pairs(Store<P> s, const int64_t *__restrict__ pi, const int64_t *__restrict__ pj, int64_t n, int raw,
             P *__restrict__ dist, P *__restrict__ weight) {
    const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (t >= n) return;
    P d, w;
    join_dist<P, A, MATRIX>(s, pi[t], pj[t], raw != 0, d, w);
    dist[t] = d;
    weight[t] = w;
}

// setBestHit (NJ.tcc:3571-3639): one thread per node slot j < maxnode
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128)
k_one_vs_all(Store<P> s, int64_t query, int64_t nActive, int64_t maxnode, P *__restrict__ dist,
             P *__restrict__ weight, P *__restrict__ crit, uint64_t *__restrict__ keys) {
    const int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (j >= maxnode) return;
    if (!s.active[j]) { keys[j] = ~0ull; return; }
    P d, w;
    join_dist<P, A, MATRIX>(s, query, j, false, d, w);
    // setCriterion (NJ.tcc:1099-1107) with every out-distance fresh at this nActive
    const double outI = (double) s.outDist[query], outJ = (double) s.outDist[j];
    const P c = (P) xsub((double) d, xadd(outI, outJ) / (double) (nActive - 2));
    dist[j] = d; weight[j] = w; crit[j] = c;
    keys[j] = order_key(c);
}

// ---- top-K in psort order: key ascending, ties by index DESCENDING -----------------------------
struct KV { uint64_t key; uint32_t idx; };
__device__ __forceinline__ bool kv_less(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) {
    return ka < kb || (ka == kb && ia > ib);
}

constexpr int SORT_N = 4096;        // elements sorted per CTA
constexpr int SORT_T = 1024;

__device__ void bitonic_sort_smem(uint64_t *k, uint32_t *v) {
    for (int size = 2; size <= SORT_N; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < SORT_N / 2; t += SORT_T) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint64_t ka = k[lo], kb = k[hi];
                const uint32_t ia = v[lo], ib = v[hi];
                const bool swap = up ? kv_less(kb, ib, ka, ia) : kv_less(ka, ia, kb, ib);
                if (swap) { k[lo] = kb; k[hi] = ka; v[lo] = ib; v[hi] = ia; }
            }
        }
    }
    __syncthreads();
}

// stage 1: CTA b sorts slots [b*SORT_N, (b+1)*SORT_N) and keeps its Kc best
__global__ void __launch_bounds__(SORT_T)
k_topk_chunks(const uint64_t *__restrict__ keys, int64_t n, int Kc, uint64_t *__restrict__ outK, uint32_t *__restrict__ outV) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *k = reinterpret_cast<uint64_t *>(smem);
    uint32_t *v = reinterpret_cast<uint32_t *>(smem + sizeof(uint64_t) * SORT_N);
    const int64_t base = (int64_t) blockIdx.x * SORT_N;
    for (int t = threadIdx.x; t < SORT_N; t += SORT_T) {
        const int64_t j = base + t;
        k[t] = j < n ? keys[j] : ~0ull;
        v[t] = (uint32_t) (j < n ? j : 0xFFFFFFFFu - t);      // padding: distinct indices, sorts after all
    }
    bitonic_sort_smem(k, v);
    for (int t = threadIdx.x; t < Kc; t += SORT_T) { outK[(int64_t) blockIdx.x * Kc + t] = k[t]; outV[(int64_t) blockIdx.x * Kc + t] = v[t]; }
}

// stage 2..: CTA b merges `R` sorted lists of Kc entries (R*Kc <= SORT_N) and keeps the Kc best
__global__ void __launch_bounds__(SORT_T)
k_topk_merge(const uint64_t *__restrict__ inK, const uint32_t *__restrict__ inV, int nLists, int R, int Kc,
             uint64_t *__restrict__ outK, uint32_t *__restrict__ outV) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *k = reinterpret_cast<uint64_t *>(smem);
    uint32_t *v = reinterpret_cast<uint32_t *>(smem + sizeof(uint64_t) * SORT_N);
    const int first = blockIdx.x * R;
    const int have = min(R, nLists - first) * Kc;
    for (int t = threadIdx.x; t < SORT_N; t += SORT_T) {
        if (t < have) { k[t] = inK[(int64_t) first * Kc + t]; v[t] = inV[(int64_t) first * Kc + t]; }
        else { k[t] = ~0ull; v[t] = 0u; }
    }
    bitonic_sort_smem(k, v);
    for (int t = threadIdx.x; t < Kc; t += SORT_T) { outK[(int64_t) blockIdx.x * Kc + t] = k[t]; outV[(int64_t) blockIdx.x * Kc + t] = v[t]; }
}

template<typename P>
struct Rec { int64_t j; P dist, weight, crit; };

template<typename P>
__global__ void k_gather_topk(const uint32_t *__restrict__ idx, int K, const P *__restrict__ dist,
                              const P *__restrict__ weight, const P *__restrict__ crit, Rec<P> *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K) return;
    const uint32_t j = idx[t];
    out[t].j = j; out[t].dist = dist[j]; out[t].weight = weight[j]; out[t].crit = crit[j];
}

// setOutDistance for a list of nodes (ids != nullptr) or for every active node (ids == nullptr)
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(128)
k_out_distance(Store<P> s, const int64_t *__restrict__ ids, int64_t n, int64_t nActive, double totdiam,
               P *__restrict__ out, int commit) {
    const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t id = ids ? ids[t] : t;
    if (!ids && !s.active[id]) return;
    const P v = out_distance<P, A, MATRIX>(s, id, nActive, totdiam);
    out[t] = v;
    if (commit) s.outDist[id] = v;
}

// averageProfile (NJ.tcc:2067-2135) + profileDist(new,new) (NJ.tcc:3040-3043); ONE CTA:
// positions in parallel, then thread 0 adds the per-position self-distance terms in order.
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(256)
k_average(Store<P> s, int64_t oid, int64_t id1, int64_t id2, double bionjWeight, P diameterOut) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *termW = reinterpret_cast<double *>(smem);            // [Lp] w*w
    double *termT = termW + s.Lp;                                // [Lp] w*w*piece
    const View<P, A> p1 = make_view<P, A>(s, id1), p2 = make_view<P, A>(s, id2);
    const int64_t row = oid - s.nSeqs;
    uint8_t *oc = s.codes + oid * s.Lp;
    P *ow = s.weights + row * s.Lp;
    P *ov = s.vecs + row * s.Lp * A;
    for (int64_t pos = threadIdx.x; pos < s.Lp; pos += blockDim.x) {
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
        termW[pos] = tw; termT[pos] = tt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double top = 0, denom = 0;
        for (int64_t pos = 0; pos < s.L; pos++)
            if (termW[pos] > 0) { denom = xadd(denom, termW[pos]); top = xadd(top, termT[pos]); }
        s.selfweight[oid] = (P) (denom > 0 ? denom : 0.01);
        s.selfdist[oid] = (P) (denom > 0 ? top / denom : 1.0);
        s.diameter[oid] = diameterOut;
        s.active[id1] = 0; s.active[id2] = 0; s.active[oid] = 1;
    }
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

// outProfile, NJ.tcc:729-815: one thread per position walks the node list in ascending order
// (the accumulation order of the reference at -threads 1)
template<typename P, int A, bool MATRIX>
__global__ void __launch_bounds__(64)
k_outprofile_rebuild(Store<P> s, const int64_t *__restrict__ ids, int64_t n) {
    const int64_t pos = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (pos >= s.L) return;
    const double inweight = 1.0 / (double) n;                                              // :732
    P wout = 0;
    P f[A];
#pragma unroll
    for (int k = 0; k < A; k++) f[k] = 0;
    for (int64_t in = 0; in < n; in++) {
        const int64_t id = ids[in];
        const uint32_t c = s.codes[id * s.Lp + pos];
        P w;
        const P *fIn = nullptr;
        if (id < s.nSeqs) w = c != VFT_DEV_NOCODE ? (P) 1 : (P) 0;
        else {
            const int64_t row = id - s.nSeqs;
            w = s.weights[row * s.Lp + pos];
            if (c == VFT_DEV_NOCODE) fIn = s.vecs + (row * s.Lp + pos) * A;
        }
        wout = (P) xadd((double) wout, xmul((double) w, inweight));                        // :741
        if (w > 0) add_to_freq<P, A, MATRIX>(s, f, (double) w, c, fIn);                    // :771-774
    }
    if (wout <= 0) wout = (P) 1e-20;                                                       // :743-745
    s.ow[pos] = wout;
    normalize_freq<P, A, MATRIX>(s, f);                                                    // :789-794
#pragma unroll
    for (int k = 0; k < A; k++) s.ov[pos * A + k] = f[k];
    if (MATRIX) code_dist_row<P, A, MATRIX>(s, f, s.ocd + pos * A);                        // :801-803
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

struct vft_ctx {
    vft_config cfg;
    int A;
    int64_t N, M, L, Lp, maxnode;
    size_t ps;
    cudaStream_t stream;
    // device
    void *codes, *weights, *vecs, *ow, *ov, *ocd, *diameter, *selfdist, *selfweight, *outDist, *active, *tables;
    void *d_dist, *d_weight, *d_crit;          // [M] one-vs-all scratch
    uint64_t *d_keys;                          // [M]
    uint64_t *d_tkA, *d_tkB;                   // top-k ping-pong
    uint32_t *d_tvA, *d_tvB;
    void *d_rec;                               // [SORT_N] Rec
    int64_t *d_ids, *d_pi, *d_pj;              // staging for lists
    void *d_out1, *d_out2;
    int64_t listCap;
    // pinned host
    void *h_in, *h_out;
    size_t hCap;
    std::vector<uint8_t> activeHost;
    vft_counters cnt;
};

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
    return s;
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
extern "C" const char *vft_backend_name(void) { return "cuda-sm100a"; }

static int ensure_lists(vft_ctx *c, int64_t n) {
    if (n <= c->listCap) return VFT_OK;
    int64_t cap = std::max<int64_t>(n, 2 * c->listCap);
    cudaFree(c->d_ids); cudaFree(c->d_pi); cudaFree(c->d_pj); cudaFree(c->d_out1); cudaFree(c->d_out2);
    CK(cudaMalloc(&c->d_ids, cap * 8)); CK(cudaMalloc(&c->d_pi, cap * 8)); CK(cudaMalloc(&c->d_pj, cap * 8));
    CK(cudaMalloc(&c->d_out1, cap * 8)); CK(cudaMalloc(&c->d_out2, cap * 8));
    c->listCap = cap;
    return VFT_OK;
}

static int ensure_pinned(vft_ctx *c, size_t bytes) {
    if (bytes <= c->hCap) return VFT_OK;
    size_t cap = std::max(bytes, 2 * c->hCap);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    CK(cudaMallocHost(&c->h_in, cap)); CK(cudaMallocHost(&c->h_out, cap));
    c->hCap = cap;
    return VFT_OK;
}

extern "C" int vft_ctx_create(const vft_config *cfg, vft_ctx **out) {
    if (!cfg || !out) return fail(VFT_EINVAL, "null argument");
    if (cfg->nSeqs < 1 || cfg->nPos < 1) return fail(VFT_EINVAL, "nSeqs and nPos must be positive");
    if (cfg->nCodes != 4 && cfg->nCodes != 20) return fail(VFT_EINVAL, "nCodes must be 4 or 20");
    if (cfg->precision != 32 && cfg->precision != 64) return fail(VFT_EINVAL, "precision must be 32 or 64");
    if (2 * cfg->nSeqs >= 0xFFFF0000ll) return fail(VFT_EINVAL, "too many sequences for 32-bit sort indices");
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
    c->cfg = *cfg; c->A = cfg->nCodes; c->N = cfg->nSeqs; c->M = 2 * cfg->nSeqs; c->L = cfg->nPos;
    c->Lp = (cfg->nPos + 15) / 16 * 16; c->ps = cfg->precision / 8; c->maxnode = 0;
    std::memset(&c->cnt, 0, sizeof c->cnt);
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t ps = c->ps, Lp = (size_t) c->Lp, A = (size_t) c->A, N = (size_t) c->N, M = (size_t) c->M;
    CK(cudaMalloc(&c->codes, M * Lp));
    CK(cudaMalloc(&c->weights, N * Lp * ps));
    CK(cudaMalloc(&c->vecs, N * Lp * A * ps));
    CK(cudaMalloc(&c->ow, Lp * ps)); CK(cudaMalloc(&c->ov, Lp * A * ps)); CK(cudaMalloc(&c->ocd, Lp * A * ps));
    CK(cudaMalloc(&c->diameter, M * ps)); CK(cudaMalloc(&c->selfdist, M * ps)); CK(cudaMalloc(&c->selfweight, M * ps));
    CK(cudaMalloc(&c->outDist, M * ps)); CK(cudaMalloc(&c->active, M));
    CK(cudaMalloc(&c->tables, 840 * ps));
    CK(cudaMalloc(&c->d_dist, M * ps)); CK(cudaMalloc(&c->d_weight, M * ps)); CK(cudaMalloc(&c->d_crit, M * ps));
    CK(cudaMalloc(&c->d_keys, M * 8));
    const size_t nChunks = (M + SORT_N - 1) / SORT_N;
    CK(cudaMalloc(&c->d_tkA, nChunks * SORT_N * 8)); CK(cudaMalloc(&c->d_tkB, nChunks * SORT_N * 8));
    CK(cudaMalloc(&c->d_tvA, nChunks * SORT_N * 4)); CK(cudaMalloc(&c->d_tvB, nChunks * SORT_N * 4));
    CK(cudaMalloc(&c->d_rec, SORT_N * 32));
    CK(cudaMemsetAsync(c->codes, VFT_NOCODE, M * Lp, c->stream));
    CK(cudaMemsetAsync(c->weights, 0, N * Lp * ps, c->stream));
    CK(cudaMemsetAsync(c->ow, 0, Lp * ps, c->stream));
    CK(cudaMemsetAsync(c->ov, 0, Lp * A * ps, c->stream));
    CK(cudaMemsetAsync(c->ocd, 0, Lp * A * ps, c->stream));
    CK(cudaMemsetAsync(c->diameter, 0, M * ps, c->stream)); CK(cudaMemsetAsync(c->selfdist, 0, M * ps, c->stream));
    CK(cudaMemsetAsync(c->selfweight, 0, M * ps, c->stream)); CK(cudaMemsetAsync(c->outDist, 0, M * ps, c->stream));
    CK(cudaMemsetAsync(c->active, 0, M, c->stream));
    CK(cudaMemsetAsync(c->tables, 0, 840 * ps, c->stream));
    c->activeHost.assign(M, 0);
    int rc = ensure_lists(c, 4096);
    if (rc == VFT_OK) rc = ensure_pinned(c, 1 << 20);
    if (rc != VFT_OK) return rc;
    cudaFuncSetAttribute(k_topk_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_N * 12);
    cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_N * 12);
    CK(cudaStreamSynchronize(c->stream));
    *out = c;
    return VFT_OK;
}

extern "C" int vft_ctx_destroy(vft_ctx *c) {
    if (!c) return VFT_OK;
    cudaStreamSynchronize(c->stream);
    void *ptrs[] = {c->codes, c->weights, c->vecs, c->ow, c->ov, c->ocd, c->diameter, c->selfdist, c->selfweight,
                    c->outDist, c->active, c->tables, c->d_dist, c->d_weight, c->d_crit, c->d_keys, c->d_tkA, c->d_tkB,
                    c->d_tvA, c->d_tvB, c->d_rec, c->d_ids, c->d_pi, c->d_pj, c->d_out1, c->d_out2};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    cudaStreamDestroy(c->stream);
    delete c;
    return VFT_OK;
}

extern "C" int vft_upload_tables(vft_ctx *c, const void *distances, const void *eigenval, const void *eigentot,
                                 const void *codeFreq) {
    if (!c || !distances || !eigenval || !eigentot || !codeFreq) return fail(VFT_EINVAL, "null argument");
    const size_t ps = c->ps;
    char *h = (char *) c->h_in;
    std::memcpy(h, distances, 400 * ps); std::memcpy(h + 400 * ps, eigenval, 20 * ps);
    std::memcpy(h + 420 * ps, eigentot, 20 * ps); std::memcpy(h + 440 * ps, codeFreq, 400 * ps);
    CK(cudaMemcpyAsync(c->tables, h, 840 * ps, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VFT_OK;
}

extern "C" int vft_upload_leaves(vft_ctx *c, const uint8_t *codes) {
    if (!c || !codes) return fail(VFT_EINVAL, "null argument");
    // clamp unknowns to NOCODE on the way (NJ.tcc:449-452) into the padded rows
    std::vector<uint8_t> padded((size_t) c->N * c->Lp, VFT_NOCODE);
    for (int64_t i = 0; i < c->N; i++)
        for (int64_t p = 0; p < c->L; p++) {
            uint8_t cd = codes[i * c->L + p];
            padded[(size_t) i * c->Lp + p] = cd >= c->A ? VFT_NOCODE : cd;
        }
    CK(cudaMemcpyAsync(c->codes, padded.data(), padded.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->active, 0, c->M, c->stream));
#define CALL_INIT(P, A_, MX) k_init_leaves<P><<<(unsigned) ((c->N + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c))
    if (c->cfg.precision == 32) { CALL_INIT(float, 0, 0); } else { CALL_INIT(double, 0, 0); }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->cnt.launches++;
    std::fill(c->activeHost.begin(), c->activeHost.end(), 0);
    std::fill(c->activeHost.begin(), c->activeHost.begin() + c->N, 1);
    c->maxnode = c->N;
    return VFT_OK;
}

extern "C" int vft_outprofile_rebuild(vft_ctx *c, const int64_t *ids, int64_t n) {
    if (!c) return fail(VFT_EINVAL, "null argument");
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
#define CALL_REB(P, A_, MX) k_outprofile_rebuild<P, A_, MX><<<(unsigned) ((c->L + 63) / 64), 64, 0, c->stream>>>(make_store<P>(c), c->d_ids, n)
    VFT_DISPATCH(c, CALL_REB);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));     // h_in is reused by the next call
    c->cnt.launches++;
    return VFT_OK;
}

extern "C" int vft_outprofile_update(vft_ctx *c, int64_t old1, int64_t old2, int64_t newnode, int64_t nActiveOld) {
    if (!c || nActiveOld < 2 || newnode < c->N || newnode >= c->maxnode || old1 < 0 || old2 < 0 || old1 >= c->maxnode
        || old2 >= c->maxnode)
        return fail(VFT_EINVAL, "bad argument");
#define CALL_UPD(P, A_, MX) k_outprofile_update<P, A_, MX><<<(unsigned) ((c->L + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c), old1, old2, newnode, nActiveOld)
    VFT_DISPATCH(c, CALL_UPD);
    CK(cudaGetLastError());
    c->cnt.launches++;
    return VFT_OK;            // asynchronous: ordered on the context's stream
}

extern "C" int vft_profile_average(vft_ctx *c, int64_t out_id, int64_t id1, int64_t id2, double bionjWeight,
                                   double diameter_out) {
    if (!c || out_id < c->N || out_id >= c->M || id1 < 0 || id2 < 0 || id1 >= c->maxnode || id2 >= c->maxnode)
        return fail(VFT_EINVAL, "bad node id");
    if (bionjWeight < 0) bionjWeight = 0.5;
    const size_t smem = (size_t) c->Lp * 16;
#define CALL_AVG(P, A_, MX)                                                                                   \
    do {                                                                                                      \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_average<P, A_, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
        k_average<P, A_, MX><<<1, 256, smem, c->stream>>>(make_store<P>(c), out_id, id1, id2, bionjWeight, (P) diameter_out); \
    } while (0)
    if (smem > 200 * 1024) return fail(VFT_EINVAL, "alignment too long for the single-CTA average kernel");
    VFT_DISPATCH(c, CALL_AVG);
    CK(cudaGetLastError());
    c->cnt.launches++; c->cnt.profileAvgOps++; c->cnt.profileOps++;
    c->activeHost[id1] = 0; c->activeHost[id2] = 0; c->activeHost[out_id] = 1;
    if (out_id >= c->maxnode) c->maxnode = out_id + 1;
    return VFT_OK;            // asynchronous
}

extern "C" int vft_get_self(vft_ctx *c, int64_t id, double *selfdist, double *selfweight) {
    if (!c || id < 0 || id >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
    char buf[16];
    CK(cudaMemcpyAsync(buf, (char *) c->selfdist + id * c->ps, c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(buf + 8, (char *) c->selfweight + id * c->ps, c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->ps == 4) { *selfdist = *(float *) buf; *selfweight = *(float *) (buf + 8); }
    else { *selfdist = *(double *) buf; *selfweight = *(double *) (buf + 8); }
    return VFT_OK;
}

static int64_t profile_bytes(vft_ctx *c, int64_t id) {      // algorithmic bytes of one node's profile
    return id >= 0 && id < c->N ? c->L : c->L * ((int64_t) c->A * c->ps + c->ps + 1);
}

extern "C" int vft_out_distance_batch(vft_ctx *c, const int64_t *ids, int64_t n, int64_t nActive, double totdiam,
                                      void *outDist) {
    if (!c || (n > 0 && (!ids || !outDist))) return fail(VFT_EINVAL, "null argument");
    if (n == 0) return VFT_OK;
    for (int64_t k = 0; k < n; k++) if (ids[k] < 0 || ids[k] >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
    int rc = ensure_lists(c, n); if (rc) return rc;
    rc = ensure_pinned(c, (size_t) n * 8); if (rc) return rc;
    std::memcpy(c->h_in, ids, (size_t) n * 8);
    CK(cudaMemcpyAsync(c->d_ids, c->h_in, (size_t) n * 8, cudaMemcpyHostToDevice, c->stream));
#define CALL_OD(P, A_, MX) k_out_distance<P, A_, MX><<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c), c->d_ids, n, nActive, totdiam, (P *) c->d_out1, 0)
    VFT_DISPATCH(c, CALL_OD);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_out, c->d_out1, (size_t) n * c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(outDist, c->h_out, (size_t) n * c->ps);
    c->cnt.launches++; c->cnt.profileOps += n; c->cnt.outprofileOps += n;
    for (int64_t k = 0; k < n; k++) c->cnt.algoBytes += profile_bytes(c, ids[k]);
    c->cnt.algoBytes += profile_bytes(c, -1);
    return VFT_OK;
}

extern "C" int vft_out_distance_all(vft_ctx *c, int64_t nActive, double totdiam, void *outDist, int64_t maxnode) {
    if (!c || !outDist || maxnode < c->maxnode) return fail(VFT_EINVAL, "bad argument");
    const int64_t n = c->maxnode;
    int rc = ensure_lists(c, n); if (rc) return rc;
    rc = ensure_pinned(c, (size_t) n * 8); if (rc) return rc;
#define CALL_ODA(P, A_, MX) k_out_distance<P, A_, MX><<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c), nullptr, n, nActive, totdiam, (P *) c->d_out1, 1)
    VFT_DISPATCH(c, CALL_ODA);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_out, c->d_out1, (size_t) n * c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int64_t nAct = 0;
    for (int64_t i = 0; i < n; i++)
        if (c->activeHost[i]) {
            std::memcpy((char *) outDist + i * c->ps, (char *) c->h_out + i * c->ps, c->ps);
            nAct++; c->cnt.algoBytes += profile_bytes(c, i);
        }
    c->cnt.launches++; c->cnt.profileOps += nAct; c->cnt.outprofileOps += nAct;
    c->cnt.algoBytes += profile_bytes(c, -1);
    return VFT_OK;
}

extern "C" int vft_dist_pairs(vft_ctx *c, const int64_t *pi, const int64_t *pj, int64_t n, int32_t flags, void *dist,
                              void *weight) {
    if (!c || (n > 0 && (!pi || !pj || !dist || !weight))) return fail(VFT_EINVAL, "null argument");
    if (n == 0) return VFT_OK;
    int rc = ensure_lists(c, n); if (rc) return rc;
    rc = ensure_pinned(c, (size_t) n * 16); if (rc) return rc;
    int64_t *h = (int64_t *) c->h_in;
    const bool raw = (flags & VFT_PAIRS_PROFILE_RAW) != 0;
    for (int64_t k = 0; k < n; k++) {
        if (pi[k] < 0 || pj[k] < 0 || pi[k] >= c->maxnode || pj[k] >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
        h[k] = pi[k]; h[n + k] = pj[k];
        if (!raw && pi[k] < c->N && pj[k] < c->N) { c->cnt.seqOps++; c->cnt.algoBytes += c->L; }
        else { c->cnt.profileOps++; c->cnt.algoBytes += profile_bytes(c, pj[k]); }
    }
    c->cnt.algoBytes += profile_bytes(c, pi[0]);     // lists share their query; count it once
    CK(cudaMemcpyAsync(c->d_pi, h, (size_t) n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_pj, h + n, (size_t) n * 8, cudaMemcpyHostToDevice, c->stream));
#define CALL_DP(P, A_, MX) k_dist_pairs<P, A_, MX><<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c), c->d_pi, c->d_pj, n, raw ? 1 : 0, (P *) c->d_out1, (P *) c->d_out2)
    VFT_DISPATCH(c, CALL_DP);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_out, c->d_out1, (size_t) n * c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync((char *) c->h_out + (size_t) n * 8, c->d_out2, (size_t) n * c->ps, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(dist, c->h_out, (size_t) n * c->ps);
    std::memcpy(weight, (char *) c->h_out + (size_t) n * 8, (size_t) n * c->ps);
    c->cnt.launches++;
    return VFT_OK;
}

static int next_pow2(int64_t x) { int p = 1; while (p < x) p <<= 1; return p; }

extern "C" int vft_dist_one_vs_all(vft_ctx *c, int64_t query, int64_t nActive, int64_t K, int64_t *j_out, void *dist,
                                   void *weight, void *criterion, int64_t *nOut) {
    if (!c || !j_out || !dist || !weight || !criterion || !nOut) return fail(VFT_EINVAL, "null argument");
    if (query < 0 || query >= c->maxnode || !c->activeHost[query]) return fail(VFT_EINVAL, "query must be an active node");
    if (K < 1) return fail(VFT_EINVAL, "K must be positive");
    const int64_t n = c->maxnode;
    int Kc = next_pow2(std::min<int64_t>(K, n));
    if (Kc > SORT_N / 2) {
        if (K > SORT_N) return fail(VFT_EINVAL, "K larger than 4096 is not supported yet");
        Kc = SORT_N;
    }
#define CALL_OVA(P, A_, MX) k_one_vs_all<P, A_, MX><<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(make_store<P>(c), query, nActive, n, (P *) c->d_dist, (P *) c->d_weight, (P *) c->d_crit, c->d_keys)
    VFT_DISPATCH(c, CALL_OVA);
    CK(cudaGetLastError());
    c->cnt.launches++;
    // chunk sort, then a merge tree until one list is left
    int nLists = (int) ((n + SORT_N - 1) / SORT_N);
    const int KcStage = std::min(Kc, SORT_N);
    k_topk_chunks<<<nLists, SORT_T, SORT_N * 12, c->stream>>>(c->d_keys, n, KcStage, c->d_tkA, c->d_tvA);
    CK(cudaGetLastError());
    c->cnt.launches++;
    uint64_t *inK = c->d_tkA, *outK = c->d_tkB;
    uint32_t *inV = c->d_tvA, *outV = c->d_tvB;
    while (nLists > 1) {
        int R = std::max(2, SORT_N / KcStage);
        if (KcStage == SORT_N) return fail(VFT_EINVAL, "K larger than 2048 needs more than one chunk: not supported yet");
        int nOutLists = (nLists + R - 1) / R;
        k_topk_merge<<<nOutLists, SORT_T, SORT_N * 12, c->stream>>>(inK, inV, nLists, R, KcStage, outK, outV);
        CK(cudaGetLastError());
        c->cnt.launches++;
        std::swap(inK, outK); std::swap(inV, outV);
        nLists = nOutLists;
    }
    const int64_t nRet = std::min<int64_t>(std::min<int64_t>(K, nActive), KcStage);
    const size_t recSz = c->ps == 4 ? sizeof(Rec<float>) : sizeof(Rec<double>);
    int rc = ensure_pinned(c, (size_t) nRet * recSz); if (rc) return rc;
    if (c->ps == 4) k_gather_topk<float><<<(unsigned) ((nRet + 127) / 128), 128, 0, c->stream>>>(inV, (int) nRet, (float *) c->d_dist, (float *) c->d_weight, (float *) c->d_crit, (Rec<float> *) c->d_rec);
    else k_gather_topk<double><<<(unsigned) ((nRet + 127) / 128), 128, 0, c->stream>>>(inV, (int) nRet, (double *) c->d_dist, (double *) c->d_weight, (double *) c->d_crit, (Rec<double> *) c->d_rec);
    CK(cudaGetLastError());
    c->cnt.launches++;
    CK(cudaMemcpyAsync(c->h_out, c->d_rec, (size_t) nRet * recSz, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->ps == 4) {
        const Rec<float> *r = (const Rec<float> *) c->h_out;
        for (int64_t k = 0; k < nRet; k++) { j_out[k] = r[k].j; ((float *) dist)[k] = r[k].dist; ((float *) weight)[k] = r[k].weight; ((float *) criterion)[k] = r[k].crit; }
    } else {
        const Rec<double> *r = (const Rec<double> *) c->h_out;
        for (int64_t k = 0; k < nRet; k++) { j_out[k] = r[k].j; ((double *) dist)[k] = r[k].dist; ((double *) weight)[k] = r[k].weight; ((double *) criterion)[k] = r[k].crit; }
    }
    *nOut = nRet;
    // accounting: one distance per active node
    for (int64_t j = 0; j < n; j++)
        if (c->activeHost[j]) {
            if (query < c->N && j < c->N) { c->cnt.seqOps++; c->cnt.algoBytes += c->L; }
            else { c->cnt.profileOps++; c->cnt.algoBytes += profile_bytes(c, j); }
        }
    c->cnt.algoBytes += profile_bytes(c, query);
    return VFT_OK;
}

extern "C" int vft_get_profile(vft_ctx *c, int64_t id, void *weights, uint8_t *codes, void *vectors) {
    if (!c || id < -1 || id >= c->maxnode) return fail(VFT_EINVAL, "bad node id");
    const size_t ps = c->ps, L = (size_t) c->L, Lp = (size_t) c->Lp, A = (size_t) c->A;
    CK(cudaStreamSynchronize(c->stream));
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

extern "C" int vft_get_counters(vft_ctx *c, vft_counters *out) {
    if (!c || !out) return fail(VFT_EINVAL, "null argument");
    *out = c->cnt;
    return VFT_OK;
}
