// vft_dist.cuh -- ONE tree sharded over the GPUs of a node (included by vft_cuda.cu).
//
// Every rank (one process per GPU) runs the same deterministic host loop over a REPLICATED profile slab, so the serial
// join chain needs no communication at all: each rank builds every new node itself, bit for bit.  What is sharded is the
// candidate axis of the three all-candidate sweeps, the work that grows as N^1.5:
//   vft_dist_one_vs_all   (setBestHit, NJ.tcc:3571-3646)     rank r evaluates slots r, r+W, r+2W, ... of the ascending active
//                                                             list, selects ITS K best, the W x K records are all-gathered
//                                                             and merged by rank (k_rank_merge) -- one exchange per search
//   vft_out_distance_all  (NJ.tcc:257-260, :4451-4464)       the same strided share, values all-gathered and scattered into
//                                                             every rank's out-distance table
//   vft_tophits_merge     (NJ.tcc:4477-4515)                  the m lists of a refresh in W contiguous chunks, the saved lists
//                                                             all-gathered
// The exchange step has three implementations behind dist_allgather():
//   PEER (default when every GPU can map its peers): each rank's exchange buffer is opened on every peer with CUDA IPC; ONE
//        kernel (k_peer_allgather) pushes the local payload into every peer's slot with 128-bit stores over NVLink, publishes
//        a system-scope flag per peer and then waits for the W flags addressed to it -- an all-gather in one launch,
//        no host involvement, no proxy thread, double-buffered by exchange parity;
//   NCCL: ncclAllGather on the context's stream (libnccl resolved with dlopen: the library has no link-time dependency);
//   HOST: a caller-provided host all-gather (MPI, gloo, ...) through pinned staging -- bring-up and the CPU tests' double.
// A context created on the group's device after vft_dist_init is sharded; everything else is unchanged.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

enum { DIST_NONE = 0, DIST_NCCL = 1, DIST_PEER = 2, DIST_HOST = 3 };
constexpr int DIST_MAXW = 16;

struct PeerTable { char *data[DIST_MAXW]; unsigned long long *flags[DIST_MAXW]; };
struct ExpectTable { unsigned long long v[DIST_MAXW]; };

struct DistGroup {
    bool ready = false;
    int rank = 0, world = 1, device = -1, mode = DIST_NONE;
    NcclApi api;
    ncclComm_t comm = nullptr;
    vft_allgather_fn hostFn = nullptr; void *hostUser = nullptr;
    char *send = nullptr; size_t sendCap = 0;              // this rank's payload
    char *recv = nullptr; size_t recvCap = 0;              // NCCL / HOST: [W][bytes]
    void *hSend = nullptr, *hRecv = nullptr; size_t hCap = 0;
    // PEER: xbuf = [2 parities][W slots][slotBytes] | flags[DIST_MAXW]
    char *xbuf = nullptr; size_t slotBytes = 0;
    PeerTable peers;
    ExpectTable expect;
    unsigned long long seq = 0;
    int64_t nExchanges = 0; int64_t bytesExchanged = 0;
    int liveContexts = 0;                                  // sharded contexts alive: vft_dist_finalize refuses while > 0
};
DistGroup g_dist;

int nccl_fail(ncclResult_t r, const char *what) {
    std::snprintf(g_err, sizeof g_err, "%s: %s", what, g_dist.api.GetErrorString ? g_dist.api.GetErrorString(r) : "NCCL error");
    return VFT_ECUDA;
}
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return nccl_fail(r_, #call); } while (0)

int nccl_load() {
    NcclApi &a = g_dist.api;
    if (a.lib) return VFT_OK;
    const char *env = std::getenv("VFT_NCCL_LIB");
    void *h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy the process already uses (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(VFT_EINVAL, "libnccl.so.2 not found (set VFT_NCCL_LIB)");
    a.GetUniqueId = (decltype(a.GetUniqueId)) dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank)) dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy)) dlsym(h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather)) dlsym(h, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString)) dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather) return fail(VFT_EINVAL, "libnccl lacks the expected entry points");
    a.lib = h;
    return VFT_OK;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// The all-gather over peer memory.  Grid = W x CT CTAs: CTA (w, part) pushes its share of the local payload into slot `rank`
// of peer w (peer == self included) with 128-bit stores, makes them visible at system scope and adds 1 to the flag that peer w
// keeps for this rank; CTA (w, 0) then waits until the local flag of SOURCE w has reached its expected count (CT per exchange).
// Every CTA pushes before it waits and no CTA waits on its own GPU's progress, so the kernel cannot deadlock as long as every
// rank launches the same sequence of exchanges (they do: the host loops are replicas).  Slots are double-buffered by the parity
// of the exchange number: a rank can run at most one exchange ahead of the slowest one, because finishing exchange s+1
// needs every peer's push of s+1, which each peer issues only after ITS exchange s completed.
__global__ void __launch_bounds__(256)
k_peer_allgather(const char *__restrict__ src, size_t bytes, PeerTable pt, int rank, int CT, size_t slotOff, ExpectTable ex) {
    const int w = blockIdx.x / CT, part = blockIdx.x % CT;
    uint4 *dst = reinterpret_cast<uint4 *>(pt.data[w] + slotOff);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    const size_t n16 = bytes >> 4;
    for (size_t i = (size_t) part * blockDim.x + threadIdx.x; i < n16; i += (size_t) CT * blockDim.x) dst[i] = s4[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd_system(pt.flags[w] + rank, 1ull);
        if (part == 0) {
            // bounded wait (~30 s of SM clock): a peer that died must not hang this GPU; the timeout is latched in the
            // error word behind the flags and reported by the next host-side check (dist_allgather / vft_dist_info)
            const unsigned long long *mine = pt.flags[rank] + w;
            const long long t0 = clock64();
            while (ld_acquire_sys(mine) < ex.v[w]) {
                if (clock64() - t0 > 60000000000ll) { atomicExch((unsigned long long *) (pt.flags[rank] + DIST_MAXW), 1ull); break; }
                __nanosleep(64);
            }
        }
    }
}

// out-distances of a sharded sweep, gathered as [W][chunk] (slot i of rank w = entry i*W + w of the active list), committed to
// this rank's table
template<typename P>
__global__ void k_scatter_outdist(const char *__restrict__ base, size_t stride, const int32_t *__restrict__ act, int64_t nAct, int W,
                                  P *__restrict__ outDist) {
    const int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (e >= nAct) return;
    const int w = (int) (e % W);
    const int64_t i = e / W;
    outDist[act[e]] = reinterpret_cast<const P *>(base + (size_t) w * stride)[i];
}

// W sorted record lists (each rank's K best in psort order: criterion ascending, node id descending) -> the K best overall.
// Every record computes its global rank = its own position + the number of records of every OTHER list that precede it
// (binary search; node ids are disjoint between ranks, so the composite order is strict) and, if < K, writes itself there.
template<typename P>
__device__ __forceinline__ bool rec_before(const Rec<P> &a, uint64_t kb, int64_t jb) {      // a strictly before (kb, jb)?
    const uint64_t ka = order_key(a.crit);
    return ka < kb || (ka == kb && a.j > jb);
}
template<typename P>
__global__ void k_rank_merge(const char *__restrict__ base, size_t stride, int W, int64_t nAct, int K, Rec<P> *__restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int w = (int) (e / K), t = (int) (e % K);
    if (w >= W) return;
    const int64_t len = (nAct - w + W - 1) / W;
    const int cnt = (int) (len < K ? len : K);
    if (t >= cnt) return;
    const Rec<P> me = reinterpret_cast<const Rec<P> *>(base + (size_t) w * stride)[t];
    const uint64_t km = order_key(me.crit);
    int64_t rank = t;
    for (int o = 0; o < W; o++) {
        if (o == w) continue;
        const Rec<P> *lst = reinterpret_cast<const Rec<P> *>(base + (size_t) o * stride);
        const int64_t lo_ = (nAct - o + W - 1) / W;
        int lo = 0, hi = (int) (lo_ < K ? lo_ : K);
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (rec_before<P>(lst[mid], km, me.j)) lo = mid + 1; else hi = mid; }
        rank += lo;
    }
    if (rank < K) out[rank] = me;
}

int dist_release_buffers() {
    DistGroup &g = g_dist;
    if (g.xbuf) {
        for (int w = 0; w < g.world; w++) if (w != g.rank && g.peers.data[w]) { cudaIpcCloseMemHandle(g.peers.data[w]); g.peers.data[w] = nullptr; }
        if (g.comm && g.send && g.recv) {             // nobody frees a buffer a peer may still have mapped (finalize is collective)
            g.api.AllGather(g.send, g.recv, 16, ncclUint8, g.comm, (cudaStream_t) 0);
            cudaDeviceSynchronize();
        }
        cudaFree(g.xbuf);
        g.xbuf = nullptr; g.slotBytes = 0;
    }
    if (g.send) { cudaFree(g.send); g.send = nullptr; g.sendCap = 0; }
    if (g.recv) { cudaFree(g.recv); g.recv = nullptr; g.recvCap = 0; }
    if (g.hSend) { cudaFreeHost(g.hSend); cudaFreeHost(g.hRecv); g.hSend = g.hRecv = nullptr; g.hCap = 0; }
    return VFT_OK;
}

// room for `bytes` per rank in the exchange buffers (collective in PEER mode when the slots have to grow: every rank reaches
// this call with the same argument)
int dist_reserve(cudaStream_t stream, size_t bytes) {
    DistGroup &g = g_dist;
    bytes = (bytes + 15) & ~(size_t) 15;
    if (bytes > g.sendCap) {
        // 1 MB to start with (VFT_XBUF_INIT_KB: a smaller start, so that the tests reach the growth path), doubled as needed
        static const size_t initCap = [] { const char *e = std::getenv("VFT_XBUF_INIT_KB"); const long kb = e ? std::atol(e) : 1024; return (size_t) (kb >= 4 ? kb : 4) << 10; }();
        size_t cap = initCap;
        while (cap < bytes) cap <<= 1;
        CK(cudaStreamSynchronize(stream));
        if (g.send) cudaFree(g.send);
        g.send = nullptr; g.sendCap = 0;
        CK(cudaMalloc((void **) &g.send, cap));
        g.sendCap = cap;
    }
    if (g.mode == DIST_NCCL || g.mode == DIST_HOST || g.mode == DIST_PEER) {           // PEER keeps recv for the handle exchange
        const size_t need = g.mode == DIST_PEER ? (size_t) g.world * 128 : (size_t) g.world * g.sendCap;
        if (need > g.recvCap) {
            CK(cudaStreamSynchronize(stream));
            if (g.recv) cudaFree(g.recv);
            g.recv = nullptr; g.recvCap = 0;
            CK(cudaMalloc((void **) &g.recv, need));
            g.recvCap = need;
        }
    }
    if (g.mode == DIST_HOST && (size_t) g.world * g.sendCap > g.hCap) {
        if (g.hSend) { cudaFreeHost(g.hSend); cudaFreeHost(g.hRecv); }
        g.hSend = g.hRecv = nullptr; g.hCap = 0;
        CK(cudaHostAlloc(&g.hSend, g.sendCap, cudaHostAllocDefault)); CK(cudaHostAlloc(&g.hRecv, (size_t) g.world * g.sendCap, cudaHostAllocDefault));
        g.hCap = (size_t) g.world * g.sendCap;
    }
    if (g.mode == DIST_PEER && g.sendCap > g.slotBytes) {
        // (re)build the peer-mapped exchange buffer: local allocation, handles all-gathered through NCCL, peers opened
        CK(cudaStreamSynchronize(stream));
        if (g.xbuf) {
            // every rank first drops its mappings of the peers' old buffers; only when ALL have done so (a small all-gather as
            // the barrier) does anybody free the buffer the others had mapped
            for (int w = 0; w < g.world; w++) if (w != g.rank && g.peers.data[w]) { cudaIpcCloseMemHandle(g.peers.data[w]); g.peers.data[w] = nullptr; }
            NK(g.api.AllGather(g.send, g.recv, 16, ncclUint8, g.comm, stream));
            CK(cudaStreamSynchronize(stream));
            cudaFree(g.xbuf); g.xbuf = nullptr;
        }
        const size_t slot = g.sendCap, dataBytes = 2 * (size_t) g.world * slot;
        CK(cudaMalloc((void **) &g.xbuf, dataBytes + 4096));
        CK(cudaMemset(g.xbuf + dataBytes, 0, 4096));
        cudaIpcMemHandle_t mine;
        CK(cudaIpcGetMemHandle(&mine, g.xbuf));
        static_assert(sizeof(cudaIpcMemHandle_t) <= 128, "handle size");
        char hbuf[128] = {0};
        std::memcpy(hbuf, &mine, sizeof mine);
        CK(cudaMemcpyAsync(g.send, hbuf, 128, cudaMemcpyHostToDevice, stream));
        NK(g.api.AllGather(g.send, g.recv, 128, ncclUint8, g.comm, stream));
        std::vector<char> all((size_t) g.world * 128);
        CK(cudaMemcpyAsync(all.data(), g.recv, all.size(), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (int w = 0; w < g.world; w++) {
            char *p = g.xbuf;
            if (w != g.rank) {
                cudaIpcMemHandle_t h;
                std::memcpy(&h, all.data() + (size_t) w * 128, sizeof h);
                void *q = nullptr;
                CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
                p = (char *) q;
            }
            g.peers.data[w] = p;
            g.peers.flags[w] = reinterpret_cast<unsigned long long *>(p + dataBytes);
            g.expect.v[w] = 0;
        }
        g.slotBytes = slot; g.seq = 0;
    }
    return VFT_OK;
}

// all-gather of `bytes` bytes per rank out of g.send, ordered on `stream`; *base / *stride locate rank w's payload afterwards
int dist_allgather(cudaStream_t stream, size_t bytes, const char **base, size_t *stride) {
    DistGroup &g = g_dist;
    bytes = (bytes + 15) & ~(size_t) 15;
    if (bytes > g.sendCap) return fail(VFT_EINVAL, "exchange larger than reserved");
    g.nExchanges++; g.bytesExchanged += (int64_t) bytes * g.world;
    if (g.mode == DIST_NCCL) {
        NK(g.api.AllGather(g.send, g.recv, bytes, ncclUint8, g.comm, stream));
        *base = g.recv; *stride = bytes;
        return VFT_OK;
    }
    if (g.mode == DIST_HOST) {
        CK(cudaMemcpyAsync(g.hSend, g.send, bytes, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (g.hostFn(g.hSend, g.hRecv, (int64_t) bytes, g.hostUser) != 0) return fail(VFT_EINVAL, "the caller's all-gather failed");
        CK(cudaMemcpyAsync(g.recv, g.hRecv, (size_t) g.world * bytes, cudaMemcpyHostToDevice, stream));
        *base = g.recv; *stride = bytes;
        return VFT_OK;
    }
    // PEER
    {   // a peer that timed out in an earlier exchange left its mark: fail loudly instead of computing on garbage
        unsigned long long err = 0;
        if (g.nExchanges % 64 == 0) {
            CK(cudaMemcpyAsync(&err, g.peers.flags[g.rank] + DIST_MAXW, 8, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            if (err) return fail(VFT_ECUDA, "peer-memory exchange timed out: a rank of the group stopped responding");
        }
    }
    const int CT = (int) std::min<size_t>(16, std::max<size_t>(1, bytes >> 16));           // one CTA per 64 KB and peer, at most 16
    const size_t parityOff = (size_t) (g.seq & 1) * g.world * g.slotBytes;
    for (int w = 0; w < g.world; w++) g.expect.v[w] += (unsigned long long) CT;
    k_peer_allgather<<<g.world * CT, 256, 0, stream>>>(g.send, bytes, g.peers, g.rank, CT, parityOff + (size_t) g.rank * g.slotBytes, g.expect);
    CK(cudaGetLastError());
    *base = g.xbuf + parityOff; *stride = g.slotBytes;
    g.seq++;
    return VFT_OK;
}

}  // namespace

extern "C" int vft_dist_unique_id(void *id128) {
    if (!id128) return fail(VFT_EINVAL, "null argument");
    int rc = nccl_load(); if (rc) return rc;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NK(g_dist.api.GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return VFT_OK;
}

extern "C" int vft_dist_finalize(void) {
    DistGroup &g = g_dist;
    if (!g.ready) return VFT_OK;
    if (g.liveContexts > 0) return fail(VFT_EINVAL, "vft_dist_finalize: sharded contexts are still alive (destroy them first)");
    if (g.device >= 0) cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    dist_release_buffers();
    if (g.comm) { g.api.CommDestroy(g.comm); g.comm = nullptr; }
    g.ready = false; g.mode = DIST_NONE; g.world = 1; g.rank = 0; g.device = -1; g.hostFn = nullptr;
    return VFT_OK;
}

extern "C" int vft_dist_init(int32_t rank, int32_t world, const void *id128, int32_t device) {
    DistGroup &g = g_dist;
    if (g.ready) return fail(VFT_EINVAL, "a group is already initialised (vft_dist_finalize first)");
    if (world < 1 || world > DIST_MAXW || rank < 0 || rank >= world || (world > 1 && !id128)) return fail(VFT_EINVAL, "bad rank / world");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(VFT_ENODEVICE, "no CUDA device; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(VFT_EINVAL, "bad device ordinal");
    CK(cudaSetDevice(device));
    g.rank = rank; g.world = world; g.device = device; g.seq = 0; g.nExchanges = 0; g.bytesExchanged = 0;
    if (world == 1) { g.mode = DIST_NONE; g.ready = true; return VFT_OK; }
    int rc = nccl_load(); if (rc) return rc;
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    NK(g.api.CommInitRank(&g.comm, world, id, rank));
    g.mode = DIST_NCCL;
    g.ready = true;
    // peer mode unless disabled: needs every pair of GPUs to be peer-capable; all ranks must agree, so the local verdict is
    // all-gathered (one byte per rank) before anybody switches
    const char *env = std::getenv("VFT_EXCHANGE");
    const bool wantPeer = !(env && (env[0] == 'n' || env[0] == 'N'));                       // VFT_EXCHANGE=nccl
    unsigned char ok = wantPeer ? 1 : 0;
    for (int d = 0; ok && d < ndev && d < world; d++) {
        // ranks use devices 0..world-1 of one node (LOCAL_RANK); a peer that cannot be mapped turns the mode off
        int can = 0;
        if (d != device && (cudaDeviceCanAccessPeer(&can, device, d) != cudaSuccess || !can)) ok = 0;
    }
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    rc = dist_reserve(st, 4096);
    if (rc == VFT_OK) {
        cudaMemcpyAsync(g.send, &ok, 1, cudaMemcpyHostToDevice, st);
        ncclResult_t r = g.api.AllGather(g.send, g.recv, 16, ncclUint8, g.comm, st);
        std::vector<unsigned char> all((size_t) world * 16);
        cudaMemcpyAsync(all.data(), g.recv, all.size(), cudaMemcpyDeviceToHost, st);
        if (r != ncclSuccess || cudaStreamSynchronize(st) != cudaSuccess) rc = fail(VFT_ECUDA, "NCCL all-gather failed during vft_dist_init");
        else {
            bool allOk = true;
            for (int w = 0; w < world; w++) allOk = allOk && all[(size_t) w * 16] == 1;
            if (allOk) {
                g.mode = DIST_PEER;
                rc = dist_reserve(st, 4096);                    // builds and maps the exchange buffer
                if (rc != VFT_OK) { cudaGetLastError(); g.mode = DIST_NCCL; }
                // the mapping must have worked everywhere
                unsigned char ok2 = rc == VFT_OK ? 1 : 0;
                cudaMemcpyAsync(g.send, &ok2, 1, cudaMemcpyHostToDevice, st);
                r = g.api.AllGather(g.send, g.recv, 16, ncclUint8, g.comm, st);
                cudaMemcpyAsync(all.data(), g.recv, all.size(), cudaMemcpyDeviceToHost, st);
                rc = (r != ncclSuccess || cudaStreamSynchronize(st) != cudaSuccess) ? fail(VFT_ECUDA, "NCCL all-gather failed during vft_dist_init") : VFT_OK;
                for (int w = 0; w < world; w++) if (all[(size_t) w * 16] != 1) g.mode = DIST_NCCL;
            }
        }
    }
    cudaStreamDestroy(st);
    if (rc != VFT_OK) { char keep[sizeof g_err]; std::memcpy(keep, g_err, sizeof keep); vft_dist_finalize(); std::memcpy(g_err, keep, sizeof keep); }
    return rc;
}

extern "C" int vft_dist_init_host(int32_t rank, int32_t world, vft_allgather_fn fn, void *user, int32_t device) {
    DistGroup &g = g_dist;
    if (g.ready) return fail(VFT_EINVAL, "a group is already initialised (vft_dist_finalize first)");
    if (world < 1 || world > DIST_MAXW || rank < 0 || rank >= world || (world > 1 && !fn)) return fail(VFT_EINVAL, "bad rank / world");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(VFT_ENODEVICE, "no CUDA device; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(VFT_EINVAL, "bad device ordinal");
    g.rank = rank; g.world = world; g.device = device; g.hostFn = fn; g.hostUser = user; g.seq = 0; g.nExchanges = 0; g.bytesExchanged = 0;
    g.mode = world == 1 ? DIST_NONE : DIST_HOST;
    g.ready = true;
    return VFT_OK;
}

extern "C" int vft_dist_info(int32_t *rank, int32_t *world, int32_t *mode, int64_t *nExchanges, int64_t *bytesExchanged) {
    if (rank) *rank = g_dist.rank;
    if (world) *world = g_dist.ready ? g_dist.world : 1;
    if (mode) *mode = g_dist.ready ? g_dist.mode : DIST_NONE;
    if (nExchanges) *nExchanges = g_dist.nExchanges;
    if (bytesExchanged) *bytesExchanged = g_dist.bytesExchanged;
    return VFT_OK;
}
