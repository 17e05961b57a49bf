// vft_ingest.cuh -- ingest: the character decoding of seqsToProfiles (NJ.tcc:415-457) and Uniquify
// (Alignment.cpp:494-526) with the byte work on the device (included by vft_cuda.cu).
//
// The alignment text goes to the GPU once (1 byte per character, chunks of rows); per chunk one kernel decodes it into the
// uint8 code slab the NJ phase consumes and hashes every row twice (two independent 64-bit polynomial hashes, a warp per row,
// 16-byte loads).  The host groups the rows by their 128-bit key in input order -- a key seen before is CONFIRMED with a
// memcmp of the two text rows, so the result is exact whatever the hashes do -- which yields the reference's
// uniqueFirst / alnToUniq; a gather kernel then compacts the codes of the distinct rows, in first-occurrence order,
// straight into the layout vft_upload_leaves / vft_nj_build take.
#pragma once
#include <unordered_map>

namespace {

struct IngestLut { uint8_t code[256]; };

// warp per row: decode + two rolling hashes (order-dependent: per-lane partial hashes are combined in lane order)
__global__ void __launch_bounds__(256)
k_ingest(const uint8_t *__restrict__ text, int64_t nRows, int64_t nPos, const __grid_constant__ IngestLut lut,
         uint8_t *__restrict__ codes, uint64_t *__restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
    if (row >= nRows) return;
    const uint8_t *src = text + row * nPos;
    uint8_t *dst = codes + row * nPos;
    // position p contributes byte * M^p (mod 2^64) for two odd multipliers: a polynomial hash, evaluated as lane-strided
    // partial sums with running powers (exactly reproducible, independent of the launch shape)
    const uint64_t M1 = 0x9E3779B97F4A7C15ull, M2 = 0xC2B2AE3D27D4EB4Full;
    uint64_t p1 = 1, p2 = 1;
    for (int k = 0; k < lane; k++) { p1 *= M1; p2 *= M2; }
    uint64_t s1 = 1, s2 = 1;                            // M^32: the stride between a lane's consecutive positions
    for (int k = 0; k < 32; k++) { s1 *= M1; s2 *= M2; }
    uint64_t h1 = 0, h2 = 0;
    for (int64_t p = lane; p < nPos; p += 32) {
        const uint8_t ch = src[p];
        dst[p] = lut.code[ch];
        h1 += (uint64_t) (ch + 1) * p1; h2 += (uint64_t) (ch + 1) * p2;
        p1 *= s1; p2 *= s2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { h1 += __shfl_xor_sync(0xFFFFFFFFu, h1, o); h2 += __shfl_xor_sync(0xFFFFFFFFu, h2, o); }
    if (lane == 0) { keys[2 * row] = h1; keys[2 * row + 1] = h2; }
}

// codes of the distinct rows of this chunk, compacted: out row k = chunk row pick[k]
__global__ void __launch_bounds__(256)
k_ingest_gather(const uint8_t *__restrict__ codes, const int32_t *__restrict__ pick, int64_t nPick, int64_t nPos, uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t k = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
    if (k >= nPick) return;
    const uint8_t *src = codes + (int64_t) pick[k] * nPos;
    uint8_t *dst = out + k * nPos;
    for (int64_t p = lane; p < nPos; p += 32) dst[p] = src[p];
}

struct KeyHash { size_t operator()(const std::pair<uint64_t, uint64_t> &k) const { return (size_t) (k.first ^ (k.second * 0x9E3779B97F4A7C15ull)); } };

}  // namespace

extern "C" int vft_ingest(const char *text, int64_t nSeqs, int64_t nPos, const char *codesString, int32_t device,
                          int64_t *uniqueFirst, int64_t *alnToUniq, uint8_t *codes, int64_t *nUnique) {
    if (!text || !codesString || !uniqueFirst || !alnToUniq || !codes || !nUnique || nSeqs < 1 || nPos < 1) return fail(VFT_EINVAL, "bad argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        std::snprintf(g_err, sizeof g_err, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
        return VFT_ENODEVICE;
    }
    if (device < 0 || device >= ndev) return fail(VFT_EINVAL, "bad device ordinal");
    CK(cudaSetDevice(device));
    // charToCode (NJ.tcc:415-425): codesString position, both cases; everything else (incl. '-') is a gap
    IngestLut lut;
    const size_t nCodes = std::strlen(codesString);
    if (nCodes < 1 || nCodes > VFT_MAXCODES) return fail(VFT_EINVAL, "bad codesString");
    for (int c = 0; c < 256; c++) lut.code[c] = VFT_NOCODE;
    for (size_t i = 0; i < nCodes; i++) {
        lut.code[(unsigned char) codesString[i]] = (uint8_t) i;
        lut.code[(unsigned char) std::tolower((unsigned char) codesString[i])] = (uint8_t) i;
    }
    lut.code[(unsigned char) '-'] = VFT_NOCODE;
    // rows per chunk: ~256 MB of text at a time
    const int64_t chunkRows = std::max<int64_t>(1, std::min<int64_t>(nSeqs, ((int64_t) 256 << 20) / nPos));
    uint8_t *dText = nullptr, *dCodes = nullptr, *dOut = nullptr;
    uint64_t *dKeys = nullptr;
    int32_t *dPick = nullptr;
    cudaStream_t st = nullptr;
    std::vector<uint64_t> keys((size_t) 2 * chunkRows);
    std::vector<int32_t> pick;
    std::unordered_map<std::pair<uint64_t, uint64_t>, std::vector<int64_t>, KeyHash> seen;      // key -> first rows with that key (usually one)
    seen.reserve((size_t) nSeqs * 2);
    int rc = VFT_OK;
    int64_t nU = 0;
    auto cleanup = [&] { mem_free(dText); mem_free(dCodes); mem_free(dOut); mem_free(dKeys); mem_free(dPick); if (st) cudaStreamDestroy(st); };
#define CKI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = cuda_fail(e_, #call); cleanup(); return rc; } } while (0)
    CKI(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CKI(mem_alloc((void **) &dText, (size_t) chunkRows * nPos, MEM_DEVICE)); CKI(mem_alloc((void **) &dCodes, (size_t) chunkRows * nPos, MEM_DEVICE));
    CKI(mem_alloc((void **) &dOut, (size_t) chunkRows * nPos, MEM_DEVICE)); CKI(mem_alloc((void **) &dKeys, (size_t) chunkRows * 16, MEM_DEVICE));
    CKI(mem_alloc((void **) &dPick, (size_t) chunkRows * 4, MEM_DEVICE));
    for (int64_t r0 = 0; r0 < nSeqs; r0 += chunkRows) {
        const int64_t n = std::min(chunkRows, nSeqs - r0);
        CKI(cudaMemcpyAsync(dText, text + r0 * nPos, (size_t) n * nPos, cudaMemcpyHostToDevice, st));
        k_ingest<<<(unsigned) ((n * 32 + 255) / 256), 256, 0, st>>>(dText, n, nPos, lut, dCodes, dKeys);
        CKI(cudaGetLastError());
        CKI(cudaMemcpyAsync(keys.data(), dKeys, (size_t) n * 16, cudaMemcpyDeviceToHost, st));
        CKI(cudaStreamSynchronize(st));
        // Uniquify (Alignment.cpp:494-526): rows in input order; the first row of each distinct text becomes a unique sequence
        pick.clear();
        for (int64_t i = 0; i < n; i++) {
            const int64_t row = r0 + i;
            auto &cands = seen[{keys[(size_t) 2 * i], keys[(size_t) 2 * i + 1]}];
            int64_t first = -1;
            for (int64_t f : cands) if (std::memcmp(text + f * nPos, text + row * nPos, (size_t) nPos) == 0) { first = f; break; }
            if (first < 0) {
                cands.push_back(row);
                uniqueFirst[nU] = row; alnToUniq[row] = nU; nU++;
                pick.push_back((int32_t) i);
            } else alnToUniq[row] = alnToUniq[first];
        }
        if (!pick.empty()) {
            CKI(cudaMemcpyAsync(dPick, pick.data(), pick.size() * 4, cudaMemcpyHostToDevice, st));
            k_ingest_gather<<<(unsigned) ((pick.size() * 32 + 255) / 256), 256, 0, st>>>(dCodes, dPick, (int64_t) pick.size(), nPos, dOut);
            CKI(cudaGetLastError());
            CKI(cudaMemcpyAsync(codes + (nU - (int64_t) pick.size()) * nPos, dOut, pick.size() * (size_t) nPos, cudaMemcpyDeviceToHost, st));
            CKI(cudaStreamSynchronize(st));
        }
    }
#undef CKI
    cleanup();
    *nUnique = nU;
    return VFT_OK;
}
