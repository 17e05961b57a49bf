// vft_device.cuh -- device-side arithmetic of the NJ hot path (sm_100a).
//
// Everything here is "reference-order arithmetic": each value is produced by the same sequence
// of individually rounded operations, in the same types, as the reference's -mavx2 (no FMA)
// build evaluates it (citations: /root/reference/src, NJ.tcc = NeighbourJoining.tcc).  That is
// what makes the top-hit indices identical to the reference rather than merely close: the
// criterion has many exact and near ties, and the order of the double-precision accumulation
// over positions decides them.  Consequences for the kernel design:
//   - a pair's distance is accumulated by ONE thread, positions in ascending order (the
//     per-position work is cheap; parallelism comes from the thousands of pairs in a batch);
//   - mul and add are kept separate (__dmul_rn/__dadd_rn, __fmul_rn/__fadd_rn; the file is also
//     compiled with -fmad=false);
//   - the P-typed 20-wide dot products reproduce the lane order of AVX256Operations.tcc.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define VFT_DEV_NOCODE 127u

namespace vft {

// ---- separately rounded arithmetic in P and in double ------------------------------------------
__device__ __forceinline__ float  pmul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double pmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  padd(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double padd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  psub(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ double psub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }

template<typename P>
struct Store {
    // profile slab (HBM): leaves are 1 byte/position, internal nodes are dense
    uint8_t *codes;        // [2*nSeqs][Lp]            NOCODE in the padding
    P *weights;            // [nSeqs][Lp]              internal node id -> row id-nSeqs; 0 in the padding
    P *vecs;               // [nSeqs][Lp][A]
    // out-profile (always dense, NJ.tcc:746)
    P *ow, *ov, *ocd;      // [Lp], [Lp][A], [Lp][A] (codeDist, matrix mode)
    // per node
    P *diameter, *selfdist, *selfweight, *outDist;   // [2*nSeqs]
    uint8_t *active;       // [2*nSeqs]
    // tables (DistanceMatrix.h:15-33), row stride 20
    const P *distances, *eigenval, *eigentot, *codeFreq;
    int64_t nSeqs, L, Lp;
    int reduction;         // VFT_REDUCE_*
    double fPostTotalTolerance;
};

// ---- lane-ordered reductions: BasicOperations.tcc:17-41, AVX256Operations.tcc:5-26,58-138 -------
template<typename P, int A>
__device__ __forceinline__ P lane_fold(const P (&prod)[A], int mode) {
    if (mode == 0) {
        P out = 0;
#pragma unroll
        for (int i = 0; i < A; i++) out = padd(out, prod[i]);
        return out;
    }
    if (sizeof(P) == 8) {                       // 4 lane accumulators, then (l0+l1)+(l2+l3)
        P l0 = 0, l1 = 0, l2 = 0, l3 = 0;
#pragma unroll
        for (int i = 0; i < A; i += 4) {
            l0 = padd(prod[i], l0); l1 = padd(prod[i + 1], l1); l2 = padd(prod[i + 2], l2); l3 = padd(prod[i + 3], l3);
        }
        return padd(padd(l0, l1), padd(l2, l3));
    }
    if (A == 4) return padd(padd(prod[0], prod[1]), padd(prod[2], prod[3]));
    {                                           // float, A == 20: two 8-lane blocks + a 4-lane tail
        P l[8];
#pragma unroll
        for (int k = 0; k < 8; k++) l[k] = 0;
        constexpr int m = A - (A % 8);
#pragma unroll
        for (int i = 0; i < m; i += 8)
#pragma unroll
            for (int k = 0; k < 8; k++) l[k] = padd(prod[i + k], l[k]);
        P t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = padd(l[k], prod[(m + k) < A ? (m + k) : 0]);
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = padd(l[4 + k], t[k]);
        return padd(padd(t[0], t[1]), padd(t[2], t[3]));
    }
}

template<typename P, int A>
__device__ __forceinline__ P vec_mul_sum(const P *f1, const P *f2, int mode) {        // vector_multiply_sum
    P prod[A];
#pragma unroll
    for (int i = 0; i < A; i++) prod[i] = pmul(f1[i], f2[i]);
    return lane_fold<P, A>(prod, mode);
}

template<typename P, int A>
__device__ __forceinline__ P vec_mul3_sum(const P *f1, const P *f2, const P *f3, int mode) {   // vector_multiply3_sum
    P prod[A];
#pragma unroll
    for (int i = 0; i < A; i++) prod[i] = pmul(pmul(f1[i], f2[i]), f3[i]);
    return lane_fold<P, A>(prod, mode);
}

// ---- a node as the distance loops see it -------------------------------------------------------
template<typename P, int A>
struct View {
    const uint8_t *codes;   // nullptr for the out-profile (every code is NOCODE)
    const P *w;             // nullptr for leaves (weight = code != NOCODE)
    const P *v;             // nullptr for leaves
    const P *cd;            // codeDist (out-profile, matrix mode) or nullptr
};

template<typename P, int A>
__device__ __forceinline__ View<P, A> make_view(const Store<P> &s, int64_t id) {
    View<P, A> r;
    if (id < 0) { r.codes = nullptr; r.w = s.ow; r.v = s.ov; r.cd = s.ocd; }
    else if (id < s.nSeqs) { r.codes = s.codes + id * s.Lp; r.w = nullptr; r.v = nullptr; r.cd = nullptr; }
    else {
        int64_t row = id - s.nSeqs;
        r.codes = s.codes + id * s.Lp; r.w = s.weights + row * s.Lp; r.v = s.vecs + row * s.Lp * A; r.cd = nullptr;
    }
    return r;
}

__device__ __forceinline__ uint32_t code_of(const uint4 &q, int b) {
    uint32_t w = (b < 4) ? q.x : (b < 8) ? q.y : (b < 12) ? q.z : q.w;
    return (w >> ((b & 3) * 8)) & 0xFFu;
}

// profileDistPiece, NJ.tcc:900-941, for a position where both weights are > 0
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ double piece(const Store<P> &s, uint32_t c1, uint32_t c2,
                                        const P *f1, const P *f2, const P *cd2) {
    if (MATRIX) {
        if (c1 != VFT_DEV_NOCODE && c2 != VFT_DEV_NOCODE) return (double) s.distances[c1 * 20 + c2];
        if (cd2 != nullptr && c1 != VFT_DEV_NOCODE) return (double) cd2[c1];
        P a[A], b[A], e[A];
        const P *p1 = (c1 != VFT_DEV_NOCODE) ? s.codeFreq + c1 * 20 : f1;
        const P *p2 = (c2 != VFT_DEV_NOCODE) ? s.codeFreq + c2 * 20 : f2;
#pragma unroll
        for (int k = 0; k < A; k++) { a[k] = p1[k]; b[k] = p2[k]; e[k] = s.eigenval[k]; }
        return (double) vec_mul3_sum<P, A>(a, b, e, s.reduction);
    } else {
        if (c1 != VFT_DEV_NOCODE) {
            if (c2 != VFT_DEV_NOCODE) return c1 == c2 ? 0.0 : 1.0;
            return xsub(1.0, (double) f2[c1]);
        }
        if (c2 != VFT_DEV_NOCODE) return xsub(1.0, (double) f1[c2]);
        double pc = 1.0;
#pragma unroll
        for (int k = 0; k < A; k++) pc = xsub(pc, (double) pmul(f1[k], f2[k]));
        return pc;
    }
}

// profileDist, NJ.tcc:1167-1190: one thread, positions in ascending order
template<typename P, int A, bool MATRIX>
__device__ void profile_dist(const Store<P> &s, const View<P, A> &p1, const View<P, A> &p2, P &dist, P &weight) {
    double top = 0, denom = 0;
    const int64_t Lp = s.Lp;
    for (int64_t base = 0; base < Lp; base += 16) {
        uint4 q1 = p1.codes ? *reinterpret_cast<const uint4 *>(p1.codes + base)
                            : make_uint4(0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu);
        uint4 q2 = p2.codes ? *reinterpret_cast<const uint4 *>(p2.codes + base)
                            : make_uint4(0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu);
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const int64_t pos = base + b;
            const uint32_t c1 = code_of(q1, b), c2 = code_of(q2, b);
            const P w1 = p1.w ? p1.w[pos] : (c1 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
            const P w2 = p2.w ? p2.w[pos] : (c2 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
            if (w1 > 0 && w2 > 0) {
                const double wt = (double) pmul(w1, w2);                                  // :1176
                denom = xadd(denom, wt);
                const double pc = piece<P, A, MATRIX>(s, c1, c2, p1.v ? p1.v + pos * A : nullptr,
                                                      p2.v ? p2.v + pos * A : nullptr, p2.cd ? p2.cd + pos * A : nullptr);
                top = xadd(top, xmul(wt, pc));
            }
        }
    }
    weight = (P) (denom > 0 ? denom : 0.01);                                              // :1187
    dist = (P) (denom > 0 ? top / denom : 1.0);                                           // :1188
}

// seqDist, NJ.tcc:1601-1624
template<typename P, bool MATRIX>
__device__ void seq_dist(const Store<P> &s, const uint8_t *c1, const uint8_t *c2, P &dist, P &weight) {
    const int64_t Lp = s.Lp;
    int nUse = 0;
    double top = 0;
    if (!MATRIX) {
        int nDiff = 0;
        for (int64_t base = 0; base < Lp; base += 16) {
            const uint4 a = *reinterpret_cast<const uint4 *>(c1 + base);
            const uint4 b = *reinterpret_cast<const uint4 *>(c2 + base);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // per-byte masks: 0xFF where the byte is a real code on both sides / where they differ
                const uint32_t use = __vcmpne4(aw[k], 0x7F7F7F7Fu) & __vcmpne4(bw[k], 0x7F7F7F7Fu);
                const uint32_t ne = __vcmpne4(aw[k], bw[k]);
                nUse += __popc(use) >> 3;
                nDiff += __popc(use & ne) >> 3;
            }
        }
        top = (double) nDiff;
    } else {
        for (int64_t base = 0; base < Lp; base += 16) {
            const uint4 a = *reinterpret_cast<const uint4 *>(c1 + base);
            const uint4 b = *reinterpret_cast<const uint4 *>(c2 + base);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t x = code_of(a, k), y = code_of(b, k);
                if (x != VFT_DEV_NOCODE && y != VFT_DEV_NOCODE) { nUse++; top = xadd(top, (double) s.distances[x * 20 + y]); }
            }
        }
    }
    weight = (P) (double) nUse;                                                           // :1621
    dist = (P) (nUse > 0 ? top / (double) nUse : 1.0);                                    // :1622
}

// distance half of setDistCriterion, NJ.tcc:1115-1122
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ void join_dist(const Store<P> &s, int64_t i, int64_t j, bool raw, P &dist, P &weight) {
    if (!raw && i < s.nSeqs && j < s.nSeqs) {
        seq_dist<P, MATRIX>(s, s.codes + i * s.Lp, s.codes + j * s.Lp, dist, weight);
    } else {
        const View<P, A> v1 = make_view<P, A>(s, i), v2 = make_view<P, A>(s, j);
        profile_dist<P, A, MATRIX>(s, v1, v2, dist, weight);
        if (raw) return;
        dist = psub(dist, padd(s.diameter[i], s.diameter[j]));                            // :1120
    }
    dist = (P) xadd((double) dist, 0.0);                 // :1122, constraintWeight * 0 (no constraints)
}

// setOutDistance, NJ.tcc:1012-1053
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ P out_distance(const Store<P> &s, int64_t iNode, int64_t nActive, double totdiam) {
    P ddist, dweight;
    const View<P, A> v1 = make_view<P, A>(s, iNode), vo = make_view<P, A>(s, -1);
    profile_dist<P, A, MATRIX>(s, v1, vo, ddist, dweight);
    const P pN = (P) nActive, pN1 = (P) (nActive - 1);
    const P t4 = psub(pmul(pmul(ddist, dweight), pN), pmul(s.selfweight[iNode], s.selfdist[iNode]));   // :1046
    const double top = (double) pmul(pN1, t4);
    const double bottom = (double) psub(pmul(dweight, pN), s.selfweight[iNode]);                        // :1047
    const double pd = top / bottom;                                                                       // :1048
    const P dn = pmul(s.diameter[iNode], pN1);
    const double r = bottom > 0.01 ? xsub(xsub(pd, (double) dn), xsub(totdiam, (double) s.diameter[iNode])) : 3.0;
    return (P) r;
}

// addToFreq / normalizeFreq on a register-resident frequency vector, NJ.tcc:821-871
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ void add_to_freq(const Store<P> &s, P (&f)[A], double weight, uint32_t codeIn, const P *fIn) {
    if (fIn != nullptr) {
        const P w = (P) weight;
#pragma unroll
        for (int k = 0; k < A; k++) f[k] = padd(f[k], pmul(fIn[k], w));
    } else if (MATRIX) {
        const P w = (P) weight;
        const P *cf = s.codeFreq + codeIn * 20;
#pragma unroll
        for (int k = 0; k < A; k++) f[k] = padd(f[k], pmul(cf[k], w));
    } else {
#pragma unroll
        for (int k = 0; k < A; k++) if ((uint32_t) k == codeIn) f[k] = (P) xadd((double) f[k], weight);   // :831
    }
}

template<typename P, int A, bool MATRIX>
__device__ __forceinline__ void normalize_freq(const Store<P> &s, P (&f)[A]) {
    double total = 0;
    if (MATRIX) {
        P e[A];
#pragma unroll
        for (int k = 0; k < A; k++) e[k] = s.eigentot[k];
        total = (double) vec_mul_sum<P, A>(f, e, s.reduction);                            // :849
    } else {
#pragma unroll
        for (int k = 0; k < A; k++) total = xadd(total, (double) f[k]);                   // :851-853
    }
    if (total > s.fPostTotalTolerance) {
        const P inv = (P) (1.0 / total);                                                  // :856
#pragma unroll
        for (int k = 0; k < A; k++) f[k] = pmul(f[k], inv);
    } else if (!MATRIX) {
#pragma unroll
        for (int k = 0; k < A; k++) f[k] = (P) (1.0 / A);
    } else {
#pragma unroll
        for (int k = 0; k < A; k++) f[k] = s.codeFreq[k];
    }
}

// setCodeDist for one position of the out-profile, NJ.tcc:873-898 (code1 = NOCODE, f = the vector)
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ void code_dist_row(const Store<P> &s, const P (&f)[A], P *cdRow) {
    if (!MATRIX) return;
    P e[A];
#pragma unroll
    for (int k = 0; k < A; k++) e[k] = s.eigenval[k];
    for (int c = 0; c < A; c++) {
        P b[A];
#pragma unroll
        for (int k = 0; k < A; k++) b[k] = s.codeFreq[c * 20 + k];
        cdRow[c] = (P) (double) vec_mul3_sum<P, A>(f, b, e, s.reduction);
    }
}


// profileDistPiece for the 4-state %different case with both vectors already in registers
template<typename P>
__device__ __forceinline__ P pick4(const P (&f)[4], uint32_t c) {
    return c == 0 ? f[0] : c == 1 ? f[1] : c == 2 ? f[2] : f[3];
}
template<typename P>
__device__ __forceinline__ double piece4(uint32_t c1, uint32_t c2, const P (&f1)[4], const P (&f2)[4]) {
    if (c1 != VFT_DEV_NOCODE) {                                                            // NJ.tcc:920-926
        if (c2 != VFT_DEV_NOCODE) return c1 == c2 ? 0.0 : 1.0;
        return xsub(1.0, (double) pick4(f2, c1));
    }
    if (c2 != VFT_DEV_NOCODE) return xsub(1.0, (double) pick4(f1, c2));                    // :928-930
    double pc = 1.0;                                                                       // :933-937
#pragma unroll
    for (int k = 0; k < 4; k++) pc = xsub(pc, (double) pmul(f1[k], f2[k]));
    return pc;
}

template<typename P> struct Vec4T;
template<> struct Vec4T<float> { typedef float4 type; };
template<> struct Vec4T<double> { typedef double4 type; };

// profileDist, NJ.tcc:1167-1190, by one WARP.
//   phase 1  the 32 lanes evaluate the positions in parallel (coalesced loads along the node-major
//            rows, 8 x 32 positions in flight at a time) and store the per-position terms
//            w1*w2 and w1*w2*piece to shared memory;
//   phase 2  lane 0 adds the `denom` terms and lane 1 the `top` terms IN POSITION ORDER.
// The additions are the reference's, in the reference's order; positions the reference skips
// contribute +0.0, and x + (+0.0) == x for every value the accumulators can take (they are never
// -0.0), so the result is bit-identical to the one-thread loop above.  Zero `top` terms (equal
// codes, the common case between close relatives) are compacted away before the ordered pass, and
// the `denom` pass is replaced by a tree sum whenever that is provably exact (see below).
// `sm` = 2*Lp doubles (+Lp/32 ints) of shared memory private to the calling warp: see warp_smem_bytes().
__host__ __device__ inline size_t warp_smem_bytes(int64_t Lp) { return (size_t) Lp * 16; }

template<typename P, int A, bool MATRIX>
__device__ __forceinline__ void profile_dist_warp(const Store<P> &s, const View<P, A> &p1, const View<P, A> &p2,
                                                  double *sm, P &dist, P &weight) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t Lp = s.Lp;
    double *termW = sm, *termT = sm + Lp;
    int nTop = 0;                                     // compacted count of non-zero top terms (uniform)
    constexpr int UC = 8;                             // chunks of 32 positions with their loads in flight together
    for (int64_t base = 0; base < Lp; base += 32 * UC) {
        uint32_t c1[UC], c2[UC];
        P w1[UC], w2[UC];
        constexpr bool PRE = (A == 4 && !MATRIX);     // 4-state vectors are prefetched with the weights
        P v1[PRE ? UC : 1][4], v2[PRE ? UC : 1][4];
#pragma unroll
        for (int u = 0; u < UC; u++) {
            const int64_t pos = base + 32 * u + lane;
            const bool in = pos < Lp;
            c1[u] = (in && p1.codes) ? (uint32_t) p1.codes[pos] : VFT_DEV_NOCODE;
            c2[u] = (in && p2.codes) ? (uint32_t) p2.codes[pos] : VFT_DEV_NOCODE;
            w1[u] = (in && p1.w) ? p1.w[pos] : (P) -1;
            w2[u] = (in && p2.w) ? p2.w[pos] : (P) -1;
            if (PRE) {
                typedef typename Vec4T<P>::type V4;
                if (in && p1.v) { const V4 t = *reinterpret_cast<const V4 *>(p1.v + pos * 4); v1[u][0] = t.x; v1[u][1] = t.y; v1[u][2] = t.z; v1[u][3] = t.w; }
                else { v1[u][0] = v1[u][1] = v1[u][2] = v1[u][3] = 0; }
                if (in && p2.v) { const V4 t = *reinterpret_cast<const V4 *>(p2.v + pos * 4); v2[u][0] = t.x; v2[u][1] = t.y; v2[u][2] = t.z; v2[u][3] = t.w; }
                else { v2[u][0] = v2[u][1] = v2[u][2] = v2[u][3] = 0; }
            }
        }
#pragma unroll
        for (int u = 0; u < UC; u++) {
            const int64_t pos = base + 32 * u + lane;
            if (base + 32 * u >= Lp) break;           // uniform
            const P a1 = p1.w ? w1[u] : (c1[u] != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
            const P a2 = p2.w ? w2[u] : (c2[u] != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
            double wt = 0, tt = 0;
            if (a1 > 0 && a2 > 0) {
                wt = (double) pmul(a1, a2);                                                // :1176
                double pc;
                if constexpr (PRE) pc = piece4<P>(c1[u], c2[u], v1[u], v2[u]);
                else pc = piece<P, A, MATRIX>(s, c1[u], c2[u], p1.v ? p1.v + pos * A : nullptr,
                                             p2.v ? p2.v + pos * A : nullptr, p2.cd ? p2.cd + pos * A : nullptr);
                tt = xmul(wt, pc);
            }
            termW[pos] = wt;
            const unsigned nz = __ballot_sync(full, tt != 0.0);
            if (tt != 0.0) termT[nTop + __popc(nz & ((1u << lane) - 1u))] = tt;
            nTop += __popc(nz);
        }
    }
    // `denom` fast path: when every partial sum of the w1*w2 terms is exactly representable (all
    // terms are multiples of 2^LB and the total stays below 2^(LB+53)), every addition in ANY order
    // is exact, so the ordered sum equals a tree sum.  True for the 0/1 and small dyadic weights
    // that NJ profiles carry in fp32; checked per pair, never assumed.
    double dLocal = 0;
    int minLB = 4096, maxE = -4096, cntNZ = 0;
    for (int64_t pos = lane; pos < Lp; pos += 32) {
        const double t = termW[pos];                  // own writes: visible without a barrier
        if (t > 0) {
            const unsigned long long bits = (unsigned long long) __double_as_longlong(t);
            const int e = (int) ((bits >> 52) & 0x7FF) - 1023;
            const unsigned long long mant = (bits & 0xFFFFFFFFFFFFFull) | 0x10000000000000ull;
            const int lb = e - 52 + (__ffsll((long long) mant) - 1);
            minLB = min(minLB, lb); maxE = max(maxE, e); cntNZ++;
            dLocal += t;                              // exactness of this partial sum is covered by the test below
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        minLB = min(minLB, __shfl_xor_sync(full, minLB, o));
        maxE = max(maxE, __shfl_xor_sync(full, maxE, o));
        cntNZ += __shfl_xor_sync(full, cntNZ, o);
    }
    int lg = 0;
    while ((1 << lg) < cntNZ) lg++;
    const bool exactDenom = cntNZ == 0 || (maxE + 1 + lg - minLB <= 53 && minLB > -1000);
    __syncwarp();
    double acc = 0;
    if (exactDenom) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dLocal += __shfl_xor_sync(full, dLocal, o);
        if (lane == 0) acc = dLocal;
    } else if (lane == 0) {
        for (int64_t k = 0; k < Lp; k += 4) {
            acc = xadd(acc, termW[k]); acc = xadd(acc, termW[k + 1]); acc = xadd(acc, termW[k + 2]); acc = xadd(acc, termW[k + 3]);
        }
    }
    if (lane == 1) {
        for (int k = 0; k < nTop; k++) acc = xadd(acc, termT[k]);
    }
    __syncwarp();
    const double denom = __shfl_sync(full, acc, 0), top = __shfl_sync(full, acc, 1);
    weight = (P) (denom > 0 ? denom : 0.01);                                              // :1187
    dist = (P) (denom > 0 ? top / denom : 1.0);                                           // :1188
}

// distance half of setDistCriterion for a pair that is NOT leaf x leaf (or raw), by one warp
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ void join_dist_warp(const Store<P> &s, int64_t i, int64_t j, bool raw, double *sm,
                                               P &dist, P &weight) {
    const View<P, A> v1 = make_view<P, A>(s, i), v2 = make_view<P, A>(s, j);
    profile_dist_warp<P, A, MATRIX>(s, v1, v2, sm, dist, weight);
    if (raw) return;
    dist = psub(dist, padd(s.diameter[i], s.diameter[j]));                                // :1120
    dist = (P) xadd((double) dist, 0.0);                                                  // :1122
}

// setOutDistance, NJ.tcc:1012-1053, by one warp
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ P out_distance_warp(const Store<P> &s, int64_t iNode, int64_t nActive, double totdiam, double *sm) {
    P ddist, dweight;
    const View<P, A> v1 = make_view<P, A>(s, iNode), vo = make_view<P, A>(s, -1);
    profile_dist_warp<P, A, MATRIX>(s, v1, vo, sm, ddist, dweight);
    const P pN = (P) nActive, pN1 = (P) (nActive - 1);
    const P t4 = psub(pmul(pmul(ddist, dweight), pN), pmul(s.selfweight[iNode], s.selfdist[iNode]));   // :1046
    const double top = (double) pmul(pN1, t4);
    const double bottom = (double) psub(pmul(dweight, pN), s.selfweight[iNode]);                        // :1047
    const double pd = top / bottom;                                                                       // :1048
    const P dn = pmul(s.diameter[iNode], pN1);
    const double r = bottom > 0.01 ? xsub(xsub(pd, (double) dn), xsub(totdiam, (double) s.diameter[iNode])) : 3.0;
    return (P) r;
}

// orderable keys: ascending unsigned order == ascending floating order
__device__ __forceinline__ uint64_t order_key(float x) {
    uint32_t u = __float_as_uint(x);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (uint64_t) u;
}
__device__ __forceinline__ uint64_t order_key(double x) {
    uint64_t u = (uint64_t) __double_as_longlong(x);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

}  // namespace vft
