// vft_device.cuh -- device-side arithmetic of the NJ hot path (sm_100a).
//
// Everything here is "reference-order arithmetic": each value is produced by the same sequence
// of individually rounded operations, in the same types, as the reference's -mavx2 (no FMA)
// build evaluates it (citations: /root/reference/src, NJ.tcc = NeighbourJoining.tcc).  That is
// what makes the top-hit indices identical to the reference rather than merely close: the
// criterion has many exact and near ties, and the order of the double-precision accumulation
// over positions decides them.  Consequences for the kernel design:
//   - a pair's distance is accumulated by ONE lane, positions in ascending order; parallelism
//     comes from running the chains of many pairs side by side (group_profile_dist);
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
    // rows of ONE node staged in shared memory by the CTA (vft_bulk.cuh): the node every pair of the sweep shares
    int32_t ovId;          // -2: none
    const uint8_t *ovCodes; const P *ovW, *ovV;
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

// profileDist, NJ.tcc:1167-1190, for a GROUP of pairs by one WARP ("tile-transposed" ordered
// accumulation).  The reference adds the per-position terms of a pair in position order into two
// doubles; the order decides the last bit of top/denom and with it the exact and near ties of the NJ
// criterion, so it is reproduced: a pair's terms are still added one after the other, by ONE lane --
// but up to R pairs are in flight per warp instead of one:
//   phase 1  for every pair (row) of the group, the 32 lanes evaluate C = 32*PPL consecutive positions
//            (PPL positions per lane, 128-bit loads along the node-major rows) and store the two terms
//            w1*w2 and w1*w2*piece into that row of a padded [R][C] shared-memory tile;
//   phase 2  lane k adds row k, left to right, into ITS pair's denom / top accumulators: R
//            independent ordered chains (denom and top interleaved: one DADD latency per position).
// Positions the reference skips contribute +0.0, and x + (+0.0) == x for every value the accumulators
// can take (they start at +0.0 and never become -0.0), so the sums are bit-identical to the
// one-thread loop in profile_dist().  Loads are software-pipelined across units (unit = one pair x C
// positions), also across chunk boundaries, so a single pair (the per-join lists) streams its rows
// without a dependent-load stall per chunk.  Consecutive pairs that share their first node (every
// list the NJ driver produces) reuse its registers.
// Tile shapes: 4-state %different profiles (nt): PPL=4, R=8, C=128; 20-state / matrix: PPL=1, R=32, C=32.
template<int A, bool MATRIX> struct TileShape {
    static constexpr int PPL = (A == 4 && !MATRIX) ? 4 : 1;
    static constexpr int R = 32 / PPL, C = 32 * PPL;
};
template<typename P, int A, bool MATRIX>
__host__ __device__ inline size_t group_smem_bytes(int rows) {      // per warp, independent of the alignment length
    constexpr int C = TileShape<A, MATRIX>::C;
    constexpr int WS = sizeof(P) == 4 ? C + 4 : C + 2;
    return (size_t) rows * ((C + 2) * 8 + WS * sizeof(P));
}

template<typename P> struct VecLd;
template<> struct VecLd<float>  { typedef float4 type;  static constexpr int N = 4; };
template<> struct VecLd<double> { typedef double2 type; static constexpr int N = 2; };

// n consecutive P's from a 16-byte aligned address with 128-bit loads
template<typename P, int N_>
__device__ __forceinline__ void load_vec(const P *__restrict__ src, P *v) {
    typedef typename VecLd<P>::type V;
    constexpr int N = VecLd<P>::N;
    if constexpr (N_ % N == 0) {
        const V *q = reinterpret_cast<const V *>(src);
#pragma unroll
        for (int k = 0; k < N_ / N; k++) {
            const V t = q[k];
            if constexpr (N == 4) { v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w; }
            else { v[2 * k] = t.x; v[2 * k + 1] = t.y; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < N_; k++) v[k] = src[k];
    }
}

// n consecutive P's to a 16-byte aligned address with 128-bit stores
template<typename P, int N_>
__device__ __forceinline__ void store_vec(P *__restrict__ dst, const P *v) {
    typedef typename VecLd<P>::type V;
    constexpr int N = VecLd<P>::N;
    if constexpr (N_ % N == 0) {
        V *q = reinterpret_cast<V *>(dst);
#pragma unroll
        for (int k = 0; k < N_ / N; k++) {
            V t;
            if constexpr (N == 4) { t.x = v[4 * k]; t.y = v[4 * k + 1]; t.z = v[4 * k + 2]; t.w = v[4 * k + 3]; }
            else { t.x = v[2 * k]; t.y = v[2 * k + 1]; }
            q[k] = t;
        }
    } else {
#pragma unroll
        for (int k = 0; k < N_; k++) dst[k] = v[k];
    }
}

// WIDE variant (one pair per CTA, the mid-sized lists of long alignments): the calling warp evaluates the chunks
// chunk0, chunk0+chunkStep, ... of ONE pair into a full-length term row (T, W: [Lp rounded up to C] each) and does
// not accumulate; the caller adds the row in order once every warp of the CTA is done (cta_ordered_sum).
template<typename P, int A, bool MATRIX, bool WIDE = false>
__device__ __forceinline__ void group_profile_dist(const Store<P> &s, int64_t myA, int64_t myB, unsigned mask, int rows,
                                                   unsigned char *smw, double &denomOut, double &topOut,
                                                   int chunk0 = 0, int chunkStep = 1) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    constexpr int PPL = TileShape<A, MATRIX>::PPL, C = TileShape<A, MATRIX>::C;
    constexpr bool REGV = (MATRIX || A == 4);          // vectors staged in registers (the generic !MATRIX A=20 case reads them in place)
    constexpr bool VCOND = MATRIX;                     // matrix mode: vectors (or codeFreq rows / table entries) are fetched once codes and weights are known
    constexpr int TS = C + 2;                          // row strides: 16-byte aligned rows, conflict-free 128-bit row reads
    constexpr int WS = sizeof(P) == 4 ? C + 4 : C + 2;
    const int nChunksAll = (int) ((s.Lp + C - 1) / C);
    double *T = reinterpret_cast<double *>(smw);       // [rows][TS]  (WIDE: [nChunksAll*C])
    P *W = reinterpret_cast<P *>(smw + (WIDE ? (size_t) nChunksAll * C * 8 : (size_t) rows * TS * 8));   // [rows][WS]; w1*w2 is a P product: exact in P
    const int nItems = __popc(mask);                   // <= rows <= R (caller's contract)
    denomOut = 0; topOut = 0;
    if (nItems == 0) return;
    // compact the group: lane k takes the k-th item (its row), and will accumulate it
    const int srcLane = lane < nItems ? __fns(mask, 0, lane + 1) : 0;
    const int rowA = (int) __shfl_sync(full, myA, srcLane), rowB = (int) __shfl_sync(full, myB, srcLane);
    const uint32_t Lp = (uint32_t) s.Lp, nSeqs = (uint32_t) s.nSeqs;
    const int nChunks = WIDE ? (nChunksAll - chunk0 + chunkStep - 1) / chunkStep : nChunksAll;   // chunks this warp evaluates
    P eig[MATRIX ? A : 1];
    if constexpr (MATRIX) {
#pragma unroll
        for (int k = 0; k < A; k++) eig[k] = s.eigenval[k];
    }

    // one side of a unit: where its codes / weights / vectors live (nullptr: leaf or out-profile conventions of View)
    struct Side { const uint8_t *codes; const P *w; const P *v; };
    auto side = [&](int id) {
        Side r;
        if (id < 0) { r.codes = nullptr; r.w = s.ow; r.v = s.ov; }
        else if (id == s.ovId) { r.codes = s.ovCodes; r.w = s.ovW; r.v = s.ovV; }
        else {
            r.codes = s.codes + (uint64_t) (uint32_t) id * Lp;
            if ((uint32_t) id < nSeqs) { r.w = nullptr; r.v = nullptr; }
            else {
                const uint64_t off = (uint64_t) ((uint32_t) id - nSeqs) * Lp;
                r.w = s.weights + off; r.v = s.vecs + off * A;
            }
        }
        return r;
    };
    struct CW { uint32_t c1[PPL], c2[PPL]; P w1[PPL], w2[PPL]; };
    struct VV { P v1[REGV ? PPL * A : 1], v2[REGV ? PPL * A : 1]; P pre[VCOND ? PPL : 1]; };
    struct Unit { int k, c, na, nb; uint32_t pos; bool in; };
    auto unitAt = [&](int k, int c) {
        Unit u;
        u.k = k; u.c = c;
        u.na = __shfl_sync(full, rowA, k); u.nb = __shfl_sync(full, rowB, k);
        u.pos = (uint32_t) (WIDE ? chunk0 + c * chunkStep : c) * C + lane * PPL;
        u.in = u.pos < Lp;                             // Lp is a multiple of 32 >= PPL: a lane's PPL positions are all in or all out
        return u;
    };
    auto nextUnit = [&](const Unit &u) { return (u.k + 1 < nItems) ? unitAt(u.k + 1, u.c) : unitAt(0, u.c + 1); };
    auto loadCodes = [&](const uint8_t *row, uint32_t pos, uint32_t (&c)[PPL]) {
        if (row == nullptr) {
#pragma unroll
            for (int i = 0; i < PPL; i++) c[i] = VFT_DEV_NOCODE;
        } else if constexpr (PPL == 4) {
            const uint32_t q = *reinterpret_cast<const uint32_t *>(row + pos);
#pragma unroll
            for (int i = 0; i < 4; i++) c[i] = (q >> (8 * i)) & 0xFFu;
        } else {
#pragma unroll
            for (int i = 0; i < PPL; i++) c[i] = (uint32_t) row[pos + i];
        }
    };
    // stage A: codes + weights; in %different mode also the (narrow) vectors, unconditionally
    auto loadA = [&](const Unit &u, CW &o, VV &q) {
        if (!u.in) {
#pragma unroll
            for (int i = 0; i < PPL; i++) { o.c1[i] = o.c2[i] = VFT_DEV_NOCODE; o.w1[i] = o.w2[i] = 0; }
            return;
        }
        const Side s1 = side(u.na), s2 = side(u.nb);
        loadCodes(s1.codes, u.pos, o.c1);
        loadCodes(s2.codes, u.pos, o.c2);
        if (s1.w) load_vec<P, PPL>(s1.w + u.pos, o.w1);
        else {
#pragma unroll
            for (int i = 0; i < PPL; i++) o.w1[i] = o.c1[i] != VFT_DEV_NOCODE ? (P) 1 : (P) 0;
        }
        if (s2.w) load_vec<P, PPL>(s2.w + u.pos, o.w2);
        else {
#pragma unroll
            for (int i = 0; i < PPL; i++) o.w2[i] = o.c2[i] != VFT_DEV_NOCODE ? (P) 1 : (P) 0;
        }
        if constexpr (REGV && !VCOND) {
            if (s1.v) load_vec<P, PPL * A>(s1.v + (uint64_t) u.pos * A, q.v1);
            else {
#pragma unroll
                for (int k = 0; k < PPL * A; k++) q.v1[k] = 0;
            }
            if (s2.v) load_vec<P, PPL * A>(s2.v + (uint64_t) u.pos * A, q.v2);
            else {
#pragma unroll
                for (int k = 0; k < PPL * A; k++) q.v2[k] = 0;
            }
        }
    };
    // stage B (matrix mode), once codes and weights are known: per position either the table entry that
    // IS the piece (code x code, or code x codeDist of the out-profile) or the two vectors of the 3-way
    // dot product -- the profile's own vector, or the codeFreq row that stands in for a known code
    // (profileDistPiece, NJ.tcc:900-916)
    auto loadB = [&](const Unit &u, const CW &cw, VV &q) {
        if constexpr (VCOND) {
            if (!u.in) return;
            const Side s1 = side(u.na), s2 = side(u.nb);
#pragma unroll
            for (int i = 0; i < PPL; i++) {
                const uint32_t c1 = cw.c1[i], c2 = cw.c2[i];
                q.pre[i] = 0;
                if (cw.w1[i] > 0 && cw.w2[i] > 0) {
                    if (c1 != VFT_DEV_NOCODE && c2 != VFT_DEV_NOCODE) q.pre[i] = s.distances[c1 * 20 + c2];
                    else if (u.nb < 0 && c1 != VFT_DEV_NOCODE) q.pre[i] = s.ocd[(uint64_t) (u.pos + i) * A + c1];
                    else {
                        const P *f1 = c1 != VFT_DEV_NOCODE ? s.codeFreq + c1 * 20 : s1.v + (uint64_t) (u.pos + i) * A;
                        const P *f2 = c2 != VFT_DEV_NOCODE ? s.codeFreq + c2 * 20 : s2.v + (uint64_t) (u.pos + i) * A;
                        load_vec<P, A>(f1, q.v1 + i * A);
                        load_vec<P, A>(f2, q.v2 + i * A);
                    }
                }
            }
        }
    };

    double den = 0, top = 0;
    auto compute = [&](const Unit &u, const CW &cw, const VV &q) {
        double tt[PPL];
        P wt[PPL];
#pragma unroll
        for (int i = 0; i < PPL; i++) {
            const uint32_t c1 = cw.c1[i], c2 = cw.c2[i];
            const bool on = cw.w1[i] > 0 && cw.w2[i] > 0;
            wt[i] = on ? pmul(cw.w1[i], cw.w2[i]) : (P) 0;                               // :1176
            double pc;
            if constexpr (MATRIX) {
                const bool table = (c1 != VFT_DEV_NOCODE && c2 != VFT_DEV_NOCODE) || (u.nb < 0 && c1 != VFT_DEV_NOCODE);
                pc = (double) q.pre[i];
                if (on && !table) pc = (double) vec_mul3_sum<P, A>(q.v1 + i * A, q.v2 + i * A, eig, s.reduction);
            } else if constexpr (A == 4) {
                // profileDistPiece's four %different cases (NJ.tcc:918-939) in one branch-free form: a known code
                // acts as the indicator vector of that code; 1*f and 0*f are exact, and subtracting +0.0 is exact,
                // so "1 - f2[c1]", "c1==c2 ? 0 : 1" and the full "1 - sum f1*f2" all come out of the same chain
                pc = 1.0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const P g1 = c1 != VFT_DEV_NOCODE ? (c1 == (uint32_t) k ? (P) 1 : (P) 0) : q.v1[i * 4 + k];
                    const P g2 = c2 != VFT_DEV_NOCODE ? (c2 == (uint32_t) k ? (P) 1 : (P) 0) : q.v2[i * 4 + k];
                    pc = xsub(pc, (double) pmul(g1, g2));
                }
            } else {
                pc = 0;
                if (on) {
                    const View<P, A> v1 = make_view<P, A>(s, u.na), v2 = make_view<P, A>(s, u.nb);
                    pc = piece<P, A, false>(s, c1, c2, v1.v ? v1.v + (uint64_t) (u.pos + i) * A : nullptr,
                                            v2.v ? v2.v + (uint64_t) (u.pos + i) * A : nullptr, nullptr);
                }
            }
            tt[i] = on ? xmul((double) wt[i], pc) : 0.0;
        }
        double *tr = WIDE ? T + u.pos : T + u.k * TS + lane * PPL;
        P *wr = WIDE ? W + u.pos : W + u.k * WS + lane * PPL;
        if constexpr (PPL == 4) {
            *reinterpret_cast<double2 *>(tr) = make_double2(tt[0], tt[1]);
            *reinterpret_cast<double2 *>(tr + 2) = make_double2(tt[2], tt[3]);
            if constexpr (sizeof(P) == 4) *reinterpret_cast<float4 *>(wr) = make_float4(wt[0], wt[1], wt[2], wt[3]);
            else { *reinterpret_cast<double2 *>(wr) = make_double2(wt[0], wt[1]); *reinterpret_cast<double2 *>(wr + 2) = make_double2(wt[2], wt[3]); }
        } else {
#pragma unroll
            for (int i = 0; i < PPL; i++) { tr[i] = tt[i]; wr[i] = wt[i]; }
        }
        // end of a chunk: every row holds its C terms -> ordered accumulation, one lane per pair
        if (!WIDE && u.k + 1 == nItems) {
            __syncwarp();
            if (lane < nItems) {
                const double *trow = T + lane * TS;
                const P *wrow = W + lane * WS;
#pragma unroll
                for (int j = 0; j < C; j += 4) {
                    const double2 ta = *reinterpret_cast<const double2 *>(trow + j), tb = *reinterpret_cast<const double2 *>(trow + j + 2);
                    P w4[4];
                    load_vec<P, 4>(wrow + j, w4);
                    den = xadd(den, (double) w4[0]); top = xadd(top, ta.x);
                    den = xadd(den, (double) w4[1]); top = xadd(top, ta.y);
                    den = xadd(den, (double) w4[2]); top = xadd(top, tb.x);
                    den = xadd(den, (double) w4[3]); top = xadd(top, tb.y);
                }
            }
            __syncwarp();
        }
    };

    const int U = nChunks * nItems;
    if (U <= 0) return;
    // Two register buffers (X, Y) alternate between "being computed" and "being loaded": the loop is
    // unrolled by two so that no buffer is ever copied.  Matrix mode adds a second, earlier stage for the
    // codes + weights (a few registers, rotated).
    Unit uX = unitAt(0, 0), uY = uX, uZ = uX;
    CW cwX, cwY, cwZ;
    VV qX, qY;
    loadA(uX, cwX, qX);
    if (U > 1) { uY = nextUnit(uX); loadA(uY, cwY, qY); }
    loadB(uX, cwX, qX);
    for (int u = 0; u < U; u += 2) {
        // ---- even unit: compute X; Y is one ahead ---------------------------------------------------
        if constexpr (VCOND) {
            if (u + 2 < U) { uZ = nextUnit(uY); loadA(uZ, cwZ, qX /*unused*/); }
            if (u + 1 < U) loadB(uY, cwY, qY);
        }
        compute(uX, cwX, qX);
        if constexpr (VCOND) { uX = uZ; cwX = cwZ; }                       // X now describes unit u+2 (codes + weights loaded)
        else if (u + 2 < U) { uX = nextUnit(uY); loadA(uX, cwX, qX); }
        if (u + 1 >= U) break;
        // ---- odd unit: compute Y; X is one ahead ----------------------------------------------------
        if constexpr (VCOND) {
            if (u + 3 < U) { uZ = nextUnit(uX); loadA(uZ, cwZ, qY /*unused*/); }
            if (u + 2 < U) loadB(uX, cwX, qX);
        }
        compute(uY, cwY, qY);
        if constexpr (VCOND) { uY = uZ; cwY = cwZ; }
        else if (u + 3 < U) { uY = nextUnit(uX); loadA(uY, cwY, qY); }
    }
    if (WIDE) return;
    // hand each item's sums back to the lane that owns it
    const int rank = __popc(mask & ((1u << lane) - 1u));
    denomOut = __shfl_sync(full, den, rank);
    topOut = __shfl_sync(full, top, rank);
}

// shared memory of the WIDE variant: one full-length term row per CTA
template<typename P, int A, bool MATRIX>
__host__ __device__ inline size_t wide_smem_bytes(int64_t Lp) {
    constexpr int C = TileShape<A, MATRIX>::C;
    const size_t cols = (size_t) ((Lp + C - 1) / C) * C;
    return cols * (8 + sizeof(P));
}

// the ordered accumulation of a full-length term row by ONE lane (call with a full warp; lane 0 returns the sums)
template<typename P, int A, bool MATRIX>
__device__ __forceinline__ void cta_ordered_sum(const Store<P> &s, const unsigned char *smw, double &denomOut, double &topOut) {
    constexpr int C = TileShape<A, MATRIX>::C;
    const int cols = (int) ((s.Lp + C - 1) / C) * C;
    const double *T = reinterpret_cast<const double *>(smw);
    const P *W = reinterpret_cast<const P *>(smw + (size_t) cols * 8);
    double den = 0, top = 0;
    if ((threadIdx.x & 31) == 0) {
        for (int j = 0; j < cols; j += 8) {
            double t8[8]; P w8[8];
#pragma unroll
            for (int q = 0; q < 8; q += 2) { const double2 t = *reinterpret_cast<const double2 *>(T + j + q); t8[q] = t.x; t8[q + 1] = t.y; }
            load_vec<P, 8>(W + j, w8);
#pragma unroll
            for (int q = 0; q < 8; q++) { den = xadd(den, (double) w8[q]); top = xadd(top, t8[q]); }
        }
    }
    denomOut = den; topOut = top;
}

// profileDist's tail, NJ.tcc:1187-1188
template<typename P>
__device__ __forceinline__ void finish_dist(double denom, double top, P &dist, P &weight) {
    weight = (P) (denom > 0 ? denom : 0.01);
    dist = (P) (denom > 0 ? top / denom : 1.0);
}

// distance half of setDistCriterion (NJ.tcc:1115-1122) applied to a finished profileDist
template<typename P>
__device__ __forceinline__ P join_correct(const Store<P> &s, int64_t i, int64_t j, P dist) {
    dist = psub(dist, padd(s.diameter[i], s.diameter[j]));                                // :1120
    return (P) xadd((double) dist, 0.0);                                                  // :1122
}

// setOutDistance's algebra, NJ.tcc:1046-1052, applied to profileDist(node, out-profile)
template<typename P>
__device__ __forceinline__ P out_distance_finish(const Store<P> &s, int64_t iNode, int64_t nActive, double totdiam, P ddist, P dweight) {
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
    uint32_t u = __float_as_uint(__fadd_rn(x, 0.0f));          // -0.0 -> +0.0: the reference compares values, not bits
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (uint64_t) u;
}
__device__ __forceinline__ uint64_t order_key(double x) {
    uint64_t u = (uint64_t) __double_as_longlong(__dadd_rn(x, 0.0));
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

}  // namespace vft
