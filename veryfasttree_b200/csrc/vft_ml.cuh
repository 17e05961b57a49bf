// vft_ml.cuh -- device-side likelihood arithmetic: expEigenRates / fastexp, pairLogLk, posteriorProfile.
// Reference-order arithmetic as in vft_device.cuh (citations: /root/reference/src, NJ.tcc =
// NeighbourJoining.tcc).  The only operations that are not bit-reproducible against the CPU are the
// libm calls the reference itself makes (exp at -fastexp 0/1 and in the Jukes-Cantor pSame, the
// final log of pairLogLk): CUDA's exp/log are within 1 ulp of glibc's, so those outputs are held
// to the north_star tolerance (1e-5 relative fp32, 1e-12 fp64) instead of bit equality.
#pragma once
#include "vft_device.cuh"

namespace vft {

template<typename P>
struct MLModel {
    const P *codeFreq;      // [(A+1)][A], last row = NOCODE (gap) row; nullptr => Jukes-Cantor
    const P *eigenval;      // [A]
    const P *eigeninv;      // [A][A]
    const P *eigeninvT;     // [A][A] (nt)
    const P *statinv;       // [A]
    const P *rates;         // [nRateCats]
    const int32_t *ratecat; // [Lp]
    int nRateCats, fastexp;
    double MLMinRelBranchLength, MLMinBranchLength;
};

// fastexp, BasicOperations.tcc:121-216 (levels 2/3: the Cephes-style rational, every operation rounded
// separately as on the CPU)
template<typename P>
__device__ __forceinline__ P fastexp1(P v, int lvl) {
    if (lvl == 0) return (P) exp((double) v);
    if (lvl == 1) return (P) expf((float) v);
    if (lvl == 2) {
        double x = (double) v;
        double px = floor(xadd(xmul(1.4426950408889634073599, x), 0.5));
        const long long m = (long long) px;
        x = xsub(x, xmul(px, 6.93145751953125E-1));
        x = xsub(x, xmul(px, 1.42860682030941723212E-6));
        const double xx = xmul(x, x);
        px = 1.26177193074810590878E-4; px = xmul(px, xx); px = xadd(px, 3.02994407707441961300E-2);
        px = xmul(px, xx); px = xadd(px, 9.99999999999999999910E-1); px = xmul(px, x);
        double qx = 3.00198505138664455042E-6; qx = xmul(qx, xx); qx = xadd(qx, 2.52448340349684104192E-3);
        qx = xmul(qx, xx); qx = xadd(qx, 2.27265548208155028766E-1); qx = xmul(qx, xx); qx = xadd(qx, 2.00000000000000000009E0);
        x = px / xsub(qx, px);
        x = xadd(1.0, xmul(2.0, x));
        const double scale = __longlong_as_double((m + 1023ll) << 52);
        return (P) xmul((double) (P) x, scale);
    }
    float x = (float) v;
    float px = floorf(__fadd_rn(__fmul_rn(1.4426950408889634073599f, x), 0.5f));
    const int m = (int) px;
    x = __fsub_rn(x, __fmul_rn(px, 6.93145751953125E-1f));
    x = __fsub_rn(x, __fmul_rn(px, 1.42860682030941723212E-6f));
    const float xx = __fmul_rn(x, x);
    px = 1.26177193074810590878E-4f; px = __fmul_rn(px, xx); px = __fadd_rn(px, 3.02994407707441961300E-2f);
    px = __fmul_rn(px, xx); px = __fadd_rn(px, 9.99999999999999999910E-1f); px = __fmul_rn(px, x);
    float qx = 3.00198505138664455042E-6f; qx = __fmul_rn(qx, xx); qx = __fadd_rn(qx, 2.52448340349684104192E-3f);
    qx = __fmul_rn(qx, xx); qx = __fadd_rn(qx, 2.27265548208155028766E-1f); qx = __fmul_rn(qx, xx); qx = __fadd_rn(qx, 2.00000000000000000009E0f);
    x = __fdiv_rn(px, __fsub_rn(qx, px));
    x = (float) xadd(1.0, xmul(2.0, (double) x));
    const float scale = __int_as_float((m + 127) << 23);
    return pmul((P) x, (P) scale);
}

// expEigenRates, NJ.tcc:2020-2038: table[iRate*A + j], filled cooperatively by `nThreads` threads
template<typename P, int A>
__device__ __forceinline__ void exp_eigen_rates(const MLModel<P> &m, double length, P *table, int tid, int nThreads) {
    for (int idx = tid; idx < m.nRateCats * A; idx += nThreads) {
        const int r = idx / A, j = idx - r * A;
        double relLen = xmul(length, (double) m.rates[r]);
        if (relLen < m.MLMinRelBranchLength) relLen = m.MLMinRelBranchLength;
        table[idx] = fastexp1<P>(pmul(m.eigenval[j], (P) relLen), m.fastexp);
    }
}

// pSameVector / pDiffVector, NJ.tcc:2005-2018
template<typename P>
__device__ __forceinline__ void jc_tables(const MLModel<P> &m, double length, double *pSame, double *pDiff, int tid, int nThreads) {
    for (int r = tid; r < m.nRateCats; r += nThreads) {
        const double ps = xadd(0.25, xmul(0.75, exp(xmul(-4.0 / 3.0, fabs(xmul(length, (double) m.rates[r]))))));
        pSame[r] = ps;
        pDiff[r] = xsub(1.0, ps) / 3.0;
    }
}

// f = vector, or codeFreq[code] (gap row for NOCODE), mixed with the gap row when 0 < w < 1
// (NJ.tcc:1283-1302 / :2281-2300)
template<typename P, int A>
__device__ __forceinline__ void ml_freq(const MLModel<P> &m, uint32_t code, P w, const P *vec, P (&f)[A]) {
    const bool hasVec = w > 0 && code == VFT_DEV_NOCODE && vec != nullptr;
    const P *src = hasVec ? vec : m.codeFreq + (code == VFT_DEV_NOCODE ? A : (int) code) * A;
#pragma unroll
    for (int j = 0; j < A; j++) f[j] = src[j];
    const double wd = (double) w;
    if (wd > 0.0 && wd < 1.0) {
        const P *g = m.codeFreq + A * A;
#pragma unroll
        for (int j = 0; j < A; j++) f[j] = (P) xadd(xmul(wd, (double) f[j]), xmul(xsub(1.0, wd), (double) g[j]));
    }
}

// per-position lkAB of pairLogLk, NJ.tcc:1212-1257 (JC) / :1273-1309 (nt matrix) / :1329-1359 (aa);
// returns 1.0 for the positions the reference skips (lk *= 1.0 is exact)
template<typename P, int A>
__device__ __forceinline__ double site_lk(const Store<P> &s, const MLModel<P> &m, const View<P, A> &p1, const View<P, A> &p2,
                                          int64_t pos, const P *expeigen, const double *pSame, const double *pDiff) {
    const uint32_t cA = p1.codes[pos], cB = p2.codes[pos];
    const P wAp = p1.w ? p1.w[pos] : (cA != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
    const P wBp = p2.w ? p2.w[pos] : (cB != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
    const int r = m.ratecat[pos];
    if (m.codeFreq == nullptr) {                                                           // Jukes-Cantor (A == 4)
        const double wA = (double) wAp, wB = (double) wBp;
        const bool vA = wAp > 0 && cA == VFT_DEV_NOCODE && p1.v, vB = wBp > 0 && cB == VFT_DEV_NOCODE && p2.v;
        const P *fA = p1.v + pos * A, *fB = p2.v + pos * A;
        if (!vA && !vB) {
            if (cA == VFT_DEV_NOCODE || cB == VFT_DEV_NOCODE) return 0.25;
            const double ww = xmul(wA, wB);
            const double p = cA == cB ? pSame[r] : pDiff[r];
            return xadd(xmul(xmul(p, wA), wB), xmul(0.25, xsub(1.0, ww)));                 // :1231 / :1233
        }
        if (!vA) {
            if (cA == VFT_DEV_NOCODE) return 0.25;
            return xadd(xmul(wA, xadd(pDiff[r], xmul((double) fB[cA], xsub(pSame[r], pDiff[r])))), xmul(xsub(1.0, wA), 0.25));   // :1240
        }
        if (!vB) {
            if (cB == VFT_DEV_NOCODE) return 0.25;
            return xadd(xmul(wB, xadd(pDiff[r], xmul((double) fA[cB], xsub(pSame[r], pDiff[r])))), xmul(xsub(1.0, wB), 0.25));   // :1250
        }
        double lk = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)                                                        // :1253-1255
            lk = xadd(lk, xmul((double) fB[j], xadd(xmul((double) fA[j], pSame[r]), xmul((double) psub((P) 1, fA[j]), pDiff[r]))));
        return lk;
    }
    if (wAp == 0 && wBp == 0 && cA == VFT_DEV_NOCODE && cB == VFT_DEV_NOCODE) return 1.0; // :1278 / :1334
    P fA[A], fB[A], ee[A];
    ml_freq<P, A>(m, cA, wAp, p1.v ? p1.v + pos * A : nullptr, fA);
    ml_freq<P, A>(m, cB, wBp, p2.v ? p2.v + pos * A : nullptr, fB);
#pragma unroll
    for (int j = 0; j < A; j++) ee[j] = expeigen[r * A + j];
    if (A == 4) {
        double lk = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) lk = xadd(lk, (double) pmul(pmul(ee[j], fA[j]), fB[j]));    // :1306-1309
        return lk;
    }
    return (double) vec_mul3_sum<P, A>(ee, fA, fB, s.reduction);                           // :1359
}

// pairLogLk, NJ.tcc:1192-1447, by the nT threads (1, 2, 4 or 8 warps) that share one (pair, length) item; every thread of
// the CTA calls this (the barriers are CTA-wide), `valid` = the group has an item.  sm: [Lp] doubles (per-site lkAB), then
// the group's exp(eigenvalue x rate x length) table.  The per-site values are computed by all threads; the product over
// the sites is the reference's own sequential chain (double products with the 1e-4 / 1e4 rescale, whose rounding depends
// on the order), run by the group's first thread over register-prefetched blocks of 8 terms.
template<typename P, int A>
__device__ __forceinline__ void pair_loglk_group(const Store<P> &s, const MLModel<P> &m, bool valid, int64_t i, int64_t j, double length,
                                                 double *termL, void *tableRaw, double *siteOut, int tid, int nT, double *out) {
    P *expeigen = reinterpret_cast<P *>(tableRaw);
    double *pSame = reinterpret_cast<double *>(tableRaw), *pDiff = pSame + m.nRateCats;
    if (valid) {
        if (m.codeFreq) exp_eigen_rates<P, A>(m, length, expeigen, tid, nT);
        else jc_tables<P>(m, length, pSame, pDiff, tid, nT);
    }
    __syncthreads();
    if (valid) {
        const View<P, A> p1 = make_view<P, A>(s, i), p2 = make_view<P, A>(s, j);
        for (int64_t pos = tid; pos < s.Lp; pos += nT) {
            const double v = pos < s.L ? site_lk<P, A>(s, m, p1, p2, pos, expeigen, pSame, pDiff) : 1.0;
            termL[pos] = v;
            if (siteOut && pos < s.L) siteOut[pos] = v;
        }
    }
    __syncthreads();
    if (valid && tid == 0) {
        const double LkUnderflow = 1.0e-4, LkUnderflowInv = 1.0e4, LogLkUnderflow = 9.21034037197618;   // Constants.h:13-15
        double lk = 1.0, loglk = 0.0;
        const bool up = m.codeFreq != nullptr;          // the JC branch has no upward rescale (:1259-1262)
        const double2 *t2 = reinterpret_cast<const double2 *>(termL);
        // (Two branch-free variants of this chain -- the rescale as a select, and as predicated multiplies decided on the
        // high word -- were measured SLOWER on the B200 than this literal form with one guarding test per site: 177 / 194 ms
        // against 123 ms per sweep of profiles/ml_opt.py.)
        for (int64_t p0 = 0; p0 < s.Lp; p0 += 8) {      // Lp is a multiple of 32; the padding holds 1.0 (exact, no rescale)
            double t[8];
#pragma unroll
            for (int q = 0; q < 4; q++) { const double2 x = t2[(p0 >> 1) + q]; t[2 * q] = x.x; t[2 * q + 1] = x.y; }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                lk = xmul(lk, t[q]);
                if (lk < LkUnderflow || (up && lk > LkUnderflowInv)) {
                    while (lk < LkUnderflow && lk > 0.0) { lk = xmul(lk, LkUnderflowInv); loglk = xsub(loglk, LogLkUnderflow); }
                    if (up) while (lk > LkUnderflowInv) { lk = xmul(lk, LkUnderflow); loglk = xadd(loglk, LogLkUnderflow); }
                }
            }
        }
        *out = xadd(loglk, log(lk));                                                       // :1444
    }
}

// posteriorProfile for one position, NJ.tcc:2176-2261 (JC) / :2266-2334 (nt matrix) / :2340-2429 (aa, exactML)
template<typename P, int A>
__device__ __forceinline__ void posterior_site(const Store<P> &s, const MLModel<P> &m, const View<P, A> &p1, const View<P, A> &p2,
                                               int64_t pos, const P *ee1, const P *ee2, const double *PS1, const double *PD1,
                                               const double *PS2, const double *PD2, P &wOut, uint32_t &cOut, P (&fOut)[A]) {
    const uint32_t c1 = p1.codes[pos], c2 = p2.codes[pos];
    const P w1p = p1.w ? p1.w[pos] : (c1 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
    const P w2p = p2.w ? p2.w[pos] : (c2 != VFT_DEV_NOCODE ? (P) 1 : (P) 0);
    const int r = m.ratecat[pos];
#pragma unroll
    for (int j = 0; j < A; j++) fOut[j] = 0;
    cOut = VFT_DEV_NOCODE; wOut = (P) 1;
    if (m.codeFreq == nullptr) {                                                           // Jukes-Cantor
        const double w1 = (double) w1p, w2 = (double) w2p;
        const bool v1 = w1p > 0 && c1 == VFT_DEV_NOCODE && p1.v, v2 = w2p > 0 && c2 == VFT_DEV_NOCODE && p2.v;
        if (!v1 && !v2) {
            if (c1 == VFT_DEV_NOCODE && c2 == VFT_DEV_NOCODE) { wOut = 0; return; }
            if (c1 == VFT_DEV_NOCODE) { cOut = c2; wOut = (P) xmul(w2, xsub(PS2[r], PD2[r])); return; }     // :2196-2197
            if (c2 == VFT_DEV_NOCODE) { cOut = c1; wOut = (P) xmul(w1, xsub(PS1[r], PD1[r])); return; }
            if (c1 == c2) {                                                                // :2203-2225
                cOut = c1;
                const double a1 = xadd(xmul(w1, PS1[r]), xmul(xsub(1.0, w1), 0.25)), a2 = xadd(xmul(w2, PS2[r]), xmul(xsub(1.0, w2), 0.25));
                const double b1 = xadd(xmul(w1, PD1[r]), xmul(xsub(1.0, w1), 0.25)), b2 = xadd(xmul(w2, PD2[r]), xmul(xsub(1.0, w2), 0.25));
                const double f12code = xmul(a1, a2), f12other = xmul(b1, b2);
                const double pcode = f12code / xadd(f12code, xmul(3.0, f12other));
                wOut = (P) (xmul(xsub(pcode, 0.25), 4.0) / 3.0);
                if (wOut < (P) 1e-6) wOut = (P) 1e-6;
                return;
            }
        }
        P f1[4], f2[4];
        if (!v1) {                                                                         // :2231-2238
#pragma unroll
            for (int j = 0; j < 4; j++) f1[j] = (P) xmul(xsub(1.0, w1), 0.25);
#pragma unroll
            for (int j = 0; j < 4; j++) if ((uint32_t) j == c1) f1[j] = (P) xadd((double) f1[j], w1);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) f1[j] = p1.v[pos * A + j];
        }
        if (!v2) {
#pragma unroll
            for (int j = 0; j < 4; j++) f2[j] = (P) xmul(xsub(1.0, w2), 0.25);
#pragma unroll
            for (int j = 0; j < 4; j++) if ((uint32_t) j == c2) f2[j] = (P) xadd((double) f2[j], w2);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) f2[j] = p2.v[pos * A + j];
        }
        double lkAB = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {                                                      // :2252-2256
            const double t1 = xadd(xmul((double) f1[j], PS1[r]), xmul(xsub(1.0, (double) f1[j]), PD1[r]));
            const double t2 = xadd(xmul((double) f2[j], PS2[r]), xmul(xsub(1.0, (double) f2[j]), PD2[r]));
            fOut[j] = (P) xmul(t1, t2);
            lkAB = xadd(lkAB, (double) fOut[j]);
        }
        const double inv = 1.0 / lkAB;
#pragma unroll
        for (int j = 0; j < 4; j++) fOut[j] = (P) xmul((double) fOut[j], inv);
        return;
    }
    if (c1 == VFT_DEV_NOCODE && c2 == VFT_DEV_NOCODE && w1p == 0 && w2p == 0) { wOut = 0; return; }   // :2267 / :2341
    P f1[A], f2[A];
    ml_freq<P, A>(m, c1, w1p, p1.v ? p1.v + pos * A : nullptr, f1);
    ml_freq<P, A>(m, c2, w2p, p2.v ? p2.v + pos * A : nullptr, f2);
#pragma unroll
    for (int j = 0; j < A; j++) { f1[j] = pmul(f1[j], ee1[r * A + j]); f2[j] = pmul(f2[j], ee2[r * A + j]); }   // fMult
    P fPost[A];
    if (A == 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) {                                                      // :2311-2320
            double o1 = 0, o2 = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                o1 = xadd(o1, (double) pmul(f1[k], m.codeFreq[j * A + k]));
                o2 = xadd(o2, (double) pmul(f2[k], m.codeFreq[j * A + k]));
            }
            fPost[j] = (P) xmul(xmul(o1, o2), (double) m.statinv[j]);
        }
        double tot = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) tot = xadd(tot, (double) fPost[j]);
        const double inv = 1.0 / tot;
#pragma unroll
        for (int j = 0; j < 4; j++) fPost[j] = (P) xmul((double) fPost[j], inv);
        if (s.reduction == 0) {                                                            // BasicOperations.tcc:110-119
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double sum = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) sum = xadd(sum, (double) pmul(fPost[k], m.eigeninvT[k * A + j]));
                fOut[j] = (P) sum;
            }
        } else {                                                                           // AVX256Operations.tcc:277-303
            P o[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int k = 0; k < 4; k++) o[k] = padd(o[k], pmul(fPost[j], m.eigeninvT[j * A + k]));
#pragma unroll
            for (int k = 0; k < 4; k++) fOut[k] = o[k];
        }
        return;
    }
    for (int j = 0; j < A; j++) {                                                          // :2380-2385
        P cf[A];
#pragma unroll
        for (int k = 0; k < A; k++) cf[k] = m.codeFreq[j * A + k];
        const P a = vec_mul_sum<P, A>(f1, cf, s.reduction), b = vec_mul_sum<P, A>(f2, cf, s.reduction);
        const P value = pmul(pmul(a, b), m.statinv[j]);
        fPost[j] = value >= 0 ? value : (P) 0;
    }
    const double tot = (double) lane_fold<P, A>(fPost, s.reduction);                       // vector_sum, :2386
    const P inv = (P) (1.0 / tot);
#pragma unroll
    for (int j = 0; j < A; j++) fPost[j] = pmul(fPost[j], inv);                            // :2389
    for (int j = 0; j < A; j++) {                                                          // :2425-2427
        P ei[A];
#pragma unroll
        for (int k = 0; k < A; k++) ei[k] = m.eigeninv[j * A + k];
        fOut[j] = vec_mul_sum<P, A>(fPost, ei, s.reduction);
    }
}

}  // namespace vft
