// ml_host.cpp -- whole-tree likelihood sweeps as level-synchronous batches over the kernel-level ABI.
//
// The reference walks the tree node by node: recomputeMLProfiles() (NeighbourJoining.tcc:3508-3542) calls
// posteriorProfile once per 2-child internal node in post-order, treeLogLk() (NJ.tcc:5114-5259) calls pairLogLk
// once per internal node (plus posteriorProfile + pairLogLk at the 3-child root).  Neither sweep has a serial
// dependency beyond "children before parents", so here
//   - every tree LEVEL (height above the leaves) is one vft_posterior_profile_batch launch,
//   - all pairLogLk terms are one vft_pair_loglk_batch call (chunked when per-site likelihoods are wanted),
//   - the terms are then added in the reference's post-order (traversePostorder, NJ.tcc:3342-3377), which is what
//     fixes the rounding of the double-precision total.
// Compiled into the product library and, for the CPU tests, into the oracle double (oracle/Makefile).
#include "../../include/vft_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr double LkUnderflow = 1.0e-4, LkUnderflowInv = 1.0e4, LogLkUnderflow = 9.21034037197618;   // Constants.h:13-15

// post-order of the reference: children in index order, a node after its children (traversePostorder, NJ.tcc:3342-3377)
int post_order(int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child, std::vector<int64_t> &order) {
    order.clear();
    order.reserve((size_t) maxnode);
    std::vector<std::pair<int64_t, int>> st;
    st.push_back({root, 0});
    while (!st.empty()) {
        auto &top = st.back();
        if (top.second < nChild[top.first]) {
            const int64_t c = child[3 * top.first + top.second++];
            if (c < 0 || c >= maxnode) return VFT_EINVAL;
            st.push_back({c, 0});
        } else { order.push_back(top.first); st.pop_back(); }
    }
    return VFT_OK;
}

// recomputeMLProfiles, NJ.tcc:3508-3542: every 2-child node = posterior of its children; one launch per tree level
template<typename P>
int recompute_profiles(vft_ctx *ctx, const std::vector<int64_t> &order, int64_t maxnode, const int32_t *nChild, const int64_t *child,
                       const P *bl) {
    // height above the leaves; level h depends only on levels < h
    std::vector<int32_t> height((size_t) maxnode, 0);
    int32_t H = 0;
    for (int64_t node : order) {
        int32_t h = 0;
        for (int k = 0; k < nChild[node]; k++) h = std::max(h, height[child[3 * node + k]] + 1);
        height[node] = h; H = std::max(H, h);
    }
    std::vector<std::vector<int64_t>> levels((size_t) H + 1);
    for (int64_t node : order) if (nChild[node] == 2) levels[height[node]].push_back(node);   // :3508-3514
    std::vector<int64_t> o, a, b;
    std::vector<double> l1, l2;
    for (int32_t h = 1; h <= H; h++) {
        const auto &lv = levels[h];
        if (lv.empty()) continue;
        o.assign(lv.begin(), lv.end()); a.resize(lv.size()); b.resize(lv.size()); l1.resize(lv.size()); l2.resize(lv.size());
        for (size_t k = 0; k < lv.size(); k++) {
            a[k] = child[3 * lv[k]]; b[k] = child[3 * lv[k] + 1];
            l1[k] = (double) bl[a[k]]; l2[k] = (double) bl[b[k]];
        }
        int rc = vft_posterior_profile_batch(ctx, (int64_t) lv.size(), o.data(), a.data(), b.data(), l1.data(), l2.data());
        if (rc != VFT_OK) return rc;
    }
    return VFT_OK;
}

template<typename P>
int tree_loglk(vft_ctx *ctx, const vft_config &cfg, bool jukesCantor, int64_t root, int64_t maxnode, const int32_t *nChild,
               const int64_t *child, const P *bl, bool recompute, const uint8_t *leafCodes, double *loglkOut, double *siteLoglk) {
    const int64_t N = cfg.nSeqs, L = cfg.nPos;
    if (N < 2) { *loglkOut = 0.0; return VFT_OK; }                                          // NJ.tcc:5161-5163
    std::vector<int64_t> order;
    if (post_order(root, maxnode, nChild, child, order) != VFT_OK) return VFT_EINVAL;
    if (recompute) {
        int rc = recompute_profiles<P>(ctx, order, maxnode, nChild, child, bl);
        if (rc != VFT_OK) return rc;
    }
    // the pairLogLk terms, in post-order; the root's second term needs the posterior of its first two children
    std::vector<int64_t> pi, pj;
    std::vector<double> pl;
    for (int64_t node : order) {
        if (nChild[node] < 2) continue;
        const int64_t c0 = child[3 * node], c1 = child[3 * node + 1];
        pi.push_back(c0); pj.push_back(c1);
        pl.push_back((double) (P) (bl[c0] + bl[c1]));                                       // numeric_t sum, :5124-5125
        if (nChild[node] == 3) {                                                            // :5141-5149
            if (node != root) return VFT_EINVAL;
            int rc = vft_posterior_profile(ctx, root, c0, c1, (double) bl[c0], (double) bl[c1]);
            if (rc != VFT_OK) return rc;
            pi.push_back(root); pj.push_back(child[3 * node + 2]); pl.push_back((double) bl[child[3 * node + 2]]);
        }
    }
    const int64_t nTerms = (int64_t) pi.size();
    std::vector<double> term((size_t) nTerms);
    double loglk = 0.0;
    if (!siteLoglk) {
        int rc = vft_pair_loglk_batch(ctx, pi.data(), pj.data(), pl.data(), nTerms, term.data(), nullptr);
        if (rc != VFT_OK) return rc;
        // each node's contribution is summed on its own first (traverseTreeLogLk returns it, :5116-5157), so the two
        // terms of the 3-child root -- the last two -- are added to each other before they join the total (:5196)
        const int64_t nPlain = nChild[root] == 3 ? nTerms - 2 : nTerms;
        for (int64_t k = 0; k < nPlain; k++) loglk += term[k];
        if (nPlain < nTerms) loglk += term[nPlain] + term[nPlain + 1];
    } else {
        std::vector<double> siteLk((size_t) L, 1.0);                                        // :5168-5175
        for (int64_t i = 0; i < L; i++) siteLoglk[i] = 0.0;
        const int64_t CH = std::max<int64_t>(1, std::min<int64_t>(nTerms, (int64_t) (64 << 20) / (8 * L)));
        std::vector<double> rows((size_t) CH * L);
        double rootFirst = 0.0;
        for (int64_t k0 = 0; k0 < nTerms; k0 += CH) {
            const int64_t m = std::min(CH, nTerms - k0);
            int rc = vft_pair_loglk_batch(ctx, pi.data() + k0, pj.data() + k0, pl.data() + k0, m, term.data() + k0, rows.data());
            if (rc != VFT_OK) return rc;
            for (int64_t k = 0; k < m; k++) {
                // (the root's two terms are added to each other first, as traverseTreeLogLk does)
                if (nChild[root] == 3 && k0 + k == nTerms - 2) rootFirst = term[k0 + k];
                else if (nChild[root] == 3 && k0 + k == nTerms - 1) loglk += rootFirst + term[k0 + k];
                else loglk += term[k0 + k];
                const double *r = rows.data() + k * L;
                for (int64_t i = 0; i < L; i++) siteLk[i] *= r[i];                          // pairLogLk's site_likelihoods[i] *= lkAB
                if (pi[k0 + k] != root)                                                     // :5127-5135 (not after the root's second term)
                    for (int64_t i = 0; i < L; i++)
                        while (siteLk[i] < LkUnderflow) { siteLk[i] *= LkUnderflowInv; siteLoglk[i] -= LogLkUnderflow; }
            }
        }
        for (int64_t i = 0; i < L; i++) siteLoglk[i] += std::log(siteLk[i]);               // :5201-5205
    }
    if (cfg.nCodes == 4 && jukesCantor) {                                                   // :5231-5257
        if (!leafCodes) return VFT_EINVAL;
        int64_t nGaps = 0;
        const double logNCodes = std::log((double) cfg.nCodes);
        for (int64_t i = 0; i < L; i++) {
            int64_t g = 0;
            for (int64_t s = 0; s < N; s++) g += leafCodes[s * L + i] >= cfg.nCodes;
            nGaps += g;
            if (siteLoglk) { siteLoglk[i] += (double) g * logNCodes; siteLoglk[i] -= logNCodes; }
        }
        loglk -= (double) L * logNCodes;
        loglk += (double) nGaps * logNCodes;
    }
    *loglkOut = loglk;
    return VFT_OK;
}

// setMLRates, NJ.tcc:5429-5488 (+ MLSiteRates :5366-5377, MLSiteLikelihoodsByRate :5381-5408): nRateCats whole-tree
// sweeps, one per candidate rate, then the per-site choice with the Gamma(3,1/3) prior
template<typename P>
int set_ml_rates(vft_ctx *ctx, const vft_config &cfg, bool jukesCantor, int64_t root, int64_t maxnode, const int32_t *nChild,
                 const int64_t *child, const P *bl, int64_t nRateCats, double minRel, double minBr, int32_t fastexp,
                 const uint8_t *leafCodes, P *ratesOut, int64_t *ratecatOut, double *siteLoglkOut) {
    const int64_t L = cfg.nPos;
    std::vector<int64_t> order;
    if (post_order(root, maxnode, nChild, child, order) != VFT_OK) return VFT_EINVAL;
    std::vector<int64_t> zeros((size_t) L, 0);
    P one = (P) 1.0;
    int rc = vft_sync_rates(ctx, &one, 1, zeros.data(), minRel, minBr, fastexp);              // rates.reset(1, nPos), :5431
    if (rc != VFT_OK) return rc;
    if (nRateCats == 1) {                                                                       // :5433-5436
        ratesOut[0] = one;
        for (int64_t i = 0; i < L; i++) ratecatOut[i] = 0;
        return recompute_profiles<P>(ctx, order, maxnode, nChild, child, bl);
    }
    std::vector<P> rates((size_t) nRateCats);                                                   // MLSiteRates, :5366-5377
    {
        const double logNCat = std::log((double) nRateCats), logMinRate = -logNCat, logMaxRate = logNCat;
        const double logd = (logMaxRate - logMinRate) / (double) (nRateCats - 1);
        for (int64_t i = 0; i < nRateCats; i++) rates[i] = (P) std::exp(logMinRate + logd * (double) i);
    }
    std::vector<double> own;
    double *site = siteLoglkOut;
    if (!site) { own.resize((size_t) (nRateCats * L)); site = own.data(); }
    for (int64_t iRate = 0; iRate < nRateCats; iRate++) {                                       // MLSiteLikelihoodsByRate, :5389-5404
        rc = vft_sync_rates(ctx, &rates[iRate], 1, zeros.data(), minRel, minBr, fastexp);
        if (rc != VFT_OK) return rc;
        double lk;
        rc = tree_loglk<P>(ctx, cfg, jukesCantor, root, maxnode, nChild, child, bl, /*recompute*/true, leafCodes, &lk, site + L * iRate);
        if (rc != VFT_OK) return rc;
    }
    double sumRates = 0;                                                                        // :5449-5470
    for (int64_t iPos = 0; iPos < L; iPos++) {
        int64_t iBest = -1;
        double dBest = -1e20;
        for (int64_t iRate = 0; iRate < nRateCats; iRate++) {
            const double v = site[L * iRate + iPos] + 2.0 * std::log(rates[iRate]) - 3.0 * rates[iRate];
            if (v > dBest) { iBest = iRate; dBest = v; }
        }
        ratecatOut[iPos] = iBest;
        sumRates += rates[iBest];
    }
    const double avgRate = sumRates / L;                                                        // :5473-5476
    for (int64_t iRate = 0; iRate < nRateCats; iRate++) rates[iRate] /= avgRate;
    for (int64_t iRate = 0; iRate < nRateCats; iRate++) ratesOut[iRate] = rates[iRate];
    rc = vft_sync_rates(ctx, rates.data(), nRateCats, ratecatOut, minRel, minBr, fastexp);       // :5479
    if (rc != VFT_OK) return rc;
    return recompute_profiles<P>(ctx, order, maxnode, nChild, child, bl);                       // :5482
}

}  // namespace

extern "C" int vft_tree_loglk(vft_ctx *ctx, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child,
                              const void *branchlength, int32_t recomputeProfiles, const uint8_t *leafCodes, double *loglk,
                              double *siteLoglk) {
    if (!ctx || !nChild || !child || !branchlength || !loglk || root < 0 || root >= maxnode) return VFT_EINVAL;
    vft_config cfg;
    int32_t hasTransmat = 0;
    int rc = vft_get_config(ctx, &cfg, &hasTransmat);
    if (rc != VFT_OK) return rc;
    if (maxnode > 2 * cfg.nSeqs) return VFT_EINVAL;
    if (cfg.precision == 32)
        return tree_loglk<float>(ctx, cfg, !hasTransmat, root, maxnode, nChild, child, (const float *) branchlength, recomputeProfiles != 0,
                                 leafCodes, loglk, siteLoglk);
    return tree_loglk<double>(ctx, cfg, !hasTransmat, root, maxnode, nChild, child, (const double *) branchlength, recomputeProfiles != 0,
                              leafCodes, loglk, siteLoglk);
}

extern "C" int vft_set_ml_rates(vft_ctx *ctx, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child,
                                const void *branchlength, int64_t nRateCats, double MLMinRelBranchLength, double MLMinBranchLength,
                                int32_t fastexpLevel, const uint8_t *leafCodes, void *rates, int64_t *ratecat, double *siteLoglk) {
    if (!ctx || !nChild || !child || !branchlength || !rates || !ratecat || root < 0 || root >= maxnode || nRateCats < 1 || nRateCats > 64)
        return VFT_EINVAL;
    vft_config cfg;
    int32_t hasTransmat = 0;
    int rc = vft_get_config(ctx, &cfg, &hasTransmat);
    if (rc != VFT_OK) return rc;
    if (maxnode > 2 * cfg.nSeqs) return VFT_EINVAL;
    if (cfg.precision == 32)
        return set_ml_rates<float>(ctx, cfg, !hasTransmat, root, maxnode, nChild, child, (const float *) branchlength, nRateCats,
                                   MLMinRelBranchLength, MLMinBranchLength, fastexpLevel, leafCodes, (float *) rates, ratecat, siteLoglk);
    return set_ml_rates<double>(ctx, cfg, !hasTransmat, root, maxnode, nChild, child, (const double *) branchlength, nRateCats,
                                MLMinRelBranchLength, MLMinBranchLength, fastexpLevel, leafCodes, (double *) rates, ratecat, siteLoglk);
}

// recomputeProfiles, NJ.tcc:3474-3506 (the minimum-evolution counterpart of recomputeMLProfiles): one
// vft_profile_average_batch per tree level
extern "C" int vft_recompute_profiles(vft_ctx *ctx, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child) {
    if (!ctx || !nChild || !child || root < 0 || root >= maxnode) return VFT_EINVAL;
    std::vector<int64_t> order;
    if (post_order(root, maxnode, nChild, child, order) != VFT_OK) return VFT_EINVAL;
    std::vector<int32_t> height((size_t) maxnode, 0);
    int32_t H = 0;
    for (int64_t node : order) {
        int32_t h = 0;
        for (int k = 0; k < nChild[node]; k++) h = std::max(h, height[child[3 * node + k]] + 1);
        height[node] = h; H = std::max(H, h);
    }
    std::vector<std::vector<int64_t>> levels((size_t) H + 1);
    for (int64_t node : order) if (nChild[node] == 2) levels[height[node]].push_back(node);       // :3476
    std::vector<int64_t> a, b;
    for (int32_t h = 1; h <= H; h++) {
        const auto &lv = levels[h];
        if (lv.empty()) continue;
        a.resize(lv.size()); b.resize(lv.size());
        for (size_t k = 0; k < lv.size(); k++) { a[k] = child[3 * lv[k]]; b[k] = child[3 * lv[k] + 1]; }
        const int rc = vft_profile_average_batch(ctx, (int64_t) lv.size(), lv.data(), a.data(), b.data());
        if (rc != VFT_OK) return rc;
    }
    return VFT_OK;
}
