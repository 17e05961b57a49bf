// TEST INFRASTRUCTURE ONLY.
//
// White-box dumper: instantiates the reference's OWN templates
// (veryfasttree::NeighbourJoining<P, AVX256Operations>, DistanceMatrix<P,32>) from the headers
// where they lie under /root/reference/src -- nothing is copied or restated here -- and writes
// kernel-level values for a given alignment as a flat list of named arrays.  The vectors pin
// oracle/vft_oracle.c (and, through it, the CUDA kernels) at a finer grain than whole trees:
// tests/golden/make_golden.py commits its output, tests/test_oracle_golden.py replays it.
//
// usage: refdump <fasta> <nt|aa> <32|64> <out.bin> [tophits]
//
// File format: repeated { u32 name_len, name, char dtype('f','d','q','B'), u32 ndim, i64 dims[], data }.
#include <algorithm>
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <omp.h>

// private members are reached with g++ -fno-access-control (see oracle/Makefile); a
// `#define private public` would also rewrite the OpenMP private() clauses of the reference.
#include "operations/AVX256Operations.h"
#include "NeighbourJoining.h"

using namespace veryfasttree;
void ran_start(long seed);            // Knuth.cpp: the global generator behind knuth_rand()

static FILE *g_out;
static int64_t envInt(const char *name, int64_t dflt) { const char *e = std::getenv(name); return e && *e ? std::atoll(e) : dflt; }

static void put(const std::string &name, char dtype, const std::vector<int64_t> &dims, const void *data) {
    uint32_t nl = (uint32_t) name.size(), nd = (uint32_t) dims.size();
    size_t es = dtype == 'f' ? 4 : dtype == 'B' ? 1 : 8, n = 1;
    for (auto d : dims) n *= (size_t) d;
    fwrite(&nl, 4, 1, g_out); fwrite(name.data(), 1, nl, g_out); fwrite(&dtype, 1, 1, g_out);
    fwrite(&nd, 4, 1, g_out); fwrite(dims.data(), 8, nd, g_out); fwrite(data, es, n, g_out);
}
template<typename P> static char dt() { return sizeof(P) == 4 ? 'f' : 'd'; }
template<typename P> static void putv(const std::string &n, const std::vector<P> &v, std::vector<int64_t> dims = {}) {
    if (dims.empty()) dims = {(int64_t) v.size()};
    put(n, dt<P>(), dims, v.data());
}
static void putq(const std::string &n, const std::vector<int64_t> &v, std::vector<int64_t> dims = {}) {
    if (dims.empty()) dims = {(int64_t) v.size()};
    put(n, 'q', dims, v.data());
}

template<typename P>
struct Dumper {
    typedef NeighbourJoining<P, AVX256Operations> NJ;
    typedef typename NJ::Profile Profile;
    NJ &nj;
    int64_t L, A, S;
    explicit Dumper(NJ &nj) : nj(nj), L(nj.nPos), A(nj.options.nCodes), S(nj.nCodeSize) {}

    void profile(const std::string &pfx, Profile &p) {           // dense expansion of the sparse Profile
        std::vector<P> w(L), v(L * A, 0);
        std::vector<uint8_t> c(L);
        int64_t iv = 0;
        for (int64_t i = 0; i < L; i++) {
            w[i] = p.weights[i];
            c[i] = (uint8_t) p.codes[i];
            P *f = nj.getFreq(p, i, iv);
            if (f) for (int64_t k = 0; k < A; k++) v[i * A + k] = f[k];
        }
        putv<P>(pfx + ".weights", w);
        put(pfx + ".codes", 'B', {L}, c.data());
        putv<P>(pfx + ".vectors", v, {L, A});
        if (p.codeDistSize) {
            std::vector<P> cd(p.codeDist, p.codeDist + L * A);
            putv<P>(pfx + ".codeDist", cd, {L, A});
        }
    }
};

// ---- "ml" mode: pairLogLk (NJ.tcc:1192) and posteriorProfile (NJ.tcc:2137) of the reference, on ML-phase
//      profiles built by the reference's own posteriorProfile; model = jc | gtr (nt) | jtt (aa)
template<typename P>
static int runML(const std::string &fasta, bool aa, const std::string &model, int fastexpLvl) {
    Options options;
    options.verbose = 0; options.showProgress = false; options.threads = 1; options.diskComputing = false;
    options.nCodes = aa ? 20 : 4; options.useMatrix = aa;
    options.codesString = aa ? Constants::codesStringAA : Constants::codesStringNT;
    options.doublePrecision = sizeof(P) == 8;
    options.fastexp = fastexpLvl;
    options.fPostTotalTolerance = sizeof(P) == 8 ? Constants::fPostTotalToleranceDouble : Constants::fPostTotalToleranceFloat;
    options.MLMinBranchLength = sizeof(P) == 8 ? Constants::MLMinBranchLengthDouble : Constants::MLMinBranchLengthFloat;
    options.MLMinRelBranchLength = sizeof(P) == 8 ? Constants::MLMinRelBranchLengthDouble : Constants::MLMinRelBranchLengthFloat;
    omp_set_num_threads(1);
    std::ifstream in(fasta);
    if (!in) { std::fprintf(stderr, "cannot read %s\n", fasta.c_str()); return 2; }
    std::ostringstream lg;
    Alignment aln(options, in, lg);
    aln.readAlignment();
    std::vector<std::string> seqs = aln.seqs;
    std::vector<std::string> seqsTree = aln.seqs;      // the NeighbourJoining ctor consumes its sequences
    int64_t N = (int64_t) seqs.size(), L = aln.nPos, A = options.nCodes;
    typedef AVX256Operations<P> op_t;
    typedef NeighbourJoining<P, AVX256Operations> NJ;
    static DistanceMatrix<P, op_t::ALIGNMENT> dmat{};
    static TransitionMatrix<P, op_t::ALIGNMENT> transmat;
    if (aa) { dmat.matrixBLOSUM45(); dmat.setupDistanceMatrix(options, lg); }
    if (model == "jtt") transmat.createTransitionMatrixJTT92(options);
    else if (model == "gtr") {
        double r[6] = {1.3, 3.1, 0.7, 0.9, 4.2, 1.0}, f[4] = {0.31, 0.19, 0.23, 0.27};
        transmat.createGTR(options, r, f);
    }
    ProgressReport progress(false, 0, false);
    std::vector<std::string> cons;
    std::unique_ptr<DiskMemory> d1, d2;
    NJ nj(options, lg, progress, seqs, L, cons, dmat, transmat, d1, d2);
    Dumper<P> D(nj);
    putq("shape", {N, L, A});
    putq("ml.fastexp", {fastexpLvl});
    putq("ml.hasTransmat", {(int64_t) (bool) transmat});
    double mins[2] = {options.MLMinRelBranchLength, options.MLMinBranchLength};
    put("ml.minlen", 'd', {2}, mins);
    if (transmat) {
        std::vector<P> cf((A + 1) * A), ev(A), ei(A * A), eit(A * A), si(A);
        for (int64_t i = 0; i < A; i++) {
            ev[i] = transmat.eigenval[i]; si[i] = transmat.statinv[i];
            for (int64_t j = 0; j < A; j++) { cf[i * A + j] = transmat.codeFreq[i][j]; ei[i * A + j] = transmat.eigeninv[i][j]; eit[i * A + j] = transmat.eigeninvT[i][j]; }
        }
        for (int64_t j = 0; j < A; j++) cf[A * A + j] = transmat.codeFreq[NOCODE][j];
        putv<P>("ml.codeFreq", cf, {A + 1, A}); putv<P>("ml.eigenval", ev); putv<P>("ml.eigeninv", ei, {A, A});
        putv<P>("ml.eigeninvT", eit, {A, A}); putv<P>("ml.statinv", si);
    }
    // CAT rates: 4 categories, position i in category (i*7+i/3) % 4
    const int64_t nCat = 4;
    nj.rates.reset(nCat, L);
    const double rv[4] = {0.25, 0.8, 1.0, 2.6};
    std::vector<P> rates(nCat);
    std::vector<int64_t> ratecat(L);
    for (int64_t c = 0; c < nCat; c++) { nj.rates.rates[c] = (P) rv[c]; rates[c] = nj.rates.rates[c]; }
    for (int64_t i = 0; i < L; i++) { nj.rates.ratecat[i] = (i * 7 + i / 3) % nCat; ratecat[i] = nj.rates.ratecat[i]; }
    putv<P>("ml.rates", rates); putq("ml.ratecat", ratecat);

    // posterior profiles: a script of (out, a, b, len1, len2)
    nj.parent.assign(nj.maxnodes, -1);
    int64_t K = std::min<int64_t>(14, N - 2);
    std::vector<int64_t> script;
    std::vector<double> lens;
    for (int64_t k = 0; k < K; k++) {
        int64_t a, b;
        if (k % 3 == 0) { a = (k * 5) % N; b = (k * 5 + 3) % N; }
        else if (k % 3 == 1) { a = N + k - 1; b = (k * 11 + 2) % N; }
        else { a = N + k - 1; b = N + k - 2; }
        double l1 = k == 4 ? 1e-7 : 0.02 + 0.013 * (double) k, l2 = k == 7 ? 0.0 : 0.11 - 0.006 * (double) k;
        int64_t out = N + k;
        nj.posteriorProfile(nj.profiles[out], nj.profiles[a], nj.profiles[b], l1, l2);
        nj.maxnode = out + 1;
        script.push_back(out); script.push_back(a); script.push_back(b);
        lens.push_back(l1); lens.push_back(l2);
        D.profile("post" + std::to_string(k), nj.profiles[out]);
    }
    putq("ml.post.script", script, {K, 3});
    put("ml.post.lens", 'd', {K, 2}, lens.data());

    // pair log-likelihoods over leaves and the posterior profiles
    std::vector<int64_t> pi, pj;
    std::vector<double> pl;
    for (int64_t x = 0; x < N + K; x += 2) {
        pi.push_back(x); pj.push_back((x * 3 + 1) % (N + K)); pl.push_back(0.01 + 0.004 * (double) (x % 37));
        pi.push_back((x * 5 + 2) % N); pj.push_back(N + (x % K)); pl.push_back(0.3 + 0.05 * (double) (x % 9));
    }
    pi.push_back(N + K - 1); pj.push_back(N + K - 2); pl.push_back(1e-9);      // below MLMinRelBranchLength
    pi.push_back(0); pj.push_back(0); pl.push_back(0.5);
    pi.push_back(N + 1); pj.push_back(N + 1); pl.push_back(2.5);
    std::vector<double> ll(pi.size());
    std::vector<double> site(L * 3, 1.0);
    for (size_t k = 0; k < pi.size(); k++) {
        double *sl = k < 3 ? &site[k * L] : nullptr;
        ll[k] = nj.pairLogLk(nj.profiles[pi[k]], nj.profiles[pj[k]], pl[k], sl);
    }
    putq("ml.lk.i", pi); putq("ml.lk.j", pj);
    put("ml.lk.len", 'd', {(int64_t) pl.size()}, pl.data());
    put("ml.lk.loglk", 'd', {(int64_t) ll.size()}, ll.data());
    put("ml.lk.site", 'd', {3, L}, site.data());

    // whole-tree sweeps on the NJ tree of the same alignment: recomputeMLProfiles (NJ.tcc:3516-3542) + treeLogLk
    // (NJ.tcc:5114-5259), with the same rates
    {
        std::unique_ptr<DiskMemory> e1, e2;
        NJ nj2(options, lg, progress, seqsTree, L, cons, dmat, transmat, e1, e2);
        nj2.fastNJ();
        nj2.rates.reset(nCat, L);
        for (int64_t c = 0; c < nCat; c++) nj2.rates.rates[c] = (P) rv[c];
        for (int64_t i = 0; i < L; i++) nj2.rates.ratecat[i] = (i * 7 + i / 3) % nCat;
        nj2.recomputeMLProfiles();
        std::vector<double> siteLk(L, 0.0);
        double lk = nj2.treeLogLk(siteLk.data());
        double lkNoSite = nj2.treeLogLk(nullptr);
        std::vector<int64_t> nch(nj2.maxnode), ch(nj2.maxnode * 3, -1), par(nj2.maxnode);
        std::vector<P> bl(nj2.maxnode);
        for (int64_t i = 0; i < nj2.maxnode; i++) {
            nch[i] = nj2.child[i].nChild; par[i] = nj2.parent[i]; bl[i] = nj2.branchlength[i];
            for (int k = 0; k < nj2.child[i].nChild; k++) ch[i * 3 + k] = nj2.child[i].child[k];
        }
        putq("ml.tree.root", {nj2.root});
        putq("ml.tree.nChild", nch); putq("ml.tree.child", ch, {nj2.maxnode, 3}); putq("ml.tree.parent", par);
        putv<P>("ml.tree.branchlength", bl);
        double lks[2] = {lk, lkNoSite};
        put("ml.tree.loglk", 'd', {2}, lks);
        put("ml.tree.site", 'd', {L}, siteLk.data());
        Dumper<P> DT(nj2);
        DT.profile("ml.tree.lastnode", nj2.profiles[nj2.root - 1]);
        // setMLRates (NJ.tcc:5429-5488) with 6 candidate rates on the same tree
        options.nRateCats = envInt("VFT_REFDUMP_NCAT", 6);        // the large on-the-fly cases of tests/test_gpu_parity.py ask for the default 20
        nj2.setMLRates();
        std::vector<P> r6(nj2.rates.rates.begin(), nj2.rates.rates.end());
        std::vector<int64_t> rc6(nj2.rates.ratecat.begin(), nj2.rates.ratecat.end());
        putv<P>("ml.cat.rates", r6); putq("ml.cat.ratecat", rc6);
        double lkCat = nj2.treeLogLk(nullptr);
        put("ml.cat.loglk", 'd', {1}, &lkCat);

        // ---- SURVEY 8a rows a15-a17 under the CAT rates just chosen: MLPairOptimize (NJ.tcc:1790), MLQuartetNNI
        //      (:4885; MLQuartetOptimize :1650, onedimenmin :7024, brent :7098), the per-node body of
        //      optimizeAllBranchLengths (:5044-5058) and the whole sweep (:5006-5112); all by the reference itself
        typedef typename NJ::Profile Profile;
        options.MLFTolBranchLength = sizeof(P) == 8 ? Constants::MLFTolBranchLengthDouble : Constants::MLFTolBranchLengthFloat;
        options.MLMinBranchLengthTolerance = sizeof(P) == 8 ? Constants::MLMinBranchLengthToleranceDouble : Constants::MLMinBranchLengthToleranceFloat;
        double optScalars[4] = {options.MLMinBranchLength, options.MLFTolBranchLength, options.MLMinBranchLengthTolerance, Constants::closeLogLkLimit};
        put("ml.opt.scalars", 'd', {4}, optScalars);
        putq("ml.opt.flags", {(int64_t) options.mlAccuracy, (int64_t) options.fastNNI});
        const int64_t M2 = nj2.maxnode;
        {
            std::vector<int64_t> pa, pb;
            std::vector<double> l0, l1, lk;
            for (int64_t k = 0; k < 10; k++) {
                int64_t a = (k * 3) % N, b = k % 2 ? N + (k * 5) % (M2 - 1 - N) : (k * 7 + 4) % N;
                if (a == b) b = (b + 1) % N;
                double len = k == 0 ? options.MLMinBranchLength : k == 1 ? 1.5 * options.MLMinBranchLength : k == 2 ? 4.0 : 0.03 + 0.09 * (double) k;
                pa.push_back(a); pb.push_back(b); l0.push_back(len);
                lk.push_back(nj2.MLPairOptimize(nj2.profiles[a], nj2.profiles[b], &len));
                l1.push_back(len);
            }
            putq("ml.opt.pair.a", pa); putq("ml.opt.pair.b", pb);
            put("ml.opt.pair.len0", 'd', {(int64_t) l0.size()}, l0.data());
            put("ml.opt.pair.len1", 'd', {(int64_t) l1.size()}, l1.data());
            put("ml.opt.pair.loglk", 'd', {(int64_t) lk.size()}, lk.data());
        }
        {
            std::vector<std::unique_ptr<Profile>> upProfiles(nj2.maxnodes);
            // the per-split body of testSplitsML (NJ.tcc:6884-6952) + SHSupport on the same quartets, 100 resamples
            options.nBootstrap = 100;
            std::vector<int64_t> shCol;
            nj2.resampleColumns(shCol);
            std::vector<double> spLk, spSite, spSupport;
            std::vector<int64_t> spChoice, spBad;
            std::vector<int64_t> qids, qnode, qchoice;
            std::vector<P> qlen0, qlen1;
            std::vector<double> qcrit;
            int64_t nQ = 0;
            const int64_t qStride = envInt("VFT_REFDUMP_QSTRIDE", 1);   // every qStride-th internal node (bounds the dump of a large case)
            for (int64_t node = N; node < M2; node += qStride) {
                if (node == nj2.root || nj2.child[node].nChild != 2) continue;
                Profile *p4[4];
                int64_t abcd[4];
                nj2.setupABCD(node, p4, upProfiles.data(), abcd, true);
                DT.profile("ml.opt.q" + std::to_string(nQ) + ".D", *p4[3]);
                for (int fast = 1; fast >= 0; fast--) {
                    P len[5] = {nj2.branchlength[abcd[0]], nj2.branchlength[abcd[1]], nj2.branchlength[abcd[2]], nj2.branchlength[abcd[3]], nj2.branchlength[node]};
                    double crit[3] = {0.0, 0.0, 0.0};
                    for (int i = 0; i < 5; i++) qlen0.push_back(len[i]);
                    int choice = (int) nj2.MLQuartetNNI(p4, crit, len, fast != 0);
                    for (int i = 0; i < 5; i++) qlen1.push_back(len[i]);
                    for (int i = 0; i < 3; i++) qcrit.push_back(crit[i]);
                    qchoice.push_back(choice);
                }
                {
                    double l5[5] = {(double) nj2.branchlength[abcd[0]], (double) nj2.branchlength[abcd[1]], (double) nj2.branchlength[abcd[2]],
                                    (double) nj2.branchlength[abcd[3]], (double) nj2.branchlength[node]};
                    double lAB[5] = {l5[0], l5[1], l5[2], l5[3], l5[4]}, lAC[5] = {l5[0], l5[2], l5[1], l5[3], l5[4]}, lAD[5] = {l5[0], l5[3], l5[2], l5[1], l5[4]};
                    std::vector<double> site(3 * L);
                    double lk3[3];
                    lk3[0] = nj2.MLQuartetLogLk(*p4[0], *p4[1], *p4[2], *p4[3], lAB, &site[0]);
                    lk3[1] = nj2.MLQuartetOptimize(*p4[0], *p4[2], *p4[1], *p4[3], lAC, nullptr, &site[L]);
                    lk3[2] = nj2.MLQuartetOptimize(*p4[0], *p4[3], *p4[2], *p4[1], lAD, nullptr, &site[2 * L]);
                    if (lk3[1] > lk3[2]) {
                        if (options.mlAccuracy > 1 || lk3[1] > lk3[0] - Constants::closeLogLkLimit)
                            lk3[1] = nj2.MLQuartetOptimize(*p4[0], *p4[2], *p4[1], *p4[3], lAC, nullptr, &site[L]);
                    } else {
                        if (options.mlAccuracy > 1 || lk3[2] > lk3[0] - Constants::closeLogLkLimit)
                            lk3[2] = nj2.MLQuartetOptimize(*p4[0], *p4[3], *p4[2], *p4[1], lAD, nullptr, &site[2 * L]);
                    }
                    int ch = lk3[0] >= lk3[1] && lk3[0] >= lk3[2] ? 0 : (lk3[1] >= lk3[0] && lk3[1] >= lk3[2] ? 1 : 2);
                    bool bad = lk3[ch] > lk3[0] + Constants::treeLogLkDelta;
                    spLk.insert(spLk.end(), lk3, lk3 + 3); spSite.insert(spSite.end(), site.begin(), site.end());
                    spChoice.push_back(ch); spBad.push_back(bad ? 1 : 0);
                    spSupport.push_back(bad ? 0.0 : nj2.SHSupport(shCol, lk3, site));
                }
                for (int i = 0; i < 4; i++) qids.push_back(abcd[i]);
                qnode.push_back(node);
                nQ++;
            }
            putq("ml.opt.q.node", qnode); putq("ml.opt.q.ids", qids, {nQ, 4}); putq("ml.opt.q.choice", qchoice, {nQ, 2});
            putv<P>("ml.opt.q.len0", qlen0, {nQ, 2, 5}); putv<P>("ml.opt.q.len1", qlen1, {nQ, 2, 5});
            put("ml.opt.q.criteria", 'd', {nQ, 2, 3}, qcrit.data());
            putq("ml.opt.q.nStar", {(int64_t) options.debug.nStarTests});
            putq("ml.split.col", shCol, {(int64_t) options.nBootstrap, L});
            put("ml.split.loglk", 'd', {nQ, 3}, spLk.data()); put("ml.split.site", 'd', {nQ, 3, L}, spSite.data());
            putq("ml.split.choice", spChoice); putq("ml.split.bad", spBad);
            put("ml.split.support", 'd', {nQ}, spSupport.data());
            options.nBootstrap = 1000;
            // the per-node body of traverseOptimizeAllBranchLengths on a few nodes (state untouched: local lengths)
            std::vector<int64_t> snode, sids;
            std::vector<P> slen0, slen1;
            int64_t nS = 0;
            for (int64_t node = N; node < M2; node += 3 * qStride) {
                const int64_t nChild = nj2.child[node].nChild;
                if (nChild < 2) continue;
                int64_t nodes[3] = {nj2.child[node].child[0], nj2.child[node].child[1], nChild == 3 ? nj2.child[node].child[2] : node};
                Profile *p3[3] = {&nj2.profiles[nodes[0]], &nj2.profiles[nodes[1]],
                                  nChild == 3 ? &nj2.profiles[nodes[2]] : nj2.getUpProfile(upProfiles.data(), node, true)};
                DT.profile("ml.opt.s" + std::to_string(nS) + ".U", *p3[2]);
                P bl3[3] = {nj2.branchlength[nodes[0]], nj2.branchlength[nodes[1]], nj2.branchlength[nodes[2]]};
                for (int i = 0; i < 3; i++) slen0.push_back(bl3[i]);
                for (int iter = 0; iter < 2; iter++)
                    for (int i = 0; i < 3; i++) {
                        int b1 = (i + 1) % 3, b2 = (i + 2) % 3;
                        Profile pB(L, nj2.nCons);
                        nj2.posteriorProfile(pB, *p3[b1], *p3[b2], bl3[b1], bl3[b2]);
                        double len = bl3[i];
                        if (len < options.MLMinBranchLength) len = options.MLMinBranchLength;
                        nj2.MLPairOptimize(*p3[i], pB, &len);
                        bl3[i] = len;
                    }
                for (int i = 0; i < 3; i++) { slen1.push_back(bl3[i]); sids.push_back(nodes[i]); }
                snode.push_back(node);
                nS++;
            }
            putq("ml.opt.s.node", snode); putq("ml.opt.s.ids", sids, {nS, 3});
            putv<P>("ml.opt.s.len0", slen0, {nS, 3}); putv<P>("ml.opt.s.len1", slen1, {nS, 3});
        }
        // the whole sweep, then the tree likelihood with the new lengths
        nj2.optimizeAllBranchLengths();
        std::vector<P> blOpt(M2);
        for (int64_t i = 0; i < M2; i++) blOpt[i] = nj2.branchlength[i];
        putv<P>("ml.opt.tree.branchlength", blOpt);
        double lkOpt = nj2.treeLogLk(nullptr);
        put("ml.opt.tree.loglk", 'd', {1}, &lkOpt);
        // testSplitsML (NJ.tcc:6800-7000) on the optimised tree: the reference draws its resampled columns from the global
        // Knuth generator, so the generator is seeded, the columns drawn here, and seeded again for the call itself
        {
            options.nBootstrap = 100;
            ran_start(20260117L);
            std::vector<int64_t> colT;
            nj2.resampleColumns(colT);
            ran_start(20260117L);
            typename NJ::SplitCount sc;
            nj2.testSplitsML(sc);
            std::vector<P> sup(M2);
            for (int64_t i = 0; i < M2; i++) sup[i] = nj2.support[i];
            putq("ml.splits.col", colT, {(int64_t) options.nBootstrap, L});
            putv<P>("ml.splits.support", sup);
            putq("ml.splits.nBad", {(int64_t) sc.nBadSplits, (int64_t) sc.nSplits});
            options.nBootstrap = 1000;
        }
    }
    return 0;
}

template<typename P>
static int run(const std::string &fasta, bool aa, bool tophitsOnly) {
    Options options;
    options.verbose = 0;
    options.showProgress = false;
    options.threads = 1;
    options.diskComputing = false;
    options.nCodes = aa ? 20 : 4;
    options.useMatrix = aa;
    options.codesString = aa ? Constants::codesStringAA : Constants::codesStringNT;
    options.doublePrecision = sizeof(P) == 8;
    options.fPostTotalTolerance = sizeof(P) == 8 ? Constants::fPostTotalToleranceDouble : Constants::fPostTotalToleranceFloat;
    options.MLMinBranchLength = sizeof(P) == 8 ? Constants::MLMinBranchLengthDouble : Constants::MLMinBranchLengthFloat;
    options.MLMinRelBranchLength = sizeof(P) == 8 ? Constants::MLMinRelBranchLengthDouble : Constants::MLMinRelBranchLengthFloat;
    omp_set_num_threads(1);

    std::ifstream in(fasta);
    if (!in) { std::fprintf(stderr, "cannot read %s\n", fasta.c_str()); return 2; }
    std::ostringstream lg;
    Alignment aln(options, in, lg);
    aln.readAlignment();
    std::vector<std::string> seqs = aln.seqs;     // keep a copy: the ctor releases its input strings
    int64_t N = (int64_t) seqs.size(), L = aln.nPos;

    typedef AVX256Operations<P> op_t;
    static DistanceMatrix<P, op_t::ALIGNMENT> dmat{};
    static TransitionMatrix<P, op_t::ALIGNMENT> transmat;
    if (options.useMatrix) {
        dmat.matrixBLOSUM45();
        dmat.setupDistanceMatrix(options, lg);
        std::vector<P> d(20 * 20), ev(20), et(20), cf(20 * 20);
        for (int i = 0; i < 20; i++) {
            ev[i] = dmat.eigenval[i]; et[i] = dmat.eigentot[i];
            for (int j = 0; j < 20; j++) { d[i * 20 + j] = dmat.distances[i][j]; cf[i * 20 + j] = dmat.codeFreq[i][j]; }
        }
        putv<P>("tables.distances", d, {20, 20});
        putv<P>("tables.eigenval", ev);
        putv<P>("tables.eigentot", et);
        putv<P>("tables.codeFreq", cf, {20, 20});
    }

    ProgressReport progress(false, 0, false);
    std::vector<std::string> cons;
    std::unique_ptr<DiskMemory> d1, d2;
    NeighbourJoining<P, AVX256Operations> nj(options, lg, progress, seqs, L, cons, dmat, transmat, d1, d2);
    Dumper<P> D(nj);
    int64_t A = options.nCodes;
    putq("shape", {N, L, A});

    if (tophitsOnly) {
        // the reference's own setAllLeafTopHits (NJ.tcc:3746-4124) at -threads 1: the top-hit list of
        // every leaf, indices and distances -- the "identical top-hit indices" golden
        int64_t m = (int64_t) (0.5 + options.tophitsMult * std::sqrt((double) N));
        if (m < 4 || 2 * m >= N) { std::fprintf(stderr, "too few leaves for top hits\n"); return 3; }
        nj.parent.assign(nj.maxnodes, -1);
        typename NeighbourJoining<P, AVX256Operations>::TopHits tophits(options, nj.maxnodes, m);
        nj.setAllLeafTopHits(tophits);
        std::vector<int64_t> js(N * m, -1), vis(N);
        std::vector<P> ds(N * m, 0);
        for (int64_t i = 0; i < N; i++) {
            auto &l = tophits.topHitsLists[i].hits;
            for (size_t k = 0; k < l.size(); k++) { js[i * m + k] = l[k].j; ds[i * m + k] = l[k].dist; }
            vis[i] = tophits.visible[i].j;
        }
        putq("tophits.m", {m});
        putq("tophits.j", js, {N, m});
        putv<P>("tophits.dist", ds, {N, m});
        putq("tophits.visible", vis);
        return 0;
    }

    // --- state right after the constructor (NJ.tcc:237-260)
    D.profile("ctor.outprofile", nj.outprofile);
    {
        std::vector<P> od(nj.outDistances.begin(), nj.outDistances.begin() + N), sw(nj.selfweight.begin(), nj.selfweight.begin() + N);
        putv<P>("ctor.outDistances", od);
        putv<P>("ctor.selfweight", sw);
    }

    // --- leaf x leaf, seqDist (NJ.tcc:1601-1624)
    {
        std::vector<int64_t> pi, pj;
        for (int64_t i = 0; i < std::min<int64_t>(N, 48); i++) {
            pi.push_back(i); pj.push_back((i * 7 + 3) % N);
            pi.push_back(i); pj.push_back(i);
            pi.push_back(N - 1 - i); pj.push_back((i * 13 + 1) % N);
        }
        std::vector<P> dd(pi.size()), ww(pi.size());
        for (size_t k = 0; k < pi.size(); k++) {
            typename NeighbourJoining<P, AVX256Operations>::Besthit h;
            nj.seqDist(nj.profiles[pi[k]].codes, nj.profiles[pj[k]].codes, h);
            dd[k] = h.dist; ww[k] = h.weight;
        }
        putq("seq.i", pi); putq("seq.j", pj); putv<P>("seq.dist", dd); putv<P>("seq.weight", ww);
    }

    // --- a fixed script of joins: averageProfile (NJ.tcc:2067), selfdist (:3041), updateOutProfile (:943),
    //     setOutDistance (:1012), profileDist (:1167); the test replays exactly this script
    nj.parent.assign(nj.maxnodes, -1);
    std::vector<int64_t> joins;
    std::vector<double> diams;
    std::vector<int64_t> active;
    for (int64_t i = 0; i < N; i++) active.push_back(i);
    int64_t nJoins = std::min<int64_t>(24, N - 4);
    int64_t nActive = N;
    std::vector<P> step_out_w, step_out_v;
    for (int64_t k = 0; k < nJoins; k++) {
        // alternate leaf+leaf, node+leaf, node+node picks, deterministic
        int64_t a, b;
        if (k % 3 == 0 || nj.maxnode - N < 2) { a = active[(k * 5) % active.size()]; b = active[(k * 5 + 1 + k) % active.size()]; }
        else if (k % 3 == 1) { a = nj.maxnode - 1; b = active[(k * 11 + 2) % active.size()]; }
        else { a = nj.maxnode - 1; b = nj.maxnode - 2; }
        if (a == b || nj.parent[a] >= 0 || nj.parent[b] >= 0) {
            a = active[0]; b = active[1];
        }
        int64_t nw = nj.maxnode++;
        double diam = 0.003 * (double) (k + 1) + 0.001 * (double) (k % 4);
        nj.diameter[nw] = (P) diam;
        nj.averageProfile(nj.profiles[nw], nj.profiles[a], nj.profiles[b], -1.0);
        typename NeighbourJoining<P, AVX256Operations>::Besthit sd;
        nj.profileDist(nj.profiles[nw], nj.profiles[nw], sd);
        nj.selfdist[nw] = sd.dist; nj.selfweight[nw] = sd.weight;
        nj.updateOutProfile(nj.outprofile, nj.profiles[a], nj.profiles[b], nj.profiles[nw], nActive);
        nj.totdiam += nj.diameter[nw] - nj.diameter[a] - nj.diameter[b];
        nj.parent[a] = nw; nj.parent[b] = nw;
        active.erase(std::find(active.begin(), active.end(), a));
        active.erase(std::find(active.begin(), active.end(), b));
        active.push_back(nw);
        std::sort(active.begin(), active.end());
        nActive--;
        joins.push_back(a); joins.push_back(b); diams.push_back(diam);
        D.profile("join" + std::to_string(k) + ".profile", nj.profiles[nw]);
        std::vector<P> self = {nj.selfdist[nw], nj.selfweight[nw]};
        putv<P>("join" + std::to_string(k) + ".self", self);
        if (k == 0 || k == nJoins - 1) D.profile("join" + std::to_string(k) + ".outprofile", nj.outprofile);
    }
    putq("joins", joins, {nJoins, 2});
    put("diameters", 'd', {nJoins}, diams.data());
    double td = nj.totdiam;
    put("totdiam", 'd', {1}, &td);
    putq("nActive", {nActive});

    // out-distances of every active node at this nActive (setOutDistance, NJ.tcc:1012-1053)
    {
        std::vector<P> od;
        for (int64_t id : active) { nj.nOutDistActive[id] = -1; nj.setOutDistance(id, nActive); od.push_back(nj.outDistances[id]); }
        putq("out.ids", active); putv<P>("out.dist", od);
    }
    // profile distances among internal nodes and leaves (profileDist, NJ.tcc:1167-1190), raw
    {
        std::vector<int64_t> pi, pj;
        for (int64_t x = N; x < nj.maxnode; x++) {
            pi.push_back(x); pj.push_back(x);
            pi.push_back(x); pj.push_back((x * 3) % N);
            pi.push_back((x * 5 + 1) % N); pj.push_back(x);
            if (x + 1 < nj.maxnode) { pi.push_back(x); pj.push_back(x + 1); }
            pi.push_back(nj.maxnode - 1); pj.push_back(x);
        }
        for (int64_t i = 0; i < std::min<int64_t>(N, 16); i++) { pi.push_back(i); pj.push_back((i * 7 + 3) % N); }  // leaf pairs through profileDist
        std::vector<P> dd(pi.size()), ww(pi.size());
        for (size_t k = 0; k < pi.size(); k++) {
            typename NeighbourJoining<P, AVX256Operations>::Besthit h;
            nj.profileDist(nj.profiles[pi[k]], nj.profiles[pj[k]], h);
            dd[k] = h.dist; ww[k] = h.weight;
        }
        putq("prof.i", pi); putq("prof.j", pj); putv<P>("prof.dist", dd); putv<P>("prof.weight", ww);
    }
    // setDistCriterion's distance (NJ.tcc:1115-1122) for mixed pairs
    {
        std::vector<int64_t> pi, pj;
        for (size_t x = 0; x + 1 < active.size() && x < 40; x++) { pi.push_back(active[x]); pj.push_back(active[active.size() - 1 - x]); }
        std::vector<P> dd(pi.size()), ww(pi.size());
        for (size_t k = 0; k < pi.size(); k++) {
            typename NeighbourJoining<P, AVX256Operations>::Besthit h;
            h.i = pi[k]; h.j = pj[k];
            nj.nOutDistActive[h.i] = nActive; nj.nOutDistActive[h.j] = nActive;
            nj.setDistCriterion(nActive, h);
            dd[k] = h.dist; ww[k] = h.weight;
        }
        putq("join.i", pi); putq("join.j", pj); putv<P>("join.dist", dd); putv<P>("join.weight", ww);
    }
    // chooseNNI (NJ.tcc:4836-4852) over quartets of active nodes: default options, then with a pseudo-count prior and
    // without the log correction
    {
        typedef typename NeighbourJoining<P, AVX256Operations>::Profile Profile;
        std::vector<int64_t> q;
        for (size_t x = 0; x + 3 < active.size() && x < 24; x++) {
            q.push_back(active[x]); q.push_back(active[active.size() - 1 - x]); q.push_back(active[(x * 7 + 2) % active.size()]); q.push_back(active[(x * 3 + 5) % active.size()]);
        }
        const int64_t nq = (int64_t) q.size() / 4;
        for (int variant = 0; variant < 3; variant++) {
            options.pseudoWeight = variant == 1 ? 1.0 : 0.0;
            options.logdist = variant != 2;
            std::vector<double> crit;
            std::vector<int64_t> ch;
            for (int64_t k = 0; k < nq; k++) {
                Profile *p4[4] = {&nj.profiles[q[4 * k]], &nj.profiles[q[4 * k + 1]], &nj.profiles[q[4 * k + 2]], &nj.profiles[q[4 * k + 3]]};
                double c[3];
                ch.push_back((int64_t) nj.chooseNNI(p4, c));
                crit.insert(crit.end(), c, c + 3);
            }
            put("nni" + std::to_string(variant) + ".criteria", 'd', {nq, 3}, crit.data());
            putq("nni" + std::to_string(variant) + ".choice", ch);
        }
        options.pseudoWeight = 0.0; options.logdist = true;
        putq("nni.ids", q, {nq, 4});
    }
    // SHSupport (NJ.tcc:1126-1165) with the reference's own resampled columns (resampleColumns, NJ.tcc:705-716): synthetic
    // per-site likelihoods for the three topologies of a few quartets -- near-ties, a clear winner, a losing first topology
    {
        options.nBootstrap = 200;
        std::vector<int64_t> col;
        nj.resampleColumns(col);
        const int64_t nq = 12;
        std::vector<double> lk3(nq * 3), sl(nq * 3 * L), sup(nq);
        uint64_t st = 0x9E3779B97F4A7C15ull;
        auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double) (st >> 11) / 9007199254740992.0; };
        for (int64_t q = 0; q < nq; q++) {
            const double tilt = q % 4 == 0 ? 0.0 : q % 4 == 1 ? 0.004 : q % 4 == 2 ? 0.05 : -0.01;
            for (int t = 0; t < 3; t++) {
                double sum = 0;
                for (int64_t i = 0; i < L; i++) {
                    const double base = 0.02 + 0.9 * rnd();
                    const double v = base * (1.0 + (t == 0 ? tilt : 0.0) * (rnd() - 0.3)) * (0.97 + 0.06 * rnd());
                    sl[(q * 3 + t) * L + i] = v;
                    sum += std::log(v);
                }
                lk3[q * 3 + t] = sum;
            }
            std::vector<double> sv(sl.begin() + q * 3 * L, sl.begin() + (q + 1) * 3 * L);
            sup[q] = nj.SHSupport(col, &lk3[q * 3], sv);
        }
        putq("sh.col", col, {(int64_t) options.nBootstrap, L});
        put("sh.loglk", 'd', {nq, 3}, lk3.data());
        put("sh.siteLk", 'd', {nq, 3, L}, sl.data());
        put("sh.support", 'd', {nq}, sup.data());
        options.nBootstrap = 1000;
    }
    // full out-profile rebuild over the active set (outProfile, NJ.tcc:729-815)
    {
        typedef typename NeighbourJoining<P, AVX256Operations>::Profile Profile;
        std::vector<Profile *> ap;
        for (int64_t id : active) ap.push_back(&nj.profiles[id]);
        nj.outProfile(nj.outprofile, ap, (int64_t) ap.size());
        D.profile("rebuild.outprofile", nj.outprofile);
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 5 || argc > 8) { std::fprintf(stderr, "usage: refdump <fasta> <nt|aa> <32|64> <out.bin> [tophits | ml <jc|gtr|jtt> <fastexp>]\n"); return 2; }
    bool th = argc == 6 && std::string(argv[5]) == "tophits";
    if (argc == 8 && std::string(argv[5]) == "ml") {
        g_out = std::fopen(argv[4], "wb");
        if (!g_out) { std::perror(argv[4]); return 2; }
        bool aa2 = std::string(argv[2]) == "aa";
        int lvl = std::atoi(argv[7]);
        int rc2 = std::string(argv[3]) == "64" ? runML<double>(argv[1], aa2, argv[6], lvl) : runML<float>(argv[1], aa2, argv[6], lvl);
        std::fclose(g_out);
        return rc2;
    }
    g_out = std::fopen(argv[4], "wb");
    if (!g_out) { std::perror(argv[4]); return 2; }
    bool aa = std::string(argv[2]) == "aa";
    int rc = std::string(argv[3]) == "64" ? run<double>(argv[1], aa, th) : run<float>(argv[1], aa, th);
    std::fclose(g_out);
    return rc;
}
