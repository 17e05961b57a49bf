// ml_opt.cpp -- branch-length optimisation and ML NNI quartets as LOCK-STEP batches over the likelihood ABI
// (MLPairOptimize / MLQuartetOptimize / onedimenmin+brent, MLQuartetNNI,
// optimizeAllBranchLengths).
//
// The reference optimises one branch at a time: Brent's minimiser (NeighbourJoining.tcc:7098-7190) asks for ONE
// pairLogLk per iteration, MLQuartetOptimize (NJ.tcc:1650-1788) chains five such minimisations with
// posteriorProfile calls in between, MLQuartetNNI (NJ.tcc:4885-5004) chains up to three of those per round.  On a
// GPU a single (pair, length) evaluation is a few microseconds of work behind a launch + synchronisation, so the
// unit that fills the machine is MANY independent optimisations advancing together: every task below is written as
// the reference writes it -- sequential code, as a C++20 coroutine -- and suspends where the reference would call
// pairLogLk or posteriorProfile; the scheduler gathers the pending requests of all tasks, issues ONE
// vft_posterior_profile_batch and ONE vft_pair_loglk_batch for the round, hands the values back and resumes the
// tasks.  n tasks of ~60 evaluations each cost ~60 device round trips instead of 60 n.
//
// Each task performs exactly the reference's `-threads 1` sequence of double-precision operations on the values
// it is given, so over the CPU double of the ABI the results are bit-identical to the reference
// (tests/test_oracle_golden.py); on the device they differ only through libm exp/log inside pairLogLk.
// Independence of the tasks of one call (no task reads a profile row another one writes) is the caller's
// contract, as it is for the reference's own OpenMP sections / tree partitions (NJ.tcc:4925-4951, :5086-5107).
//
// Compiled into the product library and, for the CPU tests, into the oracle double (oracle/Makefile).
#include "../../include/vft_b200.h"

#include <algorithm>
#include <cmath>
#include <coroutine>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <utility>
#include <vector>

namespace {

// ---- coroutine plumbing ------------------------------------------------------------------------------------
// Task<T>: lazily started, resumes its awaiter when it finishes (symmetric transfer), so tasks nest like calls.
template<class T>
struct Task {
    struct promise_type {
        T value{};
        std::coroutine_handle<> waiter;
        Task get_return_object() { return Task{std::coroutine_handle<promise_type>::from_promise(*this)}; }
        std::suspend_always initial_suspend() noexcept { return {}; }
        struct Done {
            bool await_ready() const noexcept { return false; }
            std::coroutine_handle<> await_suspend(std::coroutine_handle<promise_type> h) const noexcept {
                std::coroutine_handle<> w = h.promise().waiter;
                return w ? w : std::noop_coroutine();
            }
            void await_resume() const noexcept {}
        };
        Done final_suspend() noexcept { return {}; }
        void return_value(T v) { value = v; }
        void unhandled_exception() { std::terminate(); }
    };
    std::coroutine_handle<promise_type> h;
    explicit Task(std::coroutine_handle<promise_type> h) : h(h) {}
    Task(Task &&o) noexcept : h(std::exchange(o.h, nullptr)) {}
    Task(const Task &) = delete;
    Task &operator=(const Task &) = delete;
    Task &operator=(Task &&o) noexcept { if (this != &o) { if (h) h.destroy(); h = std::exchange(o.h, nullptr); } return *this; }
    ~Task() { if (h) h.destroy(); }
    bool await_ready() const noexcept { return false; }
    std::coroutine_handle<> await_suspend(std::coroutine_handle<> w) noexcept { h.promise().waiter = w; return h; }
    T await_resume() { return h.promise().value; }
    bool done() const { return h.done(); }
};

// One round = the requests every live task is waiting on.  A task waits on either up to two posteriorProfile
// items or up to two pairLogLk items (never both): posteriors are issued first, so the order inside a round
// cannot matter to an independent task.
struct Batcher {
    vft_ctx *ctx;
    std::vector<int64_t> lkA, lkB;
    std::vector<double> lkX;
    std::vector<double *> lkOut, lkSite;                     // lkSite[k]: per-site likelihoods are multiplied into this array (or nullptr)
    int64_t L = 0;                                           // nPos, needed only when a site array is asked for
    std::vector<double> siteBuf;
    std::vector<std::coroutine_handle<>> waiters;
    std::vector<int64_t> poOut, poA, poB;
    std::vector<double> poL1, poL2, lkVal;
    vft_ml_stats stats{};

    struct LkWait {
        Batcher *b; int n; int64_t a[2], bb[2]; double x[2], val[2]; double *site = nullptr;
        bool await_ready() const noexcept { return false; }
        void await_suspend(std::coroutine_handle<> h) {
            for (int k = 0; k < n; k++) { b->lkA.push_back(a[k]); b->lkB.push_back(bb[k]); b->lkX.push_back(x[k]); b->lkOut.push_back(&val[k]); b->lkSite.push_back(site); }
            b->waiters.push_back(h);
        }
        std::pair<double, double> await_resume() const noexcept { return {val[0], val[1]}; }
    };
    struct PostWait {
        Batcher *b; int n; int64_t o[2], p1[2], p2[2]; double l1[2], l2[2];
        bool await_ready() const noexcept { return false; }
        void await_suspend(std::coroutine_handle<> h) {
            for (int k = 0; k < n; k++) { b->poOut.push_back(o[k]); b->poA.push_back(p1[k]); b->poB.push_back(p2[k]); b->poL1.push_back(l1[k]); b->poL2.push_back(l2[k]); }
            b->waiters.push_back(h);
        }
        void await_resume() const noexcept {}
    };
    // pairLogLk(a, b, x) and a pair of them evaluated in the same round
    // site != nullptr: pairLogLk's site_likelihoods[] argument (NJ.tcc:1263-1265, :1436): every per-site value is multiplied in,
    // the first item of a pair before the second
    LkWait lk(int64_t a, int64_t b, double x, double *site = nullptr) { return LkWait{this, 1, {a, 0}, {b, 0}, {x, 0}, {0, 0}, site}; }
    LkWait lk2(int64_t a0, int64_t b0, double x0, int64_t a1, int64_t b1, double x1, double *site = nullptr) {
        return LkWait{this, 2, {a0, a1}, {b0, b1}, {x0, x1}, {0, 0}, site};
    }
    // posteriorProfile(out <- p1, p2); two independent ones in the same round
    PostWait post(int64_t o, int64_t p1, int64_t p2, double l1, double l2) { return PostWait{this, 1, {o, 0}, {p1, 0}, {p2, 0}, {l1, 0}, {l2, 0}}; }
    PostWait post2(int64_t o0, int64_t p0, int64_t q0, double l0, double m0, int64_t o1, int64_t p1, int64_t q1, double l1, double m1) {
        return PostWait{this, 2, {o0, o1}, {p0, p1}, {q0, q1}, {l0, l1}, {m0, m1}};
    }

    // one round: the pending posteriors, then the pending log-likelihoods, then every waiting task moves on
    int step() {
        stats.rounds++;
        if (!poOut.empty()) {
            int rc = vft_posterior_profile_batch(ctx, (int64_t) poOut.size(), poOut.data(), poA.data(), poB.data(), poL1.data(), poL2.data());
            if (rc != VFT_OK) return rc;
            stats.posteriorCalls++; stats.posteriorItems += (int64_t) poOut.size();
            poOut.clear(); poA.clear(); poB.clear(); poL1.clear(); poL2.clear();
        }
        if (!lkA.empty()) {
            // the (rare) items that want their per-site likelihoods go in a call of their own, so that the bulk of a round
            // carries no [n][nPos] result array
            size_t nSite = 0;
            for (double *p : lkSite) nSite += p != nullptr;
            if (nSite > 0 && nSite < lkA.size()) {               // stable partition: plain items first
                std::vector<size_t> order;
                for (size_t k = 0; k < lkA.size(); k++) if (!lkSite[k]) order.push_back(k);
                for (size_t k = 0; k < lkA.size(); k++) if (lkSite[k]) order.push_back(k);
                auto perm = [&](auto &v) { auto c = v; for (size_t k = 0; k < order.size(); k++) v[k] = c[order[k]]; };
                perm(lkA); perm(lkB); perm(lkX); perm(lkOut); perm(lkSite);
            }
            const size_t nPlain = lkA.size() - nSite;
            lkVal.resize(lkA.size());
            if (nPlain > 0) {
                int rc = vft_pair_loglk_batch(ctx, lkA.data(), lkB.data(), lkX.data(), (int64_t) nPlain, lkVal.data(), nullptr);
                if (rc != VFT_OK) return rc;
                stats.loglkCalls++;
            }
            if (nSite > 0) {
                siteBuf.resize(nSite * (size_t) L);
                int rc = vft_pair_loglk_batch(ctx, lkA.data() + nPlain, lkB.data() + nPlain, lkX.data() + nPlain, (int64_t) nSite,
                                              lkVal.data() + nPlain, siteBuf.data());
                if (rc != VFT_OK) return rc;
                stats.loglkCalls++;
                for (size_t k = 0; k < nSite; k++) {              // in request order: a task's first item before its second
                    double *dst = lkSite[nPlain + k];
                    const double *row = siteBuf.data() + k * (size_t) L;
                    for (int64_t j = 0; j < L; j++) dst[j] *= row[j];
                }
            }
            stats.loglkItems += (int64_t) lkA.size();
            for (size_t k = 0; k < lkA.size(); k++) *lkOut[k] = lkVal[k];
            lkA.clear(); lkB.clear(); lkX.clear(); lkOut.clear(); lkSite.clear();
        }
        now.swap(waiters);
        waiters.clear();
        for (auto h : now) h.resume();
        now.clear();
        return VFT_OK;
    }
    std::vector<std::coroutine_handle<>> now;

    template<class T>
    int run(std::vector<Task<T>> &tasks) {
        for (auto &t : tasks) t.h.resume();                    // up to the first request (or the end)
        while (!waiters.empty()) {
            const int rc = step();
            if (rc != VFT_OK) return rc;
        }
        return VFT_OK;
    }
};

struct Opt {                      // the scalars of Options / Constants the optimisers read
    double minLen, ftol, atol, closeLimit;
    int mlAccuracy;
    bool fast, single;            // single: numeric_t is float (branch lengths are stored narrowed)
    double store(double x) const { return single ? (double) (float) x : x; }
};

// ---- Brent's minimiser over x -> -pairLogLk(a, b, x), NJ.tcc:7098-7190 (pairNegLogLk, :1449-1458) -------------
// Golden-section steps with parabolic interpolation through the three best points; both a fractional (ftol) and an
// absolute (atol) stopping rule; at most 100 iterations.  The second derivative the reference also returns is not
// read by any caller on this path and is not computed.
struct Bracket { double lo, mid, hi, fLo, fMid, fHi; };

Task<double> brentMin(Batcher &B, int64_t pa, int64_t pb, Bracket br, double ftol, double atol, double *fOpt) {
    const double golden = 0.3819660, zeps = 1.0e-10;
    double a = br.lo < br.hi ? br.lo : br.hi, b = br.lo > br.hi ? br.lo : br.hi;
    double x = br.mid, fx = br.fMid, w, fw, v, fv;
    if (br.fLo < br.fHi) { w = br.lo; fw = br.fLo; v = br.hi; fv = br.fHi; }
    else { w = br.hi; fw = br.fHi; v = br.lo; fv = br.fLo; }
    double step = 0, prevStep = 0;                          // d and e of the textbook statement
    for (int iter = 1; iter <= 100; iter++) {
        const double xm = 0.5 * (a + b);
        const double tol1 = ftol * std::fabs(x), tol2 = 2.0 * (tol1 + zeps);
        if (std::fabs(x - xm) <= (tol2 - 0.5 * (b - a)) || std::fabs(a - b) < atol) break;
        bool goldenStep = true;
        if (std::fabs(prevStep) > tol1) {                   // try the parabola through (v, w, x)
            double r = (x - w) * (fx - fv);
            double q = (x - v) * (fx - fw);
            double p = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = std::fabs(q);
            const double before = prevStep;
            prevStep = step;
            if (!(std::fabs(p) >= std::fabs(0.5 * q * before) || p <= q * (a - x) || p >= q * (b - x))) {
                goldenStep = false;
                step = p / q;
                const double u = x + step;
                if (u - a < tol2 || b - u < tol2) step = (xm - x) >= 0.0 ? std::fabs(tol1) : -std::fabs(tol1);
            }
        }
        if (goldenStep) {
            prevStep = x >= xm ? a - x : b - x;
            step = golden * prevStep;
        }
        const double u = std::fabs(step) >= tol1 ? x + step : x + (step >= 0.0 ? std::fabs(tol1) : -std::fabs(tol1));
        const double fu = -(co_await B.lk(pa, pb, u)).first;
        if (fu <= fx) {
            if (u >= x) a = x; else b = x;
            v = w; w = x; x = u;
            fv = fw; fw = fx; fx = fu;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) { v = w; w = u; fv = fw; fw = fu; }
            else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
        }
    }
    *fOpt = fx;
    co_return x;
}

// onedimenmin, NJ.tcc:7024-7081: bracket around the guess (halved / doubled, clipped to [xmin, xmax]), widened until
// the middle point is the lowest or the bound is reached, then Brent
Task<double> oneDimMin(Batcher &B, int64_t pa, int64_t pb, double xmin, double xguess, double xmax, double ftol, double atol,
                       double *fOpt) {
    Bracket br;
    if (xguess == xmin) { br.lo = xmin; br.mid = 2.0 * xguess; br.hi = 10.0 * xguess; }
    else if (xguess <= 2.0 * xmin) { br.lo = xmin; br.mid = xguess; br.hi = 5.0 * xguess; }
    else { br.lo = 0.5 * xguess; br.mid = xguess; br.hi = 2.0 * xguess; }
    if (br.hi > xmax) br.hi = xmax;
    if (br.mid >= br.hi) br.mid = 0.5 * (br.lo + br.hi);
    br.fLo = -(co_await B.lk(pa, pb, br.lo)).first;
    br.fMid = -(co_await B.lk(pa, pb, br.mid)).first;
    br.fHi = -(co_await B.lk(pa, pb, br.hi)).first;
    while (br.fLo < br.fMid && br.lo > xmin) {
        br.lo = (br.lo + xmin) / 2.0;
        if (br.lo < 2.0 * xmin) br.lo = xmin;
        br.fLo = -(co_await B.lk(pa, pb, br.lo)).first;
    }
    while (br.fHi < br.fMid && br.hi < xmax) {
        br.hi = (br.hi + xmax) / 2.0;
        if (br.hi > xmax * 0.95) br.hi = xmax;
        br.fHi = -(co_await B.lk(pa, pb, br.hi)).first;
    }
    co_return co_await brentMin(B, pa, pb, br, ftol, atol, fOpt);
}

// MLPairOptimize, NJ.tcc:1790-1803: *len <- argmax pairLogLk(a, b, .), returns the log-likelihood there
Task<double> pairOptimize(Batcher &B, const Opt &o, int64_t pa, int64_t pb, double *len) {
    double neg = 0;
    *len = co_await oneDimMin(B, pa, pb, o.minLen, *len, 6.0, o.ftol, o.atol, &neg);
    co_return -neg;
}

enum { LEN_A = 0, LEN_B, LEN_C, LEN_D, LEN_I };            // NJ.h: order of the five quartet branches

// MLQuartetOptimize, NJ.tcc:1650-1788.  rows[3]: scratch profile rows for AB, CD and the changing third profile
// (BCD, ACD, ABD, ABC in turn -- the reference's stack Profiles).
Task<double> quartetOptimize(Batcher &B, const Opt &o, const int64_t q[4], double len[5], bool *starTest, const int64_t rows[3],
                             double *site = nullptr) {
    const int64_t pA = q[0], pB = q[1], pC = q[2], pD = q[3], AB = rows[0], CD = rows[1], X = rows[2];
    for (int j = 0; j < 5; j++) if (len[j] < o.minLen) len[j] = o.minLen;
    if (starTest) *starTest = false;
    double neg = 0;
    // the internal branch first (:1672-1687)
    co_await B.post2(AB, pA, pB, len[LEN_A], len[LEN_B], CD, pC, pD, len[LEN_C], len[LEN_D]);
    len[LEN_I] = co_await oneDimMin(B, AB, CD, o.minLen, len[LEN_I], 6.0, o.ftol, o.atol, &neg);
    if (starTest) {                                           // :1689-1698
        const double loglkStar = (co_await B.lk(AB, CD, o.minLen)).first;
        if (loglkStar < -neg - o.closeLimit) {
            *starTest = true;
            const auto off = co_await B.lk2(pA, pB, len[LEN_A] + len[LEN_B], pC, pD, len[LEN_C] + len[LEN_D]);
            co_return -neg + (off.first + off.second);
        }
    }
    co_await B.post(X, pB, CD, len[LEN_B], len[LEN_I]);       // BCD, :1700-1714
    len[LEN_A] = co_await oneDimMin(B, pA, X, o.minLen, len[LEN_A], 6.0, o.ftol, o.atol, &neg);
    co_await B.post(X, pA, CD, len[LEN_A], len[LEN_I]);       // ACD, :1716-1730
    len[LEN_B] = co_await oneDimMin(B, pB, X, o.minLen, len[LEN_B], 6.0, o.ftol, o.atol, &neg);
    co_await B.post(AB, pA, pB, len[LEN_A], len[LEN_B]);      // :1731
    co_await B.post(X, AB, pD, len[LEN_I], len[LEN_D]);       // ABD, :1733-1747
    len[LEN_C] = co_await oneDimMin(B, pC, X, o.minLen, len[LEN_C], 6.0, o.ftol, o.atol, &neg);
    co_await B.post(X, AB, pC, len[LEN_I], len[LEN_C]);       // ABC, :1749-1762
    len[LEN_D] = co_await oneDimMin(B, pD, X, o.minLen, len[LEN_D], 6.0, o.ftol, o.atol, &neg);
    // PairLogLk(ABC,D) + PairLogLk(AB,C) + PairLogLk(A,B), :1764-1775; with site likelihoods the first term is evaluated once
    // more to collect them (:1767-1772), the value of the optimisation is what enters the sum
    if (site) {
        for (int64_t j = 0; j < B.L; j++) site[j] = 1.0;
        (void) co_await B.lk(X, pD, len[LEN_D], site);
    }
    const auto rest = co_await B.lk2(AB, pC, len[LEN_I] + len[LEN_C], pA, pB, len[LEN_A] + len[LEN_B], site);
    co_return (-neg + rest.first) + rest.second;
}

struct QuartetJob { int64_t q[4]; double len[5]; double criteria[3]; int32_t choice; int32_t star; int64_t rows[3]; };

// MLQuartetNNI, NJ.tcc:4885-5004 (no topological constraints: penalties are 0).  o.fast selects which of the
// reference's two branches is followed: its serial branch (:4905-4923, taken at `-threads 1` and inside parallel
// regions) hands MLQuartetOptimize a star-test flag whatever bFast says; its sections branch (:4925-4951) never does.
Task<int> quartetNNI(Batcher &B, const Opt &o, QuartetJob *job) {
    double *len = job->len;
    double ab[5] = {len[LEN_A], len[LEN_B], len[LEN_C], len[LEN_D], len[LEN_I]};
    double ac[5] = {len[LEN_A], len[LEN_C], len[LEN_B], len[LEN_D], len[LEN_I]};          // B and C swapped
    double ad[5] = {len[LEN_A], len[LEN_D], len[LEN_C], len[LEN_B], len[LEN_I]};          // B and D swapped
    const int64_t *q = job->q;
    const int64_t qAC[4] = {q[0], q[2], q[1], q[3]}, qAD[4] = {q[0], q[3], q[2], q[1]};
    bool considerAC = true, considerAD = true;
    double *crit = job->criteria;
    const int nRounds = o.mlAccuracy < 2 ? 2 : o.mlAccuracy;
    job->star = 0;
    for (int round = 0; round < nRounds; round++) {
        bool star = false;
        crit[0] = co_await quartetOptimize(B, o, q, ab, o.fast ? &star : nullptr, job->rows);
        if (star) {                                           // :4912-4918
            crit[1] = -1e20; crit[2] = -1e20;
            len[LEN_I] = o.store(ab[LEN_I]);
            job->star = 1; job->choice = 0;
            co_return 0;
        }
        if (considerAC) crit[1] = co_await quartetOptimize(B, o, qAC, ac, nullptr, job->rows);
        if (considerAD) crit[2] = co_await quartetOptimize(B, o, qAD, ad, nullptr, job->rows);
        if (o.mlAccuracy < 2) {                               // :4962-4985
            if (crit[1] < crit[0] - o.closeLimit || (ac[LEN_I] <= 2.0 * o.minLen && crit[1] < crit[0])) considerAC = false;
            if (crit[2] < crit[0] - o.closeLimit || (ad[LEN_I] <= 2.0 * o.minLen && crit[2] < crit[0])) considerAD = false;
            if (!considerAC && !considerAD) break;
            if (crit[1] > crit[0] + o.closeLimit && crit[1] > crit[2] + o.closeLimit) break;
            if (crit[2] > crit[0] + o.closeLimit && crit[2] > crit[1] + o.closeLimit) break;
        }
    }
    const double *best = ab;
    job->choice = 0;
    if (crit[1] > crit[0] && crit[1] > crit[2]) { best = ac; job->choice = 1; }
    else if (crit[2] > crit[0] && crit[2] > crit[1]) { best = ad; job->choice = 2; }
    for (int i = 0; i < 5; i++) len[i] = o.store(best[i]);
    co_return 0;
}

// MLQuartetLogLk, NJ.tcc:5410-5427: the quartet likelihood at the given lengths, no optimisation; rows[0..1] = AB, CD
Task<double> quartetLogLk(Batcher &B, const int64_t q[4], const double len[5], const int64_t rows[3], double *site) {
    co_await B.post2(rows[0], q[0], q[1], len[0], len[1], rows[1], q[2], q[3], len[2], len[3]);
    if (site) for (int64_t j = 0; j < B.L; j++) site[j] = 1.0;
    const auto ab = co_await B.lk2(q[0], q[1], len[0] + len[1], q[2], q[3], len[2] + len[3], site);
    const double abcd = (co_await B.lk(rows[0], rows[1], len[4], site)).first;
    co_return (ab.first + ab.second) + abcd;
}

// The per-split body of traverseTestSplitsML, NJ.tcc:6884-6952: likelihood of the split as it is, the two alternatives
// optimised (the better one a second time when it is close), each with its per-site likelihoods -- the input of SHSupport
struct SplitJob { int64_t q[4]; double len[5]; double loglk[3]; int32_t choice, bad; int64_t rows[3]; double *site; };

Task<int> splitTest(Batcher &B, const Opt &o, SplitJob *job) {
    const int64_t *q = job->q;
    const double *len = job->len;
    double ab[5] = {len[LEN_A], len[LEN_B], len[LEN_C], len[LEN_D], len[LEN_I]};
    double ac[5] = {len[LEN_A], len[LEN_C], len[LEN_B], len[LEN_D], len[LEN_I]};
    double ad[5] = {len[LEN_A], len[LEN_D], len[LEN_C], len[LEN_B], len[LEN_I]};
    const int64_t qAC[4] = {q[0], q[2], q[1], q[3]}, qAD[4] = {q[0], q[3], q[2], q[1]};
    double *s0 = job->site, *s1 = s0 + B.L, *s2 = s1 + B.L;
    double *lk = job->loglk;
    lk[0] = co_await quartetLogLk(B, q, ab, job->rows, s0);
    lk[1] = co_await quartetOptimize(B, o, qAC, ac, nullptr, job->rows, s1);
    lk[2] = co_await quartetOptimize(B, o, qAD, ad, nullptr, job->rows, s2);
    if (lk[1] > lk[2]) {                                                                      // :6929-6941
        if (o.mlAccuracy > 1 || lk[1] > lk[0] - o.closeLimit) lk[1] = co_await quartetOptimize(B, o, qAC, ac, nullptr, job->rows, s1);
    } else {
        if (o.mlAccuracy > 1 || lk[2] > lk[0] - o.closeLimit) lk[2] = co_await quartetOptimize(B, o, qAD, ad, nullptr, job->rows, s2);
    }
    if (lk[0] >= lk[1] && lk[0] >= lk[2]) job->choice = 0;                                    // :6943-6950
    else if (lk[1] >= lk[0] && lk[1] >= lk[2]) job->choice = 1;
    else job->choice = 2;
    job->bad = lk[job->choice] > lk[0] + 0.1;                                                 // Constants::treeLogLkDelta, :6952
    co_return 0;
}

// The body of traverseOptimizeAllBranchLengths for one node, NJ.tcc:5044-5058: two sweeps over the three branches
// that meet at the node; branch i is optimised against the posterior of the other two.  len[] are the stored
// (numeric_t) branch lengths, updated in place.  row: one scratch profile row.
struct StarJob { int64_t p[3]; double len[3]; int64_t row; };

Task<int> starOptimize(Batcher &B, const Opt &o, StarJob *job) {
    for (int iter = 0; iter < 2; iter++)
        for (int i = 0; i < 3; i++) {
            const int b1 = (i + 1) % 3, b2 = (i + 2) % 3;
            co_await B.post(job->row, job->p[b1], job->p[b2], job->len[b1], job->len[b2]);
            double x = job->len[i];
            if (x < o.minLen) x = o.minLen;
            co_await pairOptimize(B, o, job->p[i], job->row, &x);
            job->len[i] = o.store(x);
        }
    co_return 0;
}

Task<int> pairJob(Batcher &B, const Opt &o, int64_t a, int64_t b, double *len, double *loglk) {
    *loglk = co_await pairOptimize(B, o, a, b, len);
    co_return 0;
}

// ---- the tree: post-order, parents, the reference's lazily built up-profiles -----------------------------------
struct Tree {
    int64_t root, maxnode, N;
    const int32_t *nChild;
    const int64_t *child;
    std::vector<int64_t> parent, order;
    int build() {
        parent.assign((size_t) maxnode, -1);
        order.clear();
        std::vector<std::pair<int64_t, int>> st;
        st.push_back({root, 0});
        while (!st.empty()) {                                 // traversePostorder, NJ.tcc:3342-3377
            auto &top = st.back();
            if (top.second < nChild[top.first]) {
                const int64_t c = child[3 * top.first + top.second++];
                if (c < 0 || c >= maxnode || parent[c] >= 0 || c == root) return VFT_EINVAL;
                parent[c] = top.first;
                st.push_back({c, 0});
            } else { order.push_back(top.first); st.pop_back(); }
        }
        for (int64_t node : order) {
            if (nChild[node] == 0) { if (node >= N) return VFT_EINVAL; }
            else if (node == root ? nChild[node] != 3 : nChild[node] != 2) return VFT_EINVAL;
        }
        return VFT_OK;
    }
    int64_t sibling(int64_t node) const {                     // NJ.tcc:1976-1989
        const int64_t p = parent[node];
        for (int k = 0; k < nChild[p]; k++) if (child[3 * p + k] != node) return child[3 * p + k];
        return -1;
    }
};

struct RowPool {                                              // scratch profile rows 2N .. 2N+S-1
    std::vector<int64_t> freeRows;
    RowPool(int64_t first, int64_t n) { for (int64_t k = n - 1; k >= 0; k--) freeRows.push_back(first + k); }
    int64_t take() { if (freeRows.empty()) return -1; int64_t r = freeRows.back(); freeRows.pop_back(); return r; }
    void give(int64_t r) { freeRows.push_back(r); }
};

// getUpProfile, NJ.tcc:3382-3434 (useML): the profile of everything that is NOT below `node`, built down the path
// from the root; up[n] = posterior(C, D) where, for a child of the root, C and D are its two root siblings
// (setupABCD, :1942-1974) and otherwise C is the sibling and D the up-profile of the parent.  up[]: row or -1.
template<typename P>
Task<int> upProfile(Batcher &B, const Tree &t, const P *bl, std::vector<int64_t> &up, RowPool &pool, int64_t outnode) {
    if (up[outnode] >= 0) co_return 0;
    std::vector<int64_t> path;
    for (int64_t n = outnode; n != t.root; n = t.parent[n]) path.push_back(n);
    for (int64_t k = (int64_t) path.size() - 1; k >= 0; k--) {
        const int64_t node = path[k];
        if (up[node] >= 0) continue;
        const int64_t par = t.parent[node];
        int64_t c, d, dRow;
        if (par == t.root) {                                  // rootSiblings, :1991-2003
            int64_t sibs[2], ns = 0;
            for (int j = 0; j < 3; j++) if (t.child[3 * par + j] != node) sibs[ns++] = t.child[3 * par + j];
            c = sibs[0]; d = sibs[1]; dRow = d;
        } else { c = t.sibling(node); d = par; dRow = up[par]; }
        const int64_t row = pool.take();
        if (row < 0) co_return VFT_ENOMEM;
        up[node] = row;
        co_await B.post(row, c, dRow, (double) bl[c], (double) bl[d]);        // :3419
    }
    co_return 0;
}

// optimizeAllBranchLengths in the reference's own order (NJ.tcc:5006-5112 at `-threads 1`): ONE task walks the tree
template<typename P>
Task<int> optimizeSequential(Batcher &B, const Opt &o, const Tree &t, P *bl, RowPool &pool) {
    std::vector<int64_t> up((size_t) t.maxnode, -1);
    const int64_t tmp = pool.take();
    if (tmp < 0) co_return VFT_ENOMEM;
    for (int64_t node : t.order) {
        if (t.nChild[node] == 0) continue;
        StarJob job;
        int64_t nodes[3] = {t.child[3 * node], t.child[3 * node + 1], node == t.root ? t.child[3 * node + 2] : node};
        if (node != t.root) {
            const int rc = co_await upProfile<P>(B, t, bl, up, pool, node);
            if (rc != VFT_OK) co_return rc;
        }
        for (int i = 0; i < 3; i++) { job.p[i] = nodes[i]; job.len[i] = (double) bl[nodes[i]]; }
        if (node != t.root) job.p[2] = up[node];
        job.row = tmp;
        co_await starOptimize(B, o, &job);
        for (int i = 0; i < 3; i++) bl[nodes[i]] = (P) job.len[i];
        if (node != t.root) {                                 // recomputeProfile + upProfiles[node].reset(), :5059-5062
            co_await B.post(node, nodes[0], nodes[1], (double) bl[nodes[0]], (double) bl[nodes[1]]);
            pool.give(up[node]);
            up[node] = -1;
        }
    }
    co_return 0;
}

// one node of the throughput schedule: the three-branch optimisation, the new lengths, the node's own profile rebuilt
struct NodeJob { int64_t node; StarJob job; };
template<typename P>
Task<int> nodeTask(Batcher &B, const Opt &o, const Tree &t, P *bl, const std::vector<int64_t> &up, NodeJob *nj) {
    const int64_t node = nj->node;
    const int64_t nodes[3] = {t.child[3 * node], t.child[3 * node + 1], node == t.root ? t.child[3 * node + 2] : node};
    for (int i = 0; i < 3; i++) { nj->job.p[i] = nodes[i]; nj->job.len[i] = (double) bl[nodes[i]]; }
    if (node != t.root) nj->job.p[2] = up[node];
    co_await starOptimize(B, o, &nj->job);
    for (int i = 0; i < 3; i++) bl[nodes[i]] = (P) nj->job.len[i];
    if (node != t.root) co_await B.post(node, nodes[0], nodes[1], (double) bl[nodes[0]], (double) bl[nodes[1]]);
    co_return 0;
}

// The throughput schedule: the same per-node work, but all nodes of one tree LEVEL (height above the leaves) advance
// in lock-step.  Up-profiles of the whole tree are built top-down first (one posterior batch per depth), then the
// levels are processed bottom-up, each followed by one posterior batch that rebuilds the level's own profiles.
// Within a level every node sees the state left by the previous level (Jacobi-style), where the reference's
// sequential sweep sees the updates of the nodes visited just before (Gauss-Seidel-style) -- the same kind of
// difference the reference's own tree-partitioned OpenMP mode has (NJ.tcc:5086-5107).  Needs one scratch row per
// internal node for the up-profiles plus one per node of the widest chunk.
template<typename P>
int optimizeLevels(Batcher &B, const Opt &o, const Tree &t, P *bl, int64_t firstRow, int64_t nRows) {
    const int64_t M = t.maxnode;
    std::vector<int32_t> height((size_t) M, 0), depth((size_t) M, 0);
    int32_t H = 0, D = 0;
    for (int64_t node : t.order) {
        int32_t h = 0;
        for (int k = 0; k < t.nChild[node]; k++) h = std::max(h, height[t.child[3 * node + k]] + 1);
        height[node] = h; H = std::max(H, h);
    }
    for (auto it = t.order.rbegin(); it != t.order.rend(); ++it)
        if (*it != t.root) { depth[*it] = depth[t.parent[*it]] + 1; D = std::max(D, depth[*it]); }
    // up-profile rows: internal non-root node -> firstRow + k
    std::vector<int64_t> up((size_t) M, -1);
    int64_t nUp = 0;
    for (int64_t node : t.order) if (t.nChild[node] > 0 && node != t.root) up[node] = firstRow + nUp++;
    const int64_t nTmp = nRows - nUp;
    if (nTmp < 1) return VFT_ENOMEM;
    std::vector<std::vector<int64_t>> byDepth((size_t) D + 1), byHeight((size_t) H + 1);
    for (int64_t node : t.order) if (t.nChild[node] > 0) { byHeight[height[node]].push_back(node); if (node != t.root) byDepth[depth[node]].push_back(node); }
    std::vector<int64_t> po, pa, pb;
    std::vector<double> l1, l2;
    for (int32_t d = 1; d <= D; d++) {                        // top-down: a node's up-profile needs its parent's
        po.clear(); pa.clear(); pb.clear(); l1.clear(); l2.clear();
        for (int64_t node : byDepth[d]) {
            const int64_t par = t.parent[node];
            int64_t c, dd, dRow;
            if (par == t.root) {
                int64_t sibs[2], ns = 0;
                for (int j = 0; j < 3; j++) if (t.child[3 * par + j] != node) sibs[ns++] = t.child[3 * par + j];
                c = sibs[0]; dd = sibs[1]; dRow = dd;
            } else { c = t.sibling(node); dd = par; dRow = up[par]; }
            po.push_back(up[node]); pa.push_back(c); pb.push_back(dRow); l1.push_back((double) bl[c]); l2.push_back((double) bl[dd]);
        }
        if (po.empty()) continue;
        int rc = vft_posterior_profile_batch(B.ctx, (int64_t) po.size(), po.data(), pa.data(), pb.data(), l1.data(), l2.data());
        if (rc != VFT_OK) return rc;
        B.stats.posteriorCalls++; B.stats.posteriorItems += (int64_t) po.size();
    }
    // bottom-up wavefront: a node starts as soon as its internal children are done (their profiles rebuilt); the result
    // is the same as level by level -- a node depends on nothing else -- but tasks of different levels overlap, so the
    // rounds stay wide while the stragglers of a level finish
    struct Live { std::unique_ptr<NodeJob> nj; Task<int> task; };
    std::vector<int32_t> pending((size_t) M, 0);
    std::vector<int64_t> ready;
    for (int64_t node : t.order) {
        if (t.nChild[node] == 0) continue;
        for (int k = 0; k < t.nChild[node]; k++) pending[node] += t.nChild[t.child[3 * node + k]] > 0;
        if (pending[node] == 0) ready.push_back(node);
    }
    std::reverse(ready.begin(), ready.end());                  // popped from the back: post-order first
    RowPool tmpRows(firstRow + nUp, nTmp);
    std::vector<Live> live;
    for (;;) {
        while (!ready.empty()) {
            const int64_t row = tmpRows.take();
            if (row < 0) break;
            const int64_t node = ready.back();
            ready.pop_back();
            std::unique_ptr<NodeJob> nj(new NodeJob{node, StarJob{}});
            nj->job.row = row;
            Task<int> task = nodeTask<P>(B, o, t, bl, up, nj.get());
            task.h.resume();
            live.push_back(Live{std::move(nj), std::move(task)});
        }
        size_t keep = 0;
        for (size_t k = 0; k < live.size(); k++) {
            if (!live[k].task.done()) { if (keep != k) live[keep] = std::move(live[k]); keep++; continue; }
            const int64_t node = live[k].nj->node;
            tmpRows.give(live[k].nj->job.row);
            if (node != t.root && --pending[t.parent[node]] == 0) ready.push_back(t.parent[node]);
        }
        live.erase(live.begin() + (std::ptrdiff_t) keep, live.end());
        if (live.empty() && ready.empty()) break;
        if (B.waiters.empty()) { if (ready.empty() || tmpRows.freeRows.empty()) return VFT_EINVAL; continue; }
        const int rc = B.step();
        if (rc != VFT_OK) return rc;
    }
    return VFT_OK;
}

bool readOpt(vft_ctx *ctx, const vft_ml_options *in, Opt &o, vft_config &cfg) {
    if (!ctx || !in) return false;
    if (vft_get_config(ctx, &cfg, nullptr) != VFT_OK) return false;
    o.minLen = in->MLMinBranchLength; o.ftol = in->MLFTolBranchLength; o.atol = in->MLMinBranchLengthTolerance;
    o.closeLimit = in->closeLogLkLimit; o.mlAccuracy = in->mlAccuracy; o.fast = in->fastNNI != 0; o.single = cfg.precision == 32;
    return o.minLen > 0 && o.ftol > 0 && o.atol > 0 && o.mlAccuracy >= 1;
}

double rdP(const void *p, int64_t k, bool single) { return single ? (double) ((const float *) p)[k] : ((const double *) p)[k]; }
void wrP(void *p, int64_t k, bool single, double v) { if (single) ((float *) p)[k] = (float) v; else ((double *) p)[k] = v; }

}  // namespace

extern "C" void vft_ml_default_options(int32_t precision, vft_ml_options *o) {
    const bool dbl = precision == 64;
    o->MLMinBranchLength = dbl ? 5.0e-9 : 5.0e-4;                // Constants.h:30-31
    o->MLFTolBranchLength = 0.001;                               // Constants.h:27-28
    o->MLMinBranchLengthTolerance = dbl ? 1.0e-9 : 1.0e-4;       // Constants.h:24-25
    o->closeLogLkLimit = 5.0;                                    // Constants.h:40
    o->mlAccuracy = 1;                                           // Options.h:62
    o->fastNNI = 1;                                              // Options.h:58
}

extern "C" int vft_ml_pair_optimize_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *idA, const int64_t *idB,
                                          double *length, double *loglk, vft_ml_stats *stats) {
    Opt o; vft_config cfg;
    if (!readOpt(ctx, opt, o, cfg) || n < 0 || (n > 0 && (!idA || !idB || !length || !loglk))) return VFT_EINVAL;
    Batcher B{ctx};
    std::vector<Task<int>> tasks;
    tasks.reserve((size_t) n);
    for (int64_t k = 0; k < n; k++) tasks.push_back(pairJob(B, o, idA[k], idB[k], &length[k], &loglk[k]));
    const int rc = B.run(tasks);
    if (stats) *stats = B.stats;
    return rc;
}

extern "C" int vft_ml_quartet_nni_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *ids, void *len,
                                        double *criteria, int32_t *choice, int32_t *starTest, int64_t firstScratchRow,
                                        vft_ml_stats *stats) {
    Opt o; vft_config cfg;
    if (!readOpt(ctx, opt, o, cfg) || n < 0 || (n > 0 && (!ids || !len || !criteria || !choice))) return VFT_EINVAL;
    if (firstScratchRow < 2 * cfg.nSeqs || firstScratchRow + 3 * n > 2 * cfg.nSeqs + cfg.nScratch) return VFT_EINVAL;
    Batcher B{ctx};
    std::vector<QuartetJob> jobs((size_t) n);
    std::vector<Task<int>> tasks;
    tasks.reserve((size_t) n);
    for (int64_t k = 0; k < n; k++) {
        QuartetJob &j = jobs[(size_t) k];
        for (int i = 0; i < 4; i++) j.q[i] = ids[4 * k + i];
        for (int i = 0; i < 5; i++) j.len[i] = rdP(len, 5 * k + i, o.single);
        for (int i = 0; i < 3; i++) { j.criteria[i] = criteria[3 * k + i]; j.rows[i] = firstScratchRow + 3 * k + i; }
        tasks.push_back(quartetNNI(B, o, &j));
    }
    const int rc = B.run(tasks);
    if (rc == VFT_OK)
        for (int64_t k = 0; k < n; k++) {
            const QuartetJob &j = jobs[(size_t) k];
            for (int i = 0; i < 5; i++) wrP(len, 5 * k + i, o.single, j.len[i]);
            for (int i = 0; i < 3; i++) criteria[3 * k + i] = j.criteria[i];
            choice[k] = j.choice;
            if (starTest) starTest[k] = j.star;
        }
    if (stats) *stats = B.stats;
    return rc;
}

// chooseNNI, NJ.tcc:4836-4852: the minimum-evolution counterpart of MLQuartetNNI.  The six profile distances of every
// quartet (correctedPairDistances, NJ.tcc:1460-1488) are ONE vft_dist_pairs call for the whole batch; the pseudo-count
// prior, the log correction (logCorrect, NJ.tcc:322-330) and the three-way comparison are scalar work on the results.
extern "C" int vft_choose_nni_batch(vft_ctx *ctx, int64_t n, const int64_t *ids, double pseudoWeight, int32_t logdist,
                                    double *criteria, int32_t *choice) {
    vft_config cfg;
    if (!ctx || n < 0 || (n > 0 && (!ids || !criteria || !choice)) || vft_get_config(ctx, &cfg, nullptr) != VFT_OK) return VFT_EINVAL;
    if (n == 0) return VFT_OK;
    const bool single = cfg.precision == 32;
    std::vector<int64_t> pi((size_t) (6 * n)), pj((size_t) (6 * n));
    for (int64_t k = 0; k < n; k++)
        for (int h = 0, i = 0; i < 4; i++)
            for (int j = i + 1; j < 4; j++, h++) { pi[(size_t) (6 * k + h)] = ids[4 * k + i]; pj[(size_t) (6 * k + h)] = ids[4 * k + j]; }   // qAB qAC qAD qBC qBD qCD
    std::vector<double> raw((size_t) (12 * n));                  // room for 6n distances + 6n weights of either precision
    void *dist = raw.data(), *weight = (char *) raw.data() + (size_t) (6 * n) * (single ? 4 : 8);
    const int rc = vft_dist_pairs(ctx, pi.data(), pj.data(), 6 * n, VFT_PAIRS_PROFILE_RAW, dist, weight);
    if (rc != VFT_OK) return rc;
    const bool jukesCantor = cfg.nCodes == 4 && !cfg.useMatrix;
    for (int64_t k = 0; k < n; k++) {
        double d[6];
        for (int h = 0; h < 6; h++) d[h] = rdP(dist, 6 * k + h, single);
        if (pseudoWeight > 0) {                                  // :1472-1484
            double dTop = 0, dBottom = 0;
            for (int h = 0; h < 6; h++) {
                // hit.dist * hit.weight is a numeric_t product
                const double prod = single ? (double) (((const float *) dist)[6 * k + h] * ((const float *) weight)[6 * k + h])
                                           : ((const double *) dist)[6 * k + h] * ((const double *) weight)[6 * k + h];
                dTop += prod;
                dBottom += rdP(weight, 6 * k + h, single);
            }
            const double prior = dBottom > 0.01 ? dTop / dBottom : 3.0;
            for (int h = 0; h < 6; h++) {
                const double w = rdP(weight, 6 * k + h, single);
                d[h] = (d[h] * w + prior * pseudoWeight) / (w + pseudoWeight);
            }
        }
        if (logdist)                                             // logCorrect, :322-330
            for (int h = 0; h < 6; h++) {
                const double maxscore = 3.0;
                double x = d[h];
                if (jukesCantor) x = x < 0.74 ? -0.75 * std::log(1.0 - x * 4.0 / 3.0) : maxscore;
                else x = x < 0.99 ? -1.3 * std::log(1.0 - x) : maxscore;
                d[h] = x < maxscore ? x : maxscore;
            }
        double *c = criteria + 3 * k;
        c[0] = d[0] + d[5]; c[1] = d[1] + d[4]; c[2] = d[2] + d[3];        // AB+CD, AC+BD, AD+BC (:4843-4845)
        choice[k] = 0;
        if (c[1] < c[0] && c[1] <= c[2]) choice[k] = 1;
        else if (c[2] < c[0] && c[2] <= c[1]) choice[k] = 2;
    }
    return VFT_OK;
}

extern "C" int vft_ml_split_test_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *ids, const void *len,
                                      double *loglk, double *siteLk, int32_t *choice, int32_t *badSplit, int64_t firstScratchRow,
                                      vft_ml_stats *stats) {
    Opt o; vft_config cfg;
    if (!readOpt(ctx, opt, o, cfg) || n < 0 || (n > 0 && (!ids || !len || !loglk || !siteLk || !choice || !badSplit))) return VFT_EINVAL;
    if (firstScratchRow < 2 * cfg.nSeqs || firstScratchRow + 3 * n > 2 * cfg.nSeqs + cfg.nScratch) return VFT_EINVAL;
    Batcher B{ctx};
    B.L = cfg.nPos;
    std::vector<SplitJob> jobs((size_t) n);
    std::vector<Task<int>> tasks;
    tasks.reserve((size_t) n);
    for (int64_t k = 0; k < n; k++) {
        SplitJob &j = jobs[(size_t) k];
        for (int i = 0; i < 4; i++) j.q[i] = ids[4 * k + i];
        for (int i = 0; i < 5; i++) j.len[i] = rdP(len, 5 * k + i, o.single);
        for (int i = 0; i < 3; i++) j.rows[i] = firstScratchRow + 3 * k + i;
        j.site = siteLk + (size_t) k * 3 * (size_t) cfg.nPos;
        tasks.push_back(splitTest(B, o, &j));
    }
    const int rc = B.run(tasks);
    if (rc == VFT_OK)
        for (int64_t k = 0; k < n; k++) {
            for (int i = 0; i < 3; i++) loglk[3 * k + i] = jobs[(size_t) k].loglk[i];
            choice[k] = jobs[(size_t) k].choice; badSplit[k] = jobs[(size_t) k].bad;
        }
    if (stats) *stats = B.stats;
    return rc;
}

// testSplitsML, NJ.tcc:6800-6855 + traverseTestSplitsML :6856-7000 (no constraints): the splits do not change the tree, so
// every internal node's test is independent once the up-profiles exist.  Up-profiles of the whole tree first (top-down, one
// posterior batch per depth: the values getUpProfile would build lazily, NJ.tcc:3382-3434), then the split tests in lock-step
// chunks, then SHSupport for the chunk.  support[node] as the reference sets it (:6991): 0 for a bad split.
template<typename P>
int testSplits(vft_ctx *ctx, const Opt &o, const vft_config &cfg, const Tree &t, const P *bl, int64_t nBootstrap, const int64_t *col,
               P *support, int64_t *nBadSplits, vft_ml_stats *stats) {
    const int64_t M = t.maxnode, first = 2 * cfg.nSeqs, L = cfg.nPos;
    std::vector<int32_t> depth((size_t) M, 0);
    int32_t D = 0;
    for (auto it = t.order.rbegin(); it != t.order.rend(); ++it)
        if (*it != t.root) { depth[*it] = depth[t.parent[*it]] + 1; D = std::max(D, depth[*it]); }
    std::vector<int64_t> up((size_t) M, -1);
    int64_t nUp = 0;
    for (int64_t node : t.order) if (t.nChild[node] > 0 && node != t.root) up[node] = first + nUp++;
    const int64_t CH = (cfg.nScratch - nUp) / 3;
    if (CH < 1) return VFT_ENOMEM;
    Batcher B{ctx};
    B.L = L;
    std::vector<std::vector<int64_t>> byDepth((size_t) D + 1);
    for (int64_t node : t.order) if (t.nChild[node] > 0 && node != t.root) byDepth[depth[node]].push_back(node);
    auto rootSibs = [&](int64_t node, int64_t sibs[2]) {
        int ns = 0;
        for (int j = 0; j < 3; j++) if (t.child[3 * t.root + j] != node) sibs[ns++] = t.child[3 * t.root + j];
    };
    std::vector<int64_t> po, pa, pb;
    std::vector<double> l1, l2;
    for (int32_t d = 1; d <= D; d++) {
        po.clear(); pa.clear(); pb.clear(); l1.clear(); l2.clear();
        for (int64_t node : byDepth[d]) {
            const int64_t par = t.parent[node];
            int64_t c, dd, dRow;
            if (par == t.root) { int64_t sibs[2]; rootSibs(node, sibs); c = sibs[0]; dd = sibs[1]; dRow = dd; }
            else { c = t.sibling(node); dd = par; dRow = up[par]; }
            po.push_back(up[node]); pa.push_back(c); pb.push_back(dRow); l1.push_back((double) bl[c]); l2.push_back((double) bl[dd]);
        }
        if (po.empty()) continue;
        const int rc = vft_posterior_profile_batch(ctx, (int64_t) po.size(), po.data(), pa.data(), pb.data(), l1.data(), l2.data());
        if (rc != VFT_OK) return rc;
        B.stats.posteriorCalls++; B.stats.posteriorItems += (int64_t) po.size();
    }
    // the splits, in the reference's post-order (only the order of the outputs depends on it)
    std::vector<int64_t> nodes;
    for (int64_t node : t.order) if (t.nChild[node] > 0 && node != t.root) nodes.push_back(node);
    for (int64_t i = 0; i < M; i++) support[i] = (P) -1.0;                                   // NJ.tcc: support.resize(maxnodes, -1)
    *nBadSplits = 0;
    std::vector<SplitJob> jobs;
    std::vector<double> site, lk3, sup;
    for (size_t k0 = 0; k0 < nodes.size(); k0 += (size_t) CH) {
        const size_t n = std::min(nodes.size() - k0, (size_t) CH);
        jobs.assign(n, SplitJob{});
        site.assign(n * 3 * (size_t) L, 0.0);
        std::vector<Task<int>> tasks;
        tasks.reserve(n);
        for (size_t k = 0; k < n; k++) {
            const int64_t node = nodes[k0 + k], par = t.parent[node];
            SplitJob &j = jobs[k];
            j.q[0] = t.child[3 * node]; j.q[1] = t.child[3 * node + 1];                         // setupABCD, NJ.tcc:1942-1974
            int64_t lenDNode;
            if (par == t.root) { int64_t sibs[2]; rootSibs(node, sibs); j.q[2] = sibs[0]; j.q[3] = sibs[1]; lenDNode = sibs[1]; }
            else { j.q[2] = t.sibling(node); j.q[3] = up[par]; lenDNode = par; }
            j.len[0] = (double) bl[j.q[0]]; j.len[1] = (double) bl[j.q[1]]; j.len[2] = (double) bl[j.q[2]];
            j.len[3] = (double) bl[lenDNode]; j.len[4] = (double) bl[node];                     // :6886-6890
            for (int i = 0; i < 3; i++) j.rows[i] = first + nUp + 3 * (int64_t) k + i;
            j.site = site.data() + k * 3 * (size_t) L;
            tasks.push_back(splitTest(B, o, &j));
        }
        int rc = B.run(tasks);
        if (rc != VFT_OK) return rc;
        lk3.resize(3 * n); sup.resize(n);
        for (size_t k = 0; k < n; k++) for (int i = 0; i < 3; i++) lk3[3 * k + i] = jobs[k].loglk[i];
        if (nBootstrap > 0) {
            rc = vft_sh_support_batch(ctx, (int64_t) n, nBootstrap, col, lk3.data(), site.data(), sup.data());
            if (rc != VFT_OK) return rc;
        }
        for (size_t k = 0; k < n; k++) {
            *nBadSplits += jobs[k].bad;
            if (nBootstrap > 0) support[nodes[k0 + k]] = (P) (jobs[k].bad ? 0.0 : sup[k]);       // :6991
        }
    }
    if (stats) *stats = B.stats;
    return VFT_OK;
}

extern "C" int vft_ml_test_splits(vft_ctx *ctx, const vft_ml_options *opt, int64_t root, int64_t maxnode, const int32_t *nChild,
                                  const int64_t *child, const void *branchlength, int64_t nBootstrap, const int64_t *col, void *support,
                                  int64_t *nBadSplits, vft_ml_stats *stats) {
    Opt o; vft_config cfg;
    if (!readOpt(ctx, opt, o, cfg) || !nChild || !child || !branchlength || !support || !nBadSplits || root < 0 || root >= maxnode
        || maxnode > 2 * cfg.nSeqs || nBootstrap < 0 || (nBootstrap > 0 && !col) || cfg.nSeqs < 4)
        return VFT_EINVAL;
    Tree t{root, maxnode, cfg.nSeqs, nChild, child, {}, {}};
    const int rc = t.build();
    if (rc != VFT_OK) return rc;
    if (cfg.precision == 32) return testSplits<float>(ctx, o, cfg, t, (const float *) branchlength, nBootstrap, col, (float *) support, nBadSplits, stats);
    return testSplits<double>(ctx, o, cfg, t, (const double *) branchlength, nBootstrap, col, (double *) support, nBadSplits, stats);
}

extern "C" int vft_ml_star_optimize_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *ids, void *len,
                                          int64_t firstScratchRow, vft_ml_stats *stats) {
    Opt o; vft_config cfg;
    if (!readOpt(ctx, opt, o, cfg) || n < 0 || (n > 0 && (!ids || !len))) return VFT_EINVAL;
    if (firstScratchRow < 2 * cfg.nSeqs || firstScratchRow + n > 2 * cfg.nSeqs + cfg.nScratch) return VFT_EINVAL;
    Batcher B{ctx};
    std::vector<StarJob> jobs((size_t) n);
    std::vector<Task<int>> tasks;
    tasks.reserve((size_t) n);
    for (int64_t k = 0; k < n; k++) {
        for (int i = 0; i < 3; i++) { jobs[(size_t) k].p[i] = ids[3 * k + i]; jobs[(size_t) k].len[i] = rdP(len, 3 * k + i, o.single); }
        jobs[(size_t) k].row = firstScratchRow + k;
        tasks.push_back(starOptimize(B, o, &jobs[(size_t) k]));
    }
    const int rc = B.run(tasks);
    if (rc == VFT_OK)
        for (int64_t k = 0; k < n; k++) for (int i = 0; i < 3; i++) wrP(len, 3 * k + i, o.single, jobs[(size_t) k].len[i]);
    if (stats) *stats = B.stats;
    return rc;
}

extern "C" int vft_ml_optimize_branch_lengths(vft_ctx *ctx, const vft_ml_options *opt, int64_t root, int64_t maxnode, const int32_t *nChild,
                                              const int64_t *child, void *branchlength, int32_t schedule, vft_ml_stats *stats) {
    Opt o; vft_config cfg;
    if (!readOpt(ctx, opt, o, cfg) || !nChild || !child || !branchlength || root < 0 || root >= maxnode || maxnode > 2 * cfg.nSeqs) return VFT_EINVAL;
    if (cfg.nSeqs < 3) return VFT_EINVAL;                        // the 2-leaf case (NJ.tcc:5070-5079) is one vft_ml_pair_optimize_batch item
    Tree t{root, maxnode, cfg.nSeqs, nChild, child, {}, {}};
    int rc = t.build();
    if (rc != VFT_OK) return rc;
    Batcher B{ctx};
    const int64_t first = 2 * cfg.nSeqs;
    if (schedule == VFT_ML_SCHEDULE_LEVELS) {
        rc = cfg.precision == 32 ? optimizeLevels<float>(B, o, t, (float *) branchlength, first, cfg.nScratch)
                                 : optimizeLevels<double>(B, o, t, (double *) branchlength, first, cfg.nScratch);
    } else if (schedule == VFT_ML_SCHEDULE_REFERENCE) {
        RowPool pool(first, cfg.nScratch);
        std::vector<Task<int>> tasks;
        if (cfg.precision == 32) tasks.push_back(optimizeSequential<float>(B, o, t, (float *) branchlength, pool));
        else tasks.push_back(optimizeSequential<double>(B, o, t, (double *) branchlength, pool));
        rc = B.run(tasks);
        if (rc == VFT_OK) rc = tasks[0].h.promise().value;
    } else rc = VFT_EINVAL;
    if (stats) *stats = B.stats;
    return rc;
}
