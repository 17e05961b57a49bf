// TEST INFRASTRUCTURE ONLY.
// Probes the tie behaviour of the reference's psort() (src/Utils.h:126-146 ->
// boost::sort::parallel_stable_sort -> spinsort for nthread<2 or n<65536,
// libs/boost-sort/.../parallel_stable_sort.hpp:131-136) when it is fed the
// reference's non-strict comparators (NeighbourJoining.tcc:7285-7311: they
// return true on ties).  Conjecture checked here: the result is always
//     ascending by key, ties in REVERSE original order
// i.e. identical to a strict sort on (key asc, original index desc).
// Exit code 0 iff the conjecture held for every trial.
#include "Utils.h"
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <random>

struct E { double key; int64_t idx; };
struct LeCmp { bool operator()(const E &a, const E &b) const { return !(a.key > b.key); } };

static int check(std::vector<E> v, const char *what) {
    std::vector<E> want = v;
    std::sort(want.begin(), want.end(), [](const E &a, const E &b) {
        return a.key != b.key ? a.key < b.key : a.idx > b.idx; });
    veryfasttree::psort(v.begin(), v.end(), LeCmp());
    for (size_t i = 0; i < v.size(); i++)
        if (v[i].idx != want[i].idx) {
            std::fprintf(stderr, "MISMATCH %s n=%zu at %zu: got idx %ld key %g want idx %ld key %g\n",
                         what, v.size(), i, (long) v[i].idx, v[i].key, (long) want[i].idx, want[i].key);
            return 1;
        }
    return 0;
}

int main(int argc, char **argv) {
    std::mt19937_64 rng(12345);
    int bad = 0, trials = 0;
    std::vector<size_t> sizes = {0, 1, 2, 3, 5, 16, 31, 32, 33, 64, 71, 72, 73, 74, 100, 127, 128, 129, 250, 500, 1000,
                                 1023, 1024, 1025, 1100, 2047, 2048, 2049, 3000, 4096, 5000, 8191, 10000, 16000,
                                 32000, 65535, 65536, 65537, 100000, 200000};
    for (size_t n : sizes) {
        for (int pat = 0; pat < 12; pat++) {
            for (int rep = 0; rep < (n <= 5000 ? 6 : 2); rep++) {
                std::vector<E> v(n);
                int64_t nkeys = (pat % 4 == 0) ? 3 : (pat % 4 == 1) ? 17 : (pat % 4 == 2) ? (int64_t) (n / 4 + 1) : (int64_t) (4 * n + 1);
                for (size_t i = 0; i < n; i++) v[i] = {(double) (rng() % nkeys), (int64_t) i};
                const char *what = "random";
                if (pat >= 4 && pat < 6) {          // already non-decreasing (with ties)
                    std::stable_sort(v.begin(), v.end(), [](const E &a, const E &b) { return a.key < b.key; });
                    for (size_t i = 0; i < n; i++) v[i].idx = i;
                    what = "sorted";
                } else if (pat >= 6 && pat < 8) {   // non-increasing (with ties)
                    std::stable_sort(v.begin(), v.end(), [](const E &a, const E &b) { return a.key > b.key; });
                    for (size_t i = 0; i < n; i++) v[i].idx = i;
                    what = "reverse";
                } else if (pat == 8) {              // all equal (the 1e20 sentinel case)
                    for (size_t i = 0; i < n; i++) v[i].key = 1e20;
                    what = "allequal";
                } else if (pat == 9) {              // sorted head + random tail
                    std::stable_sort(v.begin(), v.begin() + n * 7 / 8, [](const E &a, const E &b) { return a.key < b.key; });
                    for (size_t i = 0; i < n; i++) v[i].idx = i;
                    what = "head-sorted";
                } else if (pat == 10) {             // reverse head + random tail
                    std::stable_sort(v.begin(), v.begin() + n * 15 / 16, [](const E &a, const E &b) { return a.key > b.key; });
                    for (size_t i = 0; i < n; i++) v[i].idx = i;
                    what = "head-reverse";
                } else if (pat == 11) {             // many sentinels + few real values (refresh-phase shape)
                    for (size_t i = 0; i < n; i++) if (rng() % 3) v[i].key = 1e20;
                    what = "sentinels";
                }
                bad += check(v, what);
                trials++;
            }
        }
    }
    std::printf("psort_probe: %d trials, %d mismatches\n", trials, bad);
    return bad ? 1 : 0;
}
