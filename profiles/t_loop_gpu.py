"""Device-resident join loop on the GPU: trees vs the goldens / the host-driven loop, and timings.  argv: [quick]"""
import sys, os, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import replay
from veryfasttree_b200 import api, synth
lib = api.load()
def tables_for(kind, prec):
    if kind != "aa": return None
    z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
    return [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]
KEYS = ("nOutSingleFetch", "nPairSingleFetch", "nPairPrefetchHit", "nRefreshTopHits", "nVisibleUpdate", "nHillBetter", "nDeviceCalls")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for name in (["nt60", "c1"] if quick else ["nt60", "aa60", "c1", "aa300", "nt1000"]):
    for prec in (32, 64):
        chars, kind = replay.golden_case(name)
        want = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
        tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=lib, tables=tables_for(kind, prec), device_loop=1)
        got = tree.newick(["t%d" % i for i in range(chars.shape[0])])
        print(name, prec, "OK" if got == want else "DIFF", {k: tree.stats[k] for k in KEYS}, flush=True)
if quick:
    sys.exit(0)
for kind, n, L, seed in (("nt", 16000, 200, 1), ("aa", 8000, 1287, 1), ("aa", 20000, 1287, 1)):
    chars = synth.make_alignment(n, L, kind, seed); chars = chars[synth.unique_rows(chars)]
    codes = api.encode(chars, kind)
    tabs = tables_for(kind, 32)
    A = 4 if kind == "nt" else 20
    res = {}
    for mode in (0, 1, 1):
        tr = api.nj_build(codes, A, 32, lib=lib, tables=tabs, device_loop=mode)
        res[mode] = tr
        st = tr.stats
        print("%s %d x %d device_loop=%d: device %.3f s, end-to-end %.3f s (leaf %.2f s, joins %.2f s, in calls %.2f s) launches %d %s" % (
            kind, codes.shape[0], L, mode, st["deviceMsResident"] / 1e3, st["secondsEndToEnd"], st["secondsLeafTopHits"], st["secondsJoins"], st["secondsInCalls"],
            st["counters"]["launches"], {k: st[k] for k in KEYS}), flush=True)
    a, b = res[0], res[1]
    same = np.array_equal(a.joins, b.joins) and a.branchlength.tobytes() == b.branchlength.tobytes() and np.array_equal(a.parent, b.parent)
    print("   device loop vs host loop:", "IDENTICAL" if same else "DIFFERENT", flush=True)
    if not same:
        d = np.nonzero((a.joins != b.joins).any(axis=1))[0]
        print("   first differing join", d[:3], a.joins[d[0]] if len(d) else None, b.joins[d[0]] if len(d) else None)
