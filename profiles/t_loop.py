import sys, os, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import replay
from veryfasttree_b200 import api, synth
olib = api.load(replay.ORACLE_LIB)
def tables_for(kind, prec):
    if kind != "aa": return None
    z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
    return [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]
names = sys.argv[1:] or ["nt60","aa60","c1","aa300","nt1000"]
for name in names:
    for prec in (32,64):
        chars, kind = replay.golden_case(name)
        want = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
        t=time.time()
        tree = api.nj_build(api.encode(chars, kind), 4 if kind=="nt" else 20, prec, lib=olib, tables=tables_for(kind,prec), device_loop=1)
        got = tree.newick(["t%d" % i for i in range(chars.shape[0])])
        ref = api.nj_build(api.encode(chars, kind), 4 if kind=="nt" else 20, prec, lib=olib, tables=tables_for(kind,prec), device_loop=0)
        same_joins = np.array_equal(tree.joins, ref.joins)
        first = -1
        if not same_joins:
            d = np.nonzero((tree.joins != ref.joins).any(axis=1))[0]
            first = int(d[0]) if len(d) else -1
        print(name, prec, "OK" if got==want else "DIFF", "joins same" if same_joins else "first diff at join %d of %d: %s vs %s" % (first, len(ref.joins), tree.joins[first], ref.joins[first]), round(time.time()-t,2), {k:tree.stats[k] for k in ("nOutSingleFetch","nPairSingleFetch","nPairPrefetchHit","nRefreshTopHits","nVisibleUpdate","nHillBetter")})
