"""Whole-tree likelihood sweeps at scale: NJ tree of a synthetic aa alignment (vft_nj_build), then
recomputeMLProfiles + treeLogLk under JTT + 4 CAT rates through vft_tree_loglk (level-synchronous batches).
argv: taxa columns"""
import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import replay
from veryfasttree_b200 import api, synth

N = int(sys.argv[1]); L = int(sys.argv[2])
chars = synth.make_alignment(N, L, 'aa', 1)
chars = chars[synth.unique_rows(chars)]
codes = api.encode(chars, 'aa')
N = codes.shape[0]
z = np.load('tests/golden/blosum45_f32.npz')
lib = api.load()
tree = api.nj_build(codes, 20, 32, lib=lib, tables=[z['distances'], z['eigenval'], z['eigentot'], z['codeFreq']], trace=False)
print('NJ tree of %d x %d aa: %.2f s end to end' % (N, L, tree.stats['secondsEndToEnd']), flush=True)
jtt = replay.read_refdump(os.path.join(replay.GOLDEN, 'ml_jtt_f32_e3.mldump.bin'))      # model constants (JTT92) from the reference
cfg = api.make_config(N, L, 20, 32)
cfg.reserved = 1
dt = np.float32
with api.Context(lib, cfg) as ctx:
    ctx.upload_leaves(codes)
    d = lib.dll
    arrs = [np.ascontiguousarray(jtt[k], dtype=dt) for k in ("ml.codeFreq", "ml.eigenval", "ml.eigeninv", "ml.eigeninvT", "ml.statinv")]
    lib.check(d.vft_upload_transmat(ctx.h, *[api._ptr(a) for a in arrs]), "vft_upload_transmat")
    rates = np.array([0.25, 0.8, 1.0, 2.6], dtype=dt)
    ratecat = ((np.arange(L) * 7 + np.arange(L) // 3) % 4).astype(np.int64)
    lib.check(d.vft_sync_rates(ctx.h, api._ptr(rates), 4, api._ptr(ratecat), float(jtt["ml.minlen"][0]), float(jtt["ml.minlen"][1]), 3), "vft_sync_rates")
    n_child = tree.n_child; child = tree.child; bl = tree.branchlength
    for rep in range(2):
        c0 = ctx.counters()
        t0 = time.time()
        lk, _ = ctx.tree_loglk(tree.root, n_child[:tree.maxnode], child[:tree.maxnode], bl[:tree.maxnode], recompute=True)
        t1 = time.time()
        lk2, _ = ctx.tree_loglk(tree.root, n_child[:tree.maxnode], child[:tree.maxnode], bl[:tree.maxnode], recompute=False)
        t2 = time.time()
        c1 = ctx.counters()
        ms = list(c1.msKernel); ms0 = list(c0.msKernel); nk = list(c1.nKernel); nk0 = list(c0.nKernel)
        print('rep %d: recomputeMLProfiles+treeLogLk %.3f s, treeLogLk alone %.3f s, logLk %.6f (%.6f)' % (rep, t1 - t0, t2 - t1, lk, lk2))
        for nm, a, b, x, y in zip(api.KERNEL_NAMES, ms, ms0, nk, nk0):
            if x - y:
                print('   %-22s %6d launches %9.2f ms' % (nm, x - y, a - b))
    prof_bytes = L * (20 * 4 + 4 + 1)
    print('   posterior sweep: %d internal profiles, %.2f GB in+out algorithmic; pairLogLk sweep: %.2f GB' % (N - 3, 3 * (N - 3) * prof_bytes / 1e9, 2 * (N - 2) * prof_bytes / 1e9))
