"""Branch-length optimisation at scale: NJ tree of a synthetic aa alignment (vft_nj_build), then
rounds of optimizeAllBranchLengths under JTT + CAT through vft_ml_optimize_branch_lengths in the level-synchronous
schedule (lock-step Brent over k_pair_loglk / k_posterior), next to the unmodified reference binary run with
`-mllen -nni 0 -spr 0` on the same alignment (same NJ tree: the NJ phase is bit-identical), whose per-round times are
read from its own "N rounds ML lengths ... Time" lines.
argv: taxa columns [reference threads (0 = skip the reference)] [nRateCats]"""
import os, re, subprocess, sys, tempfile, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import replay
from veryfasttree_b200 import api, synth

N = int(sys.argv[1]); L = int(sys.argv[2])
ref_threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n_cat = int(sys.argv[4]) if len(sys.argv) > 4 else 20
chars = synth.make_alignment(N, L, 'aa', 1)
chars = chars[synth.unique_rows(chars)]
codes = api.encode(chars, 'aa')
N = codes.shape[0]
z = np.load('tests/golden/blosum45_f32.npz')
lib = api.load(os.environ.get('VFT_LIB'))          # (VFT_LIB: dry runs of this script over the CPU double, no timing value)
tree = api.nj_build(codes, 20, 32, lib=lib, tables=[z['distances'], z['eigenval'], z['eigentot'], z['codeFreq']], trace=False)
print('NJ tree of %d x %d aa: %.2f s end to end' % (N, L, tree.stats['secondsEndToEnd']), flush=True)
jtt = replay.read_refdump(os.path.join(replay.GOLDEN, 'ml_jtt_f32_e0.mldump.bin'))      # model constants (JTT92) from the reference
cfg = api.make_config(N, L, 20, 32, n_scratch=2 * N)
cfg.reserved = 1
dt = np.float32
M = tree.maxnode
n_child = tree.n_child[:M]; child = tree.child[:M]; bl = tree.branchlength[:M].copy()
min_rel, min_br = float(jtt["ml.minlen"][0]), float(jtt["ml.minlen"][1])
with api.Context(lib, cfg) as ctx:
    ctx.upload_leaves(codes)
    d = lib.dll
    arrs = [np.ascontiguousarray(jtt[k], dtype=dt) for k in ("ml.codeFreq", "ml.eigenval", "ml.eigeninv", "ml.eigeninvT", "ml.statinv")]
    lib.check(d.vft_upload_transmat(ctx.h, *[api._ptr(a) for a in arrs]), "vft_upload_transmat")
    one = np.array([1.0], dtype=dt); zeros = np.zeros(L, dtype=np.int64)
    lib.check(d.vft_sync_rates(ctx.h, api._ptr(one), 1, api._ptr(zeros), min_rel, min_br, 0), "vft_sync_rates")
    opt = ctx.ml_options()
    lk, _ = ctx.tree_loglk(tree.root, n_child, child, bl, recompute=True)
    print('start: LogLk = %.3f' % lk, flush=True)
    for rnd in range(1, 5):                                    # VeryFastTreeImpl.tcc:262-308: rounds of ML lengths, CAT after the first
        c0 = ctx.counters()
        t0 = time.time()
        bl, st = ctx.ml_optimize_branch_lengths(opt, tree.root, n_child, child, bl, schedule=1)
        t1 = time.time()
        lk, _ = ctx.tree_loglk(tree.root, n_child, child, bl, recompute=False)
        t2 = time.time()
        c1 = ctx.counters()
        print('%d rounds ML lengths (level schedule): LogLk = %.3f  sweep %.3f s + treeLogLk %.3f s; %d lock-step rounds, %d pairLogLk items in %d calls (%.0f per call), %d posteriors in %d calls'
              % (rnd, lk, t1 - t0, t2 - t1, st['rounds'], st['loglkItems'], st['loglkCalls'], st['loglkItems'] / max(1, st['loglkCalls']), st['posteriorItems'], st['posteriorCalls']), flush=True)
        for nm, a, b, x, y in zip(api.KERNEL_NAMES, list(c1.msKernel), list(c0.msKernel), list(c1.nKernel), list(c0.nKernel)):
            if x - y:
                print('   %-22s %6d launches %9.2f ms' % (nm, x - y, a - b))
        if rnd == 1:
            t0 = time.time()
            rates, ratecat, _ = ctx.set_ml_rates(tree.root, n_child, child, bl, n_cat, min_rel, min_br, 0)
            lk, _ = ctx.tree_loglk(tree.root, n_child, child, bl, recompute=False)
            print('   setMLRates(%d categories): %.3f s, LogLk = %.3f' % (n_cat, time.time() - t0, lk), flush=True)

    # MLQuartetNNI throughput: one quartet per internal node -- A, B = its children, C = its sibling,
    # D = the parent's sibling (or another child of the root): read-only node profiles, temporaries in scratch rows, so
    # any set of quartets is a valid batch.  (The real NNI round passes the parent's up-profile as D; same cost.)
    par = np.full(M, -1, dtype=np.int64)
    for nd in range(M):
        for k in range(n_child[nd]):
            par[child[nd, k]] = nd
    def sib(x):
        p = par[x]
        return [c for c in child[p, :n_child[p]] if c != x]
    quartets, lens = [], []
    for nd in range(N, M):
        if nd == tree.root or n_child[nd] != 2 or par[nd] < 0:
            continue
        p = par[nd]
        c = sib(nd)[0]
        d_ = sib(p)[0] if p != tree.root else sib(nd)[-1]
        if d_ == c:
            continue
        a, b = child[nd, 0], child[nd, 1]
        quartets.append([a, b, c, d_]); lens.append([bl[a], bl[b], bl[c], bl[d_], bl[nd]])
    quartets = np.array(quartets, dtype=np.int64); lens = np.array(lens, dtype=dt)
    CH = (2 * N) // 3                                          # three scratch rows per quartet
    tot = dict(rounds=0, loglkItems=0, loglkCalls=0, posteriorItems=0, posteriorCalls=0)
    n_star = n_swap = 0
    c0 = ctx.counters()
    t0 = time.time()
    for q0 in range(0, len(quartets), CH):
        ln, crit, choice, star, st = ctx.ml_quartet_nni(opt, quartets[q0:q0 + CH], lens[q0:q0 + CH], 2 * N)
        for k in tot: tot[k] += st[k]
        n_star += int(star.sum()); n_swap += int((choice != 0).sum())
    dtq = time.time() - t0
    c1 = ctx.counters()
    print('MLQuartetNNI: %d quartets in %.3f s = %.1f us per quartet (%d per call); %d star tests, %d would swap; %d lock-step rounds, %d pairLogLk items in %d calls (%.0f per call), %d posteriors'
          % (len(quartets), dtq, 1e6 * dtq / len(quartets), CH, n_star, n_swap, tot['rounds'], tot['loglkItems'], tot['loglkCalls'], tot['loglkItems'] / max(1, tot['loglkCalls']), tot['posteriorItems']), flush=True)
    for nm, a, b, x, y in zip(api.KERNEL_NAMES, list(c1.msKernel), list(c0.msKernel), list(c1.nKernel), list(c0.nKernel)):
        if x - y:
            print('   %-22s %6d launches %9.2f ms' % (nm, x - y, a - b))
    # SH-like supports of every internal split (testSplitsML): up-profiles + split tests + SHSupport, 1 000 resamples
    rng = np.random.default_rng(7)
    col = rng.integers(0, L, size=(1000, L), dtype=np.int64)
    t0 = time.time()
    sup, nbad, st = ctx.ml_test_splits(opt, tree.root, n_child, child, bl, col)
    dts = time.time() - t0
    print('testSplitsML: %d splits in %.3f s (%d bad); %d lock-step rounds, %d pairLogLk items in %d calls, %d posteriors; support quartiles %s'
          % (int((sup >= 0).sum()), dts, nbad, st['rounds'], st['loglkItems'], st['loglkCalls'], st['posteriorItems'],
             np.round(np.quantile(sup[sup >= 0], [0.25, 0.5, 0.75]), 2)), flush=True)
if ref_threads > 0 and os.path.exists(replay.REF_BIN):
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, 'a.fa')
        synth.write_fasta(fa, chars)
        cmd = [replay.REF_BIN, '-threads', str(ref_threads), '-ext', 'AVX2', '-mllen', '-nni', '0', '-spr', '0', '-nosupport', '-cat', str(n_cat),
               '-out', os.path.join(td, 'a.tree'), fa]
        t0 = time.time()
        out = subprocess.run(cmd, capture_output=True, text=True)
        print('reference (unmodified, AVX2 build, -threads %d): %.1f s wall' % (ref_threads, time.time() - t0))
        last = None
        for line in (out.stdout + out.stderr).splitlines():
            m = re.search(r'Initial topology in ([0-9.]+) seconds', line)
            if m:
                last = float(m.group(1)); print('   ' + line.strip())
            m = re.search(r'(\d+) rounds ML lengths: LogLk = ([-0-9.]+).*Time ([0-9.]+)', line)
            if m:
                t = float(m.group(3))
                print('   %s   [+%.2f s since the previous stamp]' % (line.strip(), t - (last if last is not None else 0.0)))
                last = t
            if 'rate categor' in line.lower() or 'Switched to using' in line:
                print('   ' + line.strip())
