"""Large-configuration runs of the NJ+TopHits phase (BASELINE.json configs[2]-like: amino acid, BLOSUM45 distances,
fp32) on one B200, with the per-kernel device-time table, and -- optionally -- the unmodified reference on a
subsample for a CPU taxa/s figure on the same host.  argv: taxa columns [ref_taxa]"""
import json, os, re, subprocess, sys, tempfile, time
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api, synth

N = int(sys.argv[1]); L = int(sys.argv[2])
ref_taxa = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
chars = synth.make_alignment(N, L, 'aa', 1)
chars = chars[synth.unique_rows(chars)]
codes = api.encode(chars, 'aa')
z = np.load('tests/golden/blosum45_f32.npz')
tables = [z['distances'], z['eigenval'], z['eigentot'], z['codeFreq']]
print('alignment %d x %d aa built in %.1f s' % (codes.shape[0], L, time.time() - t0), flush=True)
lib = api.load()
for rep, profile in (((0, False),) if os.environ.get("VFT_ONE_RUN") else ((0, False), (1, True))):
    t0 = time.time()
    tr = api.nj_build(codes, 20, 32, lib=lib, tables=tables, trace=False, profile=profile)
    st = tr.stats
    print('run %d (profile=%s): device %.2f s, end-to-end %.2f s, %.0f taxa/s; leaf phase %.2f s, joins %.2f s; in ABI calls %.2f s; host %s'
          % (rep, profile, st['deviceMsResident'] / 1e3, st['secondsEndToEnd'], codes.shape[0] / st['secondsEndToEnd'], st['secondsLeafTopHits'],
             st['secondsJoins'], st['secondsInCalls'], [round(x, 2) for x in st['secondsHost'][:6]]), flush=True)
    c = st['counters']
    print('   seqOps %d profileOps %d refreshes %d launches %d algorithmic GB %.1f' % (c['seqOps'], c['profileOps'], st['nRefreshTopHits'], c['launches'], c['algoBytes'] / 1e9))
    if profile:
        for nm, ms, cnt, by in zip(api.KERNEL_NAMES, c['msKernel'], c['nKernel'], c['bytesKernel']):
            if cnt:
                print('   %-24s %8d launches %10.1f ms %10.1f us/launch%s' % (nm, cnt, ms, 1e3 * ms / cnt, '   %8.1f GB algorithmic = %5.0f GB/s' % (by / 1e9, by / ms / 1e6) if by else ''))
        print('   distance kernels: %.1f GB algorithmic in %.2f s = %.0f GB/s' % (c['distBytes'] / 1e9, c['msDist'] / 1e3, c['distBytes'] / c['msDist'] / 1e6))
if ref_taxa:
    sub = chars[:ref_taxa]
    ref = os.path.join('oracle', '_ref', 'VeryFastTree')
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, 'a.fa'); synth.write_fasta(fa, sub)
        th = os.cpu_count()
        for threads in (1, th):
            t0 = time.time()
            p = subprocess.run([ref, '-threads', str(threads), '-noml', '-nni', '0', '-spr', '0', '-nosupport', '-log', os.path.join(td, 'log'), fa],
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
            log = open(os.path.join(td, 'log')).read()
            m = re.search(r'Initial topology in ([0-9.]+) seconds', log)
            print('reference (unmodified, AVX2 build) on the first %d taxa, -threads %d: initial topology %s s (wall %.1f s)' % (ref_taxa, threads, m.group(1) if m else '?', time.time() - t0), flush=True)
