"""Reference (unmodified binary, -mavx2 build) on the C3-shaped bench workload: NJ+TopHits phase time for a few
sizes and thread counts, with -ext AVX2 -fastexp 3 (what BASELINE.json's metric names), plus whether the NJ tree
of a multi-threaded run equals the product's.  argv: sizes (comma) threads (comma)"""
import os, re, subprocess, sys, tempfile, time
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api, synth

sizes = [int(x) for x in sys.argv[1].split(',')]
threads = [int(x) for x in sys.argv[2].split(',')]
L = 1287
z = np.load('tests/golden/blosum45_f32.npz')
tables = [z['distances'], z['eigenval'], z['eigentot'], z['codeFreq']]
lib = api.load()
ref = os.path.join('oracle', '_ref', 'VeryFastTree')
print('host cores', os.cpu_count(), flush=True)
for n in sizes:
    chars = synth.make_alignment(n, L, 'aa', 1)
    chars = chars[synth.unique_rows(chars)]
    codes = api.encode(chars, 'aa')
    tr = api.nj_build(codes, 20, 32, lib=lib, tables=tables, trace=False)
    tr = api.nj_build(codes, 20, 32, lib=lib, tables=tables, trace=False)
    mine = tr.newick(['t%d' % i for i in range(codes.shape[0])])
    print('n=%d: product %.2f s end to end' % (codes.shape[0], tr.stats['secondsEndToEnd']), flush=True)
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, 'a.fa'); synth.write_fasta(fa, chars)
        for th in threads:
            if th > (os.cpu_count() or 1):
                continue
            t0 = time.time()
            p = subprocess.run([ref, '-ext', 'AVX2', '-fastexp', '3', '-threads', str(th), '-noml', '-nni', '0', '-spr', '0', '-nosupport', '-log', os.path.join(td, 'log'), fa],
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(th)))
            log = open(os.path.join(td, 'log')).read()
            m = re.search(r'Initial topology in ([0-9.]+) seconds', log)
            u = re.search(r'([0-9.]+) seconds: Identified unique sequences', p.stderr.replace('\r', '\n'))
            nj = [l.split('\t', 1)[1].strip() for l in log.splitlines() if l.startswith('NJ\t')]
            print('   reference -ext AVX2 -fastexp 3 -threads %d: initial topology %s s (unique stamp %s, wall %.1f s); NJ tree identical to the product: %s'
                  % (th, m.group(1) if m else '?', u.group(1) if u else '?', time.time() - t0, bool(nj) and nj[0] == mine), flush=True)
