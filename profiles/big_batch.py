"""One large candidate batch through vft_dist_pairs (the refresh-sized sweeps of C3+): the bandwidth-bound
regime of the distance kernel.  Run plain for CUDA-event timing or under `ncu -k regex:k_eval` for metrics."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
kind = sys.argv[2] if len(sys.argv) > 2 else 'nt'
L = int(sys.argv[3]) if len(sys.argv) > 3 else (200 if kind == 'nt' else 1287)
npairs = int(sys.argv[4]) if len(sys.argv) > 4 else 262144
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
chars = synth.make_alignment(N, L, kind, 1)
chars = chars[synth.unique_rows(chars)]
N = chars.shape[0]
codes = api.encode(chars, kind)
lib = api.load()
tables = None
if kind == 'aa':
    z = np.load('tests/golden/blosum45_f32.npz')
    tables = [z['distances'], z['eigenval'], z['eigentot'], z['codeFreq']]
cfg = api.make_config(N, L, 4 if kind == 'nt' else 20, 32, use_matrix=tables is not None)
cfg.reserved = 1
with api.Context(lib, cfg) as ctx:
    if tables: ctx.upload_tables(*tables)
    ctx.upload_leaves(codes)
    ctx.outprofile_rebuild()
    ctx.out_distance_all(N, 0.0)
    rs = np.random.RandomState(3)
    nj = min(N // 2 - 2, 6000)
    active = list(range(N))
    for k in range(nj):
        a = active.pop(rs.randint(len(active))); b = active.pop(rs.randint(len(active)))
        ctx.profile_average_update(N + k, a, b, N - k, -1.0, 0.001)
        active.append(N + k)
    internal = np.arange(N, N + nj)
    m = 128
    # refresh-shaped request: m lists of npairs/m candidates, one query per list
    pi = np.repeat(internal[rs.randint(0, nj, size=m)], npairs // m)
    pj = internal[rs.randint(0, nj, size=npairs)]
    for r in range(reps):
        c0 = ctx.counters()
        ctx.dist_pairs(pi, pj)
        c1 = ctx.counters()
        ms = c1.msDist - c0.msDist
        by = c1.algoBytes - c0.algoBytes
        print('%s N=%d L=%d pairs=%d: %.3f ms, %.1f MB algorithmic, %.1f GB/s' % (kind, N, L, npairs, ms, by / 1e6, by / ms / 1e6))
