"""Leaf-phase one-vs-all (seqDist of one leaf against every leaf + criterion + top-K): the K1 sweep of SURVEY 8d.
argv: taxa columns kind(nt|aa) [reps]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api, synth

N, L, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
chars = synth.make_alignment(N, L, kind, 1)
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
    K = 2 * int(0.5 + np.sqrt(N))
    ctx.dist_one_vs_all(0, N, K)
    c0 = ctx.counters()
    for r in range(reps):
        ctx.dist_one_vs_all((r * 7919) % N, N, K)
    c1 = ctx.counters()
    ms = (c1.msKernel[2] - c0.msKernel[2]) / reps
    sel = (c1.msKernel[4] - c0.msKernel[4]) / reps
    by = (c1.bytesKernel[2] - c0.bytesKernel[2]) / reps
    print('%s %d x %d: k_one_vs_all_leaf %.3f ms per sweep (%.1f MB algorithmic, %.0f GB/s), k_topk_select %.3f ms (K=%d)' % (kind, N, L, ms, by / 1e6, by / ms / 1e6, sel, K))
