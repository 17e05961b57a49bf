"""Calls the kernel-level ABI with the request shapes of the NJ driver, a few times each; run under
`ncu --metrics gpu__time_duration.sum` to get true per-kernel durations (no event/launch overhead)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
kind = sys.argv[2] if len(sys.argv) > 2 else 'nt'
L = int(sys.argv[3]) if len(sys.argv) > 3 else (200 if kind == 'nt' else 1287)
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
with api.Context(lib, cfg) as ctx:
    if tables: ctx.upload_tables(*tables)
    ctx.upload_leaves(codes)
    ctx.outprofile_rebuild()
    ctx.out_distance_all(N, 0.0)
    rs = np.random.RandomState(3)
    nj = min(N // 2 - 2, 3000)
    active = list(range(N))
    for k in range(nj):
        a = active.pop(rs.randint(len(active))); b = active.pop(rs.randint(len(active)))
        ctx.profile_average_update(N + k, a, b, N - k, -1.0, 0.001)
        active.append(N + k)
    nA = N - nj
    internal = np.arange(N, N + nj)
    print('MARK begin')
    for n in (1, 16, 256, 1024, 65536):
        ids = internal[rs.randint(0, nj, size=n)]
        for _ in range(3): ctx.out_distance_batch(ids, nA, 1.0)
        pi = np.full(n, internal[5]); pj = internal[rs.randint(0, nj, size=n)]
        for _ in range(3): ctx.dist_pairs(pi, pj)
    ctx.out_distance_all(nA, 1.0); ctx.out_distance_all(nA, 1.0)
    K = 2 * int(0.5 + np.sqrt(N))
    for _ in range(2): ctx.dist_one_vs_all(int(internal[-1]), nA, K)
    for _ in range(2): ctx.dist_one_vs_all(int(active[0]) if active[0] < N else 0, nA, K)
    ctx.outprofile_rebuild()
