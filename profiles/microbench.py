"""Latency / throughput microbenchmarks of the kernel-level ABI on a warm device (run under gpurun).
Prints wall time per call (host clock, includes launch + synchronisation) and device time per call
(CUDA events, profile mode) for the request shapes the NJ driver produces."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api, synth

N, L = int(sys.argv[1]) if len(sys.argv) > 1 else 16000, 200
kind = sys.argv[2] if len(sys.argv) > 2 else 'nt'
if kind == 'aa':
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 1287
chars = synth.make_alignment(N, L, kind, 1)
chars = chars[synth.unique_rows(chars)]
N = chars.shape[0]
codes = api.encode(chars, kind)
lib = api.load()
tables = None
if kind == 'aa':
    z = np.load('tests/golden/blosum45_f32.npz')
    tables = [z['distances'], z['eigenval'], z['eigentot'], z['codeFreq']]

def run(profile):
    cfg = api.make_config(N, L, 4 if kind == 'nt' else 20, 32, use_matrix=tables is not None)
    cfg.reserved = 1 if profile else 0
    out = {}
    with api.Context(lib, cfg) as ctx:
        if tables: ctx.upload_tables(*tables)
        ctx.upload_leaves(codes)
        ctx.outprofile_rebuild()
        ctx.out_distance_all(N, 0.0)
        rs = np.random.RandomState(3)
        nj = min(N // 2 - 2, 6000)
        active = list(range(N))
        t0 = time.perf_counter()
        for k in range(nj):
            a = active.pop(rs.randint(len(active))); b = active.pop(rs.randint(len(active)))
            ctx.profile_average(N + k, a, b, -1.0, 0.001)
            ctx.outprofile_update(a, b, N + k, N - k)
            active.append(N + k)
        ctx.get_self(N)
        out['average+update per join (wall us)'] = (time.perf_counter() - t0) / nj * 1e6
        nA = N - nj
        internal = np.arange(N, N + nj)
        def timeit(f, reps):
            f(); c0 = ctx.counters()
            t0 = time.perf_counter()
            for _ in range(reps): f()
            wall = (time.perf_counter() - t0) / reps * 1e6
            c1 = ctx.counters()
            dev = (c1.msDist - c0.msDist + c1.msSelect - c0.msSelect + c1.msProfile - c0.msProfile) / reps * 1e3
            byt = (c1.algoBytes - c0.algoBytes) / reps
            return {'wall_us': round(wall, 2), 'dev_us': round(dev, 2), 'algoMB': round(byt / 1e6, 3),
                    'GBps_dev': round(byt / dev / 1e3, 1) if dev > 0 else None}
        for n in (1, 16, 64, 256, 1024, 8192, 65536):
            ids = internal[rs.randint(0, nj, size=n)]
            out['out_distance_batch n=%d' % n] = timeit(lambda: ctx.out_distance_batch(ids, nA, 1.0), 200 if n <= 1024 else 20)
        for n in (1, 16, 64, 256, 1024, 8192, 65536, 524288):
            pi = np.full(n, internal[5]); pj = internal[rs.randint(0, nj, size=n)]
            out['dist_pairs internal n=%d' % n] = timeit(lambda: ctx.dist_pairs(pi, pj), 200 if n <= 1024 else 10)
        for n in (256, 8192, 65536):
            pi = rs.randint(0, N, size=n); pj = rs.randint(0, N, size=n)
            out['dist_pairs leaf n=%d' % n] = timeit(lambda: ctx.dist_pairs(pi, pj), 50)
        ctx.out_distance_all(nA, 1.0)
        out['out_distance_all'] = timeit(lambda: ctx.out_distance_all(nA, 1.0), 20)
        q = int(internal[-1])
        K = 2 * int(0.5 + np.sqrt(N))
        out['one_vs_all internal query K=%d' % K] = timeit(lambda: ctx.dist_one_vs_all(q, nA, K), 20)
        ql = int(active[0]) if active[0] < N else 0
        out['one_vs_all leaf query'] = timeit(lambda: ctx.dist_one_vs_all(ql, nA, K), 20)
        out['outprofile_rebuild'] = timeit(lambda: ctx.outprofile_rebuild(), 5)
    return out

res = {'N': N, 'L': L, 'kind': kind, 'plain': run(False), 'profiled': run(True)}
for k in res['plain']:
    print("%-40s plain %-60s profiled %s" % (k, json.dumps(res['plain'][k]), json.dumps(res['profiled'][k])))
json.dump(res, open('gpurun_out/microbench_%s_%d.json' % (kind, N), 'w'), indent=1)
