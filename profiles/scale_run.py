"""One NJ+TopHits phase at a BASELINE.json configs[2]-like shape (amino acids, BLOSUM45 matrix, fp32) on one
GPU: wall/device time, counters, and -- when small enough for the reference to finish -- the reference's time."""
import sys, time, json, os
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api, synth

N, L = int(sys.argv[1]), int(sys.argv[2])
t0 = time.time()
chars = synth.make_alignment(N, L, "aa", 1)
chars = chars[synth.unique_rows(chars)]
codes = api.encode(chars, "aa")
print("generated", chars.shape, "in %.1fs" % (time.time() - t0), flush=True)
z = np.load('tests/golden/blosum45_f32.npz')
tables = [z['distances'], z['eigenval'], z['eigentot'], z['codeFreq']]
t0 = time.time()
tr = api.nj_build(codes, 20, 32, tables=tables, trace=False, profile=len(sys.argv) > 3)
wall = time.time() - t0
st = tr.stats
out = {"taxa": int(chars.shape[0]), "columns": L, "wall_s": round(wall, 2), "device_ms": round(st["deviceMsResident"], 1),
       "taxa_per_s": round(chars.shape[0] / (st["deviceMsResident"] * 1e-3), 1), "in_calls_s": round(st["secondsInCalls"], 2),
       "host_s": [round(x, 2) for x in st["secondsHost"][:6]], "refreshes": st["nRefreshTopHits"], "seeds": st["nSeeds"],
       "device_calls": st["nDeviceCalls"], "counters": st["counters"]}
print(json.dumps(out))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/scale_aa_%d_%d.json' % (N, L), 'w'), indent=1)
