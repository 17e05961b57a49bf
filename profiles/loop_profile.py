"""Per-kernel device time of one NJ+TopHits step with the device-resident join loop (CUDA events around every launch).
argv: kind taxa columns [device_loop]"""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import replay
from veryfasttree_b200 import api, synth
kind, n, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
lib = api.load()
chars = synth.make_alignment(n, L, kind, 1); chars = chars[synth.unique_rows(chars)]
codes = api.encode(chars, kind)
tabs = None
if kind == "aa":
    z = np.load(os.path.join(replay.GOLDEN, "blosum45_f32.npz")); tabs = [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]
A = 4 if kind == "nt" else 20
for profile in (False, True):
    tr = api.nj_build(codes, A, 32, lib=lib, tables=tabs, device_loop=mode, profile=profile, trace=False)
    st = tr.stats; c = st["counters"]
    import zlib
    print("%s %d x %d device_loop=%d profile=%s: device %.3f s, end-to-end %.3f s (leaf %.2f, joins %.2f, in calls %.2f) launches %d refreshes %d tree crc %08x" % (
        kind, codes.shape[0], L, mode, profile, st["deviceMsResident"] / 1e3, st["secondsEndToEnd"], st["secondsLeafTopHits"], st["secondsJoins"], st["secondsInCalls"], c["launches"], st["nRefreshTopHits"],
        zlib.crc32(tr.parent.tobytes() + tr.branchlength.tobytes())), flush=True)
    if profile:
        for nm, ms, cnt, by in zip(api.KERNEL_NAMES, c["msKernel"], c["nKernel"], c["bytesKernel"]):
            if cnt:
                print("   %-24s %8d events %10.1f ms %10.1f us/event%s" % (nm, cnt, ms, 1e3 * ms / cnt, "   %8.1f GB algorithmic = %5.0f GB/s" % (by / 1e9, by / ms / 1e6) if by else ""))
