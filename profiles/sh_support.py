"""SHSupport throughput: n quartets x 1000 resamples x nPos columns through vft_sh_support_batch.  argv: quartets columns"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from veryfasttree_b200 import api
n, L = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(1)
col = rng.integers(0, L, size=(1000, L), dtype=np.int64)
site = 0.02 + 0.9 * rng.random((n, 3, L))
loglk = np.log(site).sum(axis=2)
lib = api.load()
cfg = api.make_config(4, L, 20, 32)
with api.Context(lib, cfg) as ctx:
    ctx.upload_leaves(np.zeros((4, L), dtype=np.uint8))
    ctx.sh_support(col, loglk[:8], site[:8])
    t0 = time.time()
    sup = ctx.sh_support(col, loglk, site)
    dt = time.time() - t0
    # the reference's loop on one host core, on a slice
    m = min(n, 8)
    t0 = time.time()
    sl = np.log(site[:m])
    ref = np.empty(m)
    for q in range(m):
        r = -loglk[q][None, :] + sl[q][:, col].sum(axis=2).T          # [nBoot, 3] (numpy pairwise sums: timing only)
        best = r.argmax(axis=1)
        rb = r[np.arange(1000), best]
        d = np.minimum(rb - r[np.arange(1000), (best + 1) % 3], rb - r[np.arange(1000), (best + 2) % 3])
        delta = min(loglk[q][0] - loglk[q][1], loglk[q][0] - loglk[q][2])
        ref[q] = (d < delta).mean()
    dtn = (time.time() - t0) / m
    print('%d quartets x 1000 resamples x %d columns: %.3f s = %.1f us per quartet (%.2f G ordered adds/s); numpy on one core: %.0f us per quartet; supports %s' %
          (n, L, dt, 1e6 * dt / n, 3e3 * n * L / dt / 1e9, 1e6 * dtn, np.round(sup[:6], 3)))
