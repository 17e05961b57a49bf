"""GPU parity at the sizes BASELINE.json quotes (B200): the shapes the small committed goldens cannot reach.

The fixtures here would be tens of megabytes, so they are not committed: the unmodified reference
(oracle/_ref/VeryFastTree) and the white-box dumper over its templates (oracle/_ref/refdump) travel with the
snapshot and are run ON THE BOX, at `-threads 1`, on the same seeded alignment, inside the test.

  C2 at size        16 000 x 200 nt fp32: Newick of the NJ phase == the reference's `NJ` log line
  C3 shape          4 000 x 1287 aa fp32 (BLOSUM45): tree == reference `-ext AVX2`, leaf top-hit lists == the
                    reference's own setAllLeafTopHits
  C4 shape          300 x 1287 aa, JTT, CAT with 20 candidate rates, fp64 `-fastexp 2` and fp32 `-fastexp 3`:
                    pairLogLk, posteriorProfile, treeLogLk, setMLRates, the optimisers, one optimizeAllBranchLengths
                    sweep and testSplitsML over the whole tree (tolerances of tests/replay.py)
  chunked launches  vft_posterior_profile_batch beyond the grid.y chunk, k_pair_loglk with rows over 100 KB,
                    vft_sh_support_batch beyond one launch -- against the CPU restatement
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import replay
from veryfasttree_b200 import api, synth

pytestmark = pytest.mark.gpu

sys.path.insert(0, replay.GOLDEN)


@pytest.fixture(scope="module")
def glib():
    lib = api.load()
    assert lib.backend == "cuda-sm100a"
    return lib


@pytest.fixture(scope="module")
def olib():
    replay.ensure_oracle_built()
    return api.load(replay.ORACLE_LIB)


def need_reference():
    if not (os.path.exists(replay.REF_BIN) and os.path.exists(replay.REFDUMP_BIN)):
        pytest.fail("oracle/_ref/{VeryFastTree,refdump} did not travel with the snapshot: build them with __graft_entry__.build() "
                    "where /root/reference exists")


def tables_for(prec):
    z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
    return [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]


def test_c2_full_size_tree_identical_to_reference(glib, tmp_path):
    """BASELINE.json configs[1] at its real size (the bench workload of round 1)."""
    need_reference()
    import make_golden
    chars = synth.make_alignment(16000, 200, "nt", seed=1)
    chars = chars[synth.unique_rows(chars)]
    n = chars.shape[0]
    tree = api.nj_build(api.encode(chars, "nt"), 4, 32, lib=glib)
    fa = str(tmp_path / "c2.fa")
    synth.write_fasta(fa, chars)
    want = make_golden.ref_tree(fa, "nt", 32)
    assert tree.newick(["t%d" % i for i in range(n)]) == want
    assert tree.stats["counters"]["launches"] > 0


def test_c3_shape_tree_and_top_hits_identical_to_reference(glib, tmp_path):
    """1287-column amino-acid alignment, BLOSUM45 distances, fp32: the C3 regime (long rows, matrix mode, the AVX2 lane
    order of the 20-wide dot products) against the reference run with -ext AVX2."""
    need_reference()
    import make_golden
    chars = synth.make_alignment(4000, 1287, "aa", seed=31)
    chars = chars[synth.unique_rows(chars)]
    n = chars.shape[0]
    fa = str(tmp_path / "c3.fa")
    synth.write_fasta(fa, chars)
    tree = api.nj_build(api.encode(chars, "aa"), 20, 32, lib=glib, tables=tables_for(32))
    th = str(tmp_path / "c3.tophits.bin")
    subprocess.run([replay.REFDUMP_BIN, fa, "aa", "32", th, "tophits"], check=True)
    dump = replay.read_refdump(th)
    assert np.array_equal(tree.leaf_top_hits, dump["tophits.j"])
    want = make_golden.ref_tree(fa, "aa", 32)
    assert tree.newick(["t%d" % i for i in range(n)]) == want


@pytest.mark.parametrize("prec,lvl", [(64, 2), (32, 3)])
def test_c4_shape_likelihood_kernels_match_reference(glib, tmp_path, prec, lvl):
    """300 taxa x 1287 aa, JTT, 20 CAT candidate rates: the long-row paths of k_pair_loglk (2-8 warps per item), k_posterior,
    the level-synchronous sweeps, setMLRates with 20 categories, the lock-step optimisers and testSplitsML on the device."""
    need_reference()
    chars = synth.make_alignment(300, 1287, "aa", seed=21)
    chars = chars[synth.unique_rows(chars)]
    fa = str(tmp_path / "ml.fa")
    synth.write_fasta(fa, chars)
    out = str(tmp_path / "ml.bin")
    env = dict(os.environ, VFT_REFDUMP_NCAT="20", VFT_REFDUMP_QSTRIDE="3")
    subprocess.run([replay.REFDUMP_BIN, fa, "aa", str(prec), out, "ml", "jtt", str(lvl)], check=True, env=env)
    dump = replay.read_refdump(out)
    assert len(dump["ml.cat.rates"]) == 20
    bad, rel = replay.replay_ml(glib, dump, chars, "aa", prec, exact_log=False)
    assert bad == [], (bad, rel)
    bad, info = replay.replay_ml_opt(glib, dump, chars, "aa", prec, exact=False)
    print(info)
    assert bad == [], (bad, info)
    assert info["splits.n"] > 250


def _gtr_context(lib, codes, prec, n_scratch, lvl):
    """A context with the GTR model of the small goldens loaded (tables from tests/golden/ml_gtr_*.mldump.bin)."""
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "ml_gtr_f%d_e%d.mldump.bin" % (prec, lvl)))
    dt = api.np_dtype(prec)
    N, L = codes.shape
    ctx = api.Context(lib, api.make_config(N, L, 4, prec, n_scratch=n_scratch))
    ctx.upload_leaves(codes)
    arrs = [np.ascontiguousarray(dump[k], dtype=dt) for k in ("ml.codeFreq", "ml.eigenval", "ml.eigeninv", "ml.eigeninvT", "ml.statinv")]
    lib.check(lib.dll.vft_upload_transmat(ctx.h, *[api._ptr(a) for a in arrs]), "vft_upload_transmat")
    rates = np.array([0.25, 0.8, 1.0, 2.6], dtype=dt)
    ratecat = ((np.arange(L) * 7 + np.arange(L) // 3) % 4).astype(np.int64)
    lib.check(lib.dll.vft_sync_rates(ctx.h, api._ptr(rates), 4, api._ptr(ratecat), float(dump["ml.minlen"][0]), float(dump["ml.minlen"][1]), lvl),
              "vft_sync_rates")
    return ctx


def _posterior_chunk_script(lib, codes, prec, n_items):
    import ctypes as C
    N, L = codes.shape
    rs = np.random.RandomState(5)
    with _gtr_context(lib, codes, prec, n_items, 3 if prec == 32 else 2) as ctx:
        out = 2 * N + np.arange(n_items, dtype=np.int64)
        a = rs.randint(0, N, size=n_items).astype(np.int64)
        b = rs.randint(0, N, size=n_items).astype(np.int64)
        l1 = 0.01 + rs.random_sample(n_items); l2 = 0.01 + rs.random_sample(n_items)
        lib.check(lib.dll.vft_posterior_profile_batch(ctx.h, n_items, api._ptr(out), api._ptr(a), api._ptr(b), api._ptr(l1), api._ptr(l2)),
                  "vft_posterior_profile_batch")
        # a second level on top of the first (dense inputs), still one call
        out2 = out[:1000]
        a2 = out[1000:2000].copy(); b2 = out[n_items - 1000:].copy()
        lib.check(lib.dll.vft_posterior_profile_batch(ctx.h, 1000, api._ptr(out2), api._ptr(a2), api._ptr(b2), api._ptr(l1[:1000].copy()), api._ptr(l2[:1000].copy())),
                  "vft_posterior_profile_batch")
        res = []
        for k in (0, 999, 1000, 32767, 32768, n_items - 1):
            res += list(ctx.get_profile(int(out[k])))
        pi = out[::7].copy(); pj = out[3::7][:len(pi)].copy(); pi = pi[:len(pj)]
        ll = np.zeros(len(pi)); ln = np.ascontiguousarray(0.02 + rs.random_sample(len(pi)))
        lib.check(lib.dll.vft_pair_loglk_batch(ctx.h, api._ptr(pi), api._ptr(pj), api._ptr(ln), len(pi), api._ptr(ll), None), "vft_pair_loglk_batch")
    return res, ll


def test_posterior_batch_beyond_one_launch_matches_oracle(glib, olib):
    """40 000 posteriorProfile items in one call: more than the 32 768 items one launch carries (grid.y chunking)."""
    chars = synth.make_alignment(500, 64, "nt", seed=12)
    chars[::9, ::4] = ord("-")
    codes = api.encode(chars, "nt")
    ga, gl = _posterior_chunk_script(glib, codes, 32, 40000)
    oa, ol = _posterior_chunk_script(olib, codes, 32, 40000)
    assert all(replay.bits_equal(x, y) for x, y in zip(ga, oa))          # -fastexp 3 with a transition matrix: no libm call
    assert np.allclose(gl, ol, rtol=1e-5, atol=0)


@pytest.mark.parametrize("prec", [32, 64])
def test_pair_loglk_very_long_rows_matches_oracle(glib, olib, prec):
    """13 500 columns: one item's per-site buffer is over 100 KB, the kernel falls back to one item per CTA."""
    import ctypes as C
    chars = synth.make_alignment(40, 13500, "nt", seed=4)
    chars[::5, ::11] = ord("-")
    codes = api.encode(chars, "nt")
    N, L = codes.shape
    got = []
    for lib in (glib, olib):
        with _gtr_context(lib, codes, prec, 8, 3 if prec == 32 else 2) as ctx:
            for k in range(6):
                lib.check(lib.dll.vft_posterior_profile(ctx.h, N + k, 2 * k, 2 * k + 1, 0.05 + 0.01 * k, 0.2 - 0.02 * k), "vft_posterior_profile")
            pi = np.array([0, 1, N, N + 1, N + 2, 3, N + 5], dtype=np.int64); pj = np.array([5, N, N + 1, N + 3, 7, 9, N + 4], dtype=np.int64)
            ln = np.array([0.1, 0.03, 0.4, 0.25, 0.01, 1.5, 0.07])
            ll = np.zeros(len(pi)); site = np.zeros((len(pi), L))
            lib.check(lib.dll.vft_pair_loglk_batch(ctx.h, api._ptr(pi), api._ptr(pj), api._ptr(ln), len(pi), api._ptr(ll), api._ptr(site)), "vft_pair_loglk_batch")
            got.append((ll, site, ctx.get_profile(N + 5)))
    (gl, gs, gp), (ol, os_, op) = got
    assert all(replay.bits_equal(x, y) for x, y in zip(gp, op))
    assert replay.bits_equal(gs, os_)                                     # per-site likelihoods: no libm with -fastexp 2/3
    assert np.allclose(gl, ol, rtol=1e-5 if prec == 32 else 1e-12, atol=0)


def test_sh_support_beyond_one_launch_matches_oracle(glib, olib):
    """More quartets than one k_sh_support launch takes at 1287 columns (the 256 MB staging chunk)."""
    L, nq, nboot = 1287, 9100, 24
    rs = np.random.RandomState(8)
    site = np.ascontiguousarray(np.exp(-3.0 * rs.random_sample((nq, 3, L))))
    loglk = np.log(site).sum(axis=2)
    col = rs.randint(0, L, size=(nboot, L)).astype(np.int64)
    codes = api.encode(synth.make_alignment(8, L, "aa", seed=2), "aa")
    out = []
    for lib in (glib, olib):
        with api.Context(lib, api.make_config(8, L, 20, 64)) as ctx:
            ctx.upload_leaves(codes)
            out.append(ctx.sh_support(col, loglk, site))
    assert replay.bits_equal(out[0], out[1])
    assert 0 < out[0].mean() < 1
