"""GPU parity tests (B200): the CUDA product library, called through the C-ABI, against
  (1) the white-box vectors of the reference's own templates (tests/golden/*.refdump.bin),
  (2) the CPU restatement (oracle/) on seeded inputs, every output array bit-identical,
  (3) the NJ trees / top-hit lists of the unmodified reference binary (tests/golden/*.nj.tree),
  (4) size-independent properties at the bench size.
Bit-exact means bit-exact: float payloads are compared as bytes, indices as integers.
"""
import os

import numpy as np
import pytest

import replay
from veryfasttree_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def glib():
    lib = api.load()                      # fails loudly when the CUDA library is missing
    assert lib.backend == "cuda-sm100a"
    return lib


@pytest.fixture(scope="module")
def olib():
    replay.ensure_oracle_built()
    return api.load(replay.ORACLE_LIB)


def tables_for(kind, prec):
    if kind != "aa":
        return None
    z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
    return [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["nt60", "aa60"])
def test_cuda_matches_reference_templates(glib, name, prec):
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d.refdump.bin" % (name, prec)))
    chars, kind = replay.golden_case(name)
    assert replay.replay(glib, dump, chars, kind, prec) == []


def run_script(lib, codes, n_codes, prec, tables, seed, n_joins, k_top):
    """A seeded sequence of ABI calls; returns every output for comparison."""
    rs = np.random.RandomState(seed)
    N, L = codes.shape
    out = {}
    cfg = api.make_config(N, L, n_codes, prec, use_matrix=tables is not None)
    with api.Context(lib, cfg) as ctx:
        if tables is not None:
            ctx.upload_tables(*tables)
        ctx.upload_leaves(codes)
        ctx.outprofile_rebuild()
        out["od0"] = ctx.out_distance_all(N, 0.0)
        for q in rs.randint(0, N, size=3):
            j, d, w, c = ctx.dist_one_vs_all(int(q), N, k_top)
            out["ova0_%d" % q] = [j, d, w, c]
        pi, pj = rs.randint(0, N, size=500), rs.randint(0, N, size=500)
        out["leafpairs"] = list(ctx.dist_pairs(pi, pj))
        active = list(range(N))
        diam = np.zeros(2 * N, dtype=ctx.dt)
        totdiam = 0.0
        n_active = N
        for k in range(n_joins):
            a, b = rs.choice(len(active), size=2, replace=False)
            a, b = active[a], active[b]
            nw = N + k
            dm = float(ctx.dt(0.001 * (1 + k % 7)))
            if k % 2:
                ctx.profile_average_update(nw, a, b, n_active, -1.0, dm)      # the fused join launch
            else:
                ctx.profile_average(nw, a, b, -1.0, dm)
                ctx.outprofile_update(a, b, nw, n_active)
            diam[nw] = dm
            totdiam += float(ctx.dt(ctx.dt(diam[nw] - diam[a]) - diam[b]))
            active.remove(a); active.remove(b); active.append(nw)
            n_active -= 1
            if k % 16 == 5:
                ctx.outprofile_rebuild()
        out["self"] = [ctx.get_self(N + k) for k in range(n_joins)]
        out["outprofile"] = list(ctx.get_profile(-1))
        out["lastprofile"] = list(ctx.get_profile(N + n_joins - 1))
        ids = np.array(sorted(active))[::3]
        out["od_batch"] = ctx.out_distance_batch(ids, n_active, totdiam)
        out["od_all"] = ctx.out_distance_all(n_active, totdiam)
        for q in [N + n_joins - 1, active[0], active[len(active) // 2]]:
            j, d, w, c = ctx.dist_one_vs_all(int(q), n_active, k_top)
            out["ova1_%d" % q] = [j, d, w, c]
        act = np.array(active)
        pi, pj = act[rs.randint(0, len(act), size=800)], act[rs.randint(0, len(act), size=800)]
        out["pairs_join"] = list(ctx.dist_pairs(pi, pj, 0))
        out["pairs_raw"] = list(ctx.dist_pairs(pi, pj, 1))
        # the list-merging pass of a top-hits refresh (vft_tophits_merge): own lists with stale entries (negative
        # distance), entries without an ancestor (-1), self hits and duplicates of the transferred hits
        q = N + n_joins - 1
        m = max(4, k_top // 2)
        aj, ad, aw, ac = ctx.dist_one_vs_all(q, n_active, 2 * m)
        lists = [int(x) for x in aj[:m]]
        if q not in lists:
            lists[-1] = q                                   # the iNode == newnode case keeps the transferred distances
        own_off, own_j, own_d = [0], [], []
        for ln in lists:
            k_own = int(rs.randint(0, m + 1))
            js = act[rs.randint(0, len(act), size=k_own)].astype(np.int64)
            ds = rs.random_sample(k_own).astype(ctx.dt)
            ds[rs.random_sample(k_own) < 0.3] = -1e20       # ancestor changed: distance unknown
            js[rs.random_sample(k_own) < 0.1] = -1          # no active ancestor
            if k_own > 2:
                js[0] = ln                                  # self
                js[1] = aj[min(3, len(aj) - 1)]             # duplicate of a transferred hit
            own_j += list(js); own_d += list(ds); own_off.append(len(own_j))
        out["merge"] = list(ctx.tophits_merge(q, n_active, m, lists, own_off, own_j, np.array(own_d, dtype=ctx.dt), aj, ad))
        cn = ctx.counters()
        out["ops"] = np.array([cn.seqOps, cn.profileOps, cn.outprofileOps, cn.profileAvgOps])
    return out


def flatten(o, pfx=""):
    if isinstance(o, dict):
        for k, v in o.items():
            yield from flatten(v, pfx + "/" + str(k))
    elif isinstance(o, (list, tuple)):
        for k, v in enumerate(o):
            yield from flatten(v, pfx + "/" + str(k))
    else:
        yield pfx, np.asarray(o)


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("kind,n,L,joins", [("nt", 700, 333, 300), ("aa", 400, 207, 150), ("nt", 5000, 200, 600),
                                            ("aa", 300, 610, 120), ("nt", 400, 1100, 150)])   # long rows: the CTA-per-pair kernel
def test_cuda_bit_exact_vs_oracle(glib, olib, kind, n, L, joins, prec):
    chars = synth.make_alignment(n, L, kind, seed=100 + n)
    chars[::17, ::5] = ord("-")           # ragged gaps, also in nt
    chars[3] = ord("-")                   # an all-gap sequence: weight 0.01 / dist 1 sentinels
    codes = api.encode(chars, kind)
    A = 4 if kind == "nt" else 20
    k_top = 2 * int(0.5 + np.sqrt(n))
    a = run_script(glib, codes, A, prec, tables_for(kind, prec), 7, joins, k_top)
    b = run_script(olib, codes, A, prec, tables_for(kind, prec), 7, joins, k_top)
    bad = [ka for (ka, va), (kb, vb) in zip(flatten(a), flatten(b)) if not replay.bits_equal(va, vb)]
    assert bad == []


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["nt60", "aa60", "c1", "aa300", "nt1000"])
def test_nj_tree_identical_to_reference(glib, name, prec):
    chars, kind = replay.golden_case(name)
    want = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
    tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=glib,
                        tables=tables_for(kind, prec))
    assert tree.newick(["t%d" % i for i in range(chars.shape[0])]) == want
    assert tree.stats["counters"]["launches"] > 0


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["c1", "aa300"])
def test_leaf_top_hits_identical_to_reference(glib, name, prec):
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d.tophits.bin" % (name, prec)))
    chars, kind = replay.golden_case(name)
    tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=glib,
                        tables=tables_for(kind, prec))
    assert np.array_equal(tree.leaf_top_hits, dump["tophits.j"])


def test_bench_size_properties_and_reference(glib, tmp_path):
    """BASELINE.json configs[1] shape scaled to what the reference finishes in ~10 s here (4000 taxa):
    same tree as the reference binary when it travelled with the snapshot; always: a valid binary
    tree over all leaves, every join between active nodes, one-vs-all sorted and self on top."""
    n, L = 4000, 200
    chars = synth.make_alignment(n, L, "nt", seed=1)
    chars = chars[synth.unique_rows(chars)]
    n = chars.shape[0]
    codes = api.encode(chars, "nt")
    tree = api.nj_build(codes, 4, 32, lib=glib)
    assert tree.root == 2 * n - 3 and tree.maxnode == 2 * n - 2
    assert (tree.parent[:tree.root] >= 0).all() and tree.parent[tree.root] == -1
    assert (np.bincount(tree.parent[:tree.root], minlength=2 * n)[n:tree.root] == 2).all()
    seen = np.zeros(2 * n, dtype=bool)
    for k, (a, b) in enumerate(tree.joins):
        assert not seen[a] and not seen[b] and max(a, b) < n + k
        seen[a] = seen[b] = True
    if os.path.exists(replay.REF_BIN):
        import sys
        sys.path.insert(0, replay.GOLDEN)
        import make_golden
        fa = str(tmp_path / "b.fa")
        synth.write_fasta(fa, chars)
        assert tree.newick(["t%d" % i for i in range(n)]) == make_golden.ref_tree(fa, "nt", 32)


def test_no_device_argument_errors(glib):
    cfg = api.make_config(10, 20, 4, 32, device=99)
    with pytest.raises(api.VftError):
        api.Context(glib, cfg)


ML_PARAMS = [(name, prec, lvl) for name, case in replay.ML_CASES.items() for prec, lvl in case[5]]


@pytest.mark.parametrize("name,prec,lvl", ML_PARAMS)
def test_cuda_likelihood_matches_reference(glib, name, prec, lvl):
    """k_pair_loglk / k_posterior vs the reference's own pairLogLk / posteriorProfile (tests/golden/*.mldump.bin).
    Bit-identical posterior profiles and per-site likelihoods wherever the reference makes no libm call
    (-fastexp 2/3 with a transition matrix); otherwise, and for the final log(), the north_star tolerance:
    1e-5 relative (fp32) / 1e-12 (fp64)."""
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d_e%d.mldump.bin" % (name, prec, lvl)))
    chars, kind, model = replay.ml_case_chars(name)
    bad, rel = replay.replay_ml(glib, dump, chars, kind, prec, exact_log=False)
    assert bad == [], (bad, rel)


@pytest.mark.parametrize("name,prec,lvl", ML_PARAMS)
def test_cuda_ml_optimizers_match_reference(glib, name, prec, lvl):
    """SURVEY 8a rows a15-a17 on the device: MLPairOptimize, MLQuartetNNI (MLQuartetOptimize / onedimenmin / brent), the
    per-node body of optimizeAllBranchLengths and the whole sweep as lock-step batches over k_pair_loglk / k_posterior,
    vs the reference's own functions (tests/golden/*.mldump.bin).  The device's exp/log differ from libm in the last
    bits, which can move a Brent iterate: optimised lengths within 5x Brent's own fractional tolerance (1e-3),
    log-likelihoods within the north_star tolerance (1e-5 fp32 / 1e-10 fp64 relative), NNI choices identical except
    between criteria closer than that tolerance."""
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d_e%d.mldump.bin" % (name, prec, lvl)))
    chars, kind, model = replay.ml_case_chars(name)
    bad, info = replay.replay_ml_opt(glib, dump, chars, kind, prec, exact=False)
    print(info)
    assert bad == [], (bad, info)


@pytest.mark.parametrize("name,prec", [("c1", 32), ("aa300", 32), ("aa60", 64)])
def test_bionj_tree_identical_to_reference_on_device(glib, name, prec):
    """-bionj on the device: bare profileDist against the out-profile (vft_dist_pairs RAW with j = -1), weighted
    k_average; the reference's -bionj tree byte for byte."""
    from test_oracle_golden import _bionj_tree
    want = open(os.path.join(replay.GOLDEN, "%s_f%d.bionj.tree" % (name, prec))).read().strip()
    assert _bionj_tree(glib, name, prec) == want


def test_speculative_join_on_device(glib, monkeypatch):
    """vft_spec_join_* on the device (shadow out-profile, raw distances, state commit): with the speculative path on, the
    1000-taxon tree is still the reference's, byte for byte."""
    monkeypatch.setenv("VFT_SPECULATION", "1")
    chars, kind = replay.golden_case("nt1000")
    want = open(os.path.join(replay.GOLDEN, "nt1000_f32.nj.tree")).read().strip()
    tree = api.nj_build(api.encode(chars, kind), 4, 32, lib=glib)
    assert tree.newick(["t%d" % i for i in range(chars.shape[0])]) == want
    assert tree.stats["nSpecHit"] > 500


def test_second_device_and_host_threads(glib):
    """A context on device 1 (when there is one): every entry point binds the calling thread to the context's device,
    and the host-thread regions of the driver make no device call of their own (bench.py --gpus N, rank > 0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    chars, kind = replay.golden_case("nt1000")
    want = open(os.path.join(replay.GOLDEN, "nt1000_f32.nj.tree")).read().strip()
    tree = api.nj_build(api.encode(chars, kind), 4, 32, lib=glib, device=1, host_threads=8)
    assert tree.newick(["t%d" % i for i in range(chars.shape[0])]) == want


# ---- the device-resident join loop (vft_nj_options.deviceLoop = 1: k_nj_step / k_nj_eval / the device-side rebuild and refresh) ----
@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["nt60", "aa60", "c1", "aa300", "nt1000"])
def test_device_resident_loop_gives_the_reference_tree(glib, name, prec):
    chars, kind = replay.golden_case(name)
    want = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
    tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=glib, tables=tables_for(kind, prec), device_loop=1)
    assert tree.newick(["t%d" % i for i in range(chars.shape[0])]) == want
    assert tree.stats["counters"]["nKernel"] is not None and tree.stats["nDeviceCalls"] < chars.shape[0]      # far fewer host interventions than joins


@pytest.mark.parametrize("kind,n,L", [("nt", 16000, 200), ("aa", 6000, 1287)])
def test_device_resident_loop_identical_to_host_driven_loop_at_size(glib, kind, n, L):
    """Join order, tree and branch lengths of the device-resident loop == the host-driven loop (itself pinned to the reference at
    these shapes by tests/test_gpu_at_size.py), incl. the device-side top-visible rebuilds and refreshes (hundreds per tree)."""
    chars = synth.make_alignment(n, L, kind, seed=1)
    chars = chars[synth.unique_rows(chars)]
    codes = api.encode(chars, kind)
    A = 4 if kind == "nt" else 20
    dev = api.nj_build(codes, A, 32, lib=glib, tables=tables_for(kind, 32), device_loop=1)
    host = api.nj_build(codes, A, 32, lib=glib, tables=tables_for(kind, 32), device_loop=0)
    assert np.array_equal(dev.joins, host.joins) and np.array_equal(dev.parent, host.parent)
    assert dev.branchlength.tobytes() == host.branchlength.tobytes()
    assert dev.stats["nRefreshTopHits"] == host.stats["nRefreshTopHits"] > 100


def test_tma_staged_sweeps_bit_identical(glib, olib, monkeypatch):
    """VFT_STAGING=1: the all-candidate sweeps with the shared profile staged in shared memory by cp.async.bulk (vft_bulk.cuh);
    every output of the seeded script identical to the CPU restatement."""
    monkeypatch.setenv("VFT_STAGING", "1")
    chars = synth.make_alignment(400, 610, "aa", seed=77)
    chars[::13, ::7] = ord("-")
    codes = api.encode(chars, "aa")
    a = run_script(glib, codes, 20, 32, tables_for("aa", 32), 7, 150, 40)
    b = run_script(olib, codes, 20, 32, tables_for("aa", 32), 7, 150, 40)
    bad = [ka for (ka, va), (kb, vb) in zip(flatten(a), flatten(b)) if not replay.bits_equal(va, vb)]
    assert bad == []
