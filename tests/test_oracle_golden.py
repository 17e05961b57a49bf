"""CPU tests (no GPU): pin the oracle restatement and the host NJ driver to the reference.

  - oracle/vft_oracle.c replays the white-box script that oracle/refdump.cpp ran on the
    reference's own templates: every array bit-identical (nt / aa, fp32 / fp64);
  - the product's host driver (nj_host.cpp) on top of the oracle reproduces the reference
    binary's NJ tree byte for byte and the reference's leaf top-hit index lists exactly;
  - the prefetch hints do not change a single decision.
"""
import os

import numpy as np
import pytest

import replay
from veryfasttree_b200 import api, synth


@pytest.fixture(scope="module")
def olib():
    replay.ensure_oracle_built()
    lib = api.load(replay.ORACLE_LIB)
    assert lib.backend == "oracle-cpu"
    return lib


def tables_for(kind, prec):
    if kind != "aa":
        return None
    z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
    return [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["nt60", "aa60"])
def test_oracle_matches_reference_templates(olib, name, prec):
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d.refdump.bin" % (name, prec)))
    chars, kind = replay.golden_case(name)
    assert replay.replay(olib, dump, chars, kind, prec) == []


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["nt60", "aa60", "c1", "aa300", "nt1000"])
def test_nj_tree_identical_to_reference(olib, name, prec):
    chars, kind = replay.golden_case(name)
    want = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
    tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=olib,
                        tables=tables_for(kind, prec))
    names = ["t%d" % i for i in range(chars.shape[0])]
    assert tree.newick(names) == want
    assert tree.stats["nOutSingleFetch"] == 0 and tree.stats["nPairSingleFetch"] == 0


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["c1", "aa300"])
def test_leaf_top_hits_identical_to_reference(olib, name, prec):
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d.tophits.bin" % (name, prec)))
    chars, kind = replay.golden_case(name)
    tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=olib,
                        tables=tables_for(kind, prec))
    assert tree.m == int(dump["tophits.m"][0])
    assert np.array_equal(tree.leaf_top_hits, dump["tophits.j"])


def test_prefetch_is_only_a_hint(olib):
    chars, kind = replay.golden_case("c1")
    codes = api.encode(chars, kind)
    a = api.nj_build(codes, 4, 32, lib=olib, prefetch=True)
    b = api.nj_build(codes, 4, 32, lib=olib, prefetch=False)
    assert np.array_equal(a.joins, b.joins)
    assert a.branchlength.tobytes() == b.branchlength.tobytes()
    assert b.stats["nOutPrefetchHit"] == 0 and b.stats["nOutSingleFetch"] > 0


def test_tiny_inputs_visible_set_mode(olib, tmp_path):
    # fewer than 13 leaves: top-hits are switched off (m < 4, NJ.tcc:2829) and the visible-set
    # search (fastNJSearch, NJ.tcc:3686) is used.  Compared with the reference binary when it is here.
    import sys
    sys.path.insert(0, replay.GOLDEN)
    from veryfasttree_b200 import synth
    for n in (3, 4, 5, 9, 12, 15, 20):
        chars, kind = replay.golden_case("nt60")
        chars = chars[:n]
        t = api.nj_build(api.encode(chars, kind), 4, 64, lib=olib)
        assert (t.m == 0) == (n < 13)
        assert t.root == 2 * n - 3 and (t.parent[:t.root] >= 0).all()
        if os.path.exists(replay.REF_BIN):
            import make_golden
            fa = str(tmp_path / ("tiny%d.fa" % n))
            synth.write_fasta(fa, chars)
            assert t.newick(["t%d" % i for i in range(n)]) == make_golden.ref_tree(fa, kind, 64)


def test_b200operations_drops_into_reference_templates(tmp_path):
    """oracle/_ref/plugin_probe = the reference's own NeighbourJoining<> compiled over B200Operations<>
    (built only where /root/reference exists): same NJ tree as BasicOperations."""
    import subprocess
    probe = os.path.join(replay.ROOT, "oracle", "_ref", "plugin_probe")
    if not os.path.exists(probe):
        pytest.skip("oracle/_ref/plugin_probe not built on this box")
    from veryfasttree_b200 import synth
    for kind, name in (("nt", "nt60"), ("aa", "aa300")):
        chars, _ = replay.golden_case(name)
        fa = str(tmp_path / (name + ".fa"))
        synth.write_fasta(fa, chars)
        out = subprocess.run([probe, fa, kind], capture_output=True, text=True)
        assert out.returncode == 0 and "IDENTICAL" in out.stdout


ML_PARAMS = [(name, prec, lvl) for name, case in replay.ML_CASES.items() for prec, lvl in case[5]]


@pytest.mark.parametrize("name,prec,lvl", ML_PARAMS)
def test_oracle_likelihood_matches_reference(olib, name, prec, lvl):
    """pairLogLk (NJ.tcc:1192) and posteriorProfile (NJ.tcc:2137) of the restatement vs the reference's own
    functions: posterior profiles, per-site likelihoods and log-likelihoods bit-identical (same libm)."""
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d_e%d.mldump.bin" % (name, prec, lvl)))
    chars, kind, model = replay.ml_case_chars(name)
    bad, rel = replay.replay_ml(olib, dump, chars, kind, prec, exact_log=True)
    assert bad == [] and rel == 0.0


@pytest.mark.parametrize("name,prec,lvl", ML_PARAMS)
def test_oracle_ml_optimizers_match_reference(olib, name, prec, lvl):
    """SURVEY 8a rows a15-a17 -- MLPairOptimize (NJ.tcc:1790), MLQuartetNNI (:4885) through MLQuartetOptimize (:1650),
    onedimenmin (:7024) and brent (:7098), the per-node body of optimizeAllBranchLengths (:5044) and the whole sweep
    (:5006) -- as lock-step batches (csrc/ml_opt.cpp) over the CPU double of the ABI vs the reference's own functions:
    every optimised length, log-likelihood, NNI choice and the final tree bit-identical."""
    dump = replay.read_refdump(os.path.join(replay.GOLDEN, "%s_f%d_e%d.mldump.bin" % (name, prec, lvl)))
    chars, kind, model = replay.ml_case_chars(name)
    bad, info = replay.replay_ml_opt(olib, dump, chars, kind, prec, exact=True)
    assert bad == [], (bad, info)
    # the batching itself: all quartets advance together, so a call makes far fewer device calls than evaluations
    st = info["quartet.stats"]
    assert st["loglkItems"] > 8 * st["loglkCalls"]


@pytest.mark.parametrize("name,prec", [("nt1000", 32), ("aa300", 32), ("c1", 64)])
def test_speculative_join_is_only_a_hint(olib, name, prec, monkeypatch):
    """The speculative-join path (vft_spec_join_launch / _take / _discard: the guessed NEXT join computed ahead, raw
    distances finished on the host in the reference's arithmetic) must leave the tree byte-identical to the reference's,
    whether a guess is taken or dropped."""
    monkeypatch.setenv("VFT_SPECULATION", "1")
    chars, kind = replay.golden_case(name)
    tables = None
    if kind == "aa":
        z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
        tables = [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]
    tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=olib, tables=tables)
    want = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
    assert tree.newick(["t%d" % i for i in range(chars.shape[0])]) == want
    assert tree.stats["nSpecHit"] > 0.5 * (chars.shape[0] - 3) and tree.stats["nSpecMiss"] > 0


def _bionj_tree(lib, name, prec):
    chars, kind = replay.golden_case(name)
    tables = None
    if kind == "aa":
        z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
        tables = [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]
    tree = api.nj_build(api.encode(chars, kind), 4 if kind == "nt" else 20, prec, lib=lib, tables=tables, bionj=True)
    return tree.newick(["t%d" % i for i in range(chars.shape[0])])


@pytest.mark.parametrize("name", ["nt60", "aa60", "c1", "aa300"])
@pytest.mark.parametrize("prec", [32, 64])
def test_bionj_tree_identical_to_reference(olib, name, prec):
    """-bionj: the BIONJ weight of every join (NJ.tcc:2921-2966, variances read off the out-profile) and the weighted
    averageProfile / diameters that follow -- the tree of the reference binary run with -bionj, byte for byte."""
    want = open(os.path.join(replay.GOLDEN, "%s_f%d.bionj.tree" % (name, prec))).read().strip()
    plain = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
    assert _bionj_tree(olib, name, prec) == want
    assert want != plain          # the option does change the tree: the golden is not vacuous


def test_new_entry_points_validate_their_arguments(olib):
    """Scratch rows, speculative join, SHSupport, the lock-step optimisers: bad ids / sizes are errors, not crashes."""
    import ctypes as C
    chars, kind = replay.golden_case("nt60")
    codes = api.encode(chars, kind)
    N, L = codes.shape
    cfg = api.make_config(N, L, 4, 32, n_scratch=4)
    with api.Context(olib, cfg) as ctx:
        ctx.upload_leaves(codes)
        d = olib.dll
        w = np.ones(L, dtype=np.float32); cd = np.full(L, 127, dtype=np.uint8); v = np.zeros((L, 4), dtype=np.float32)
        assert d.vft_put_profile(ctx.h, 2 * N + 4, api._ptr(w), api._ptr(cd), api._ptr(v)) == -1      # past the scratch rows
        assert d.vft_put_profile(ctx.h, 3, api._ptr(w), api._ptr(cd), api._ptr(v)) == -1              # a leaf row
        assert d.vft_put_profile(ctx.h, 2 * N + 3, api._ptr(w), api._ptr(cd), api._ptr(v)) == 0
        ids = np.array([0, 1], dtype=np.int64)
        assert d.vft_spec_join_launch(ctx.h, N + 5, 0, 1, -1.0, N, api._ptr(ids), 2, None, 0) == -1   # not the next free row
        assert d.vft_spec_join_launch(ctx.h, N, 0, 0, -1.0, N, api._ptr(ids), 2, None, 0) == -1       # a node with itself
        assert d.vft_spec_join_take(ctx.h, 0.0, None, None, None, None, None) == -1                   # nothing pending
        assert d.vft_spec_join_launch(ctx.h, N, 2, 3, -1.0, N, api._ptr(ids), 2, None, 0) == 0
        assert d.vft_spec_join_discard(ctx.h) == 0
        col = np.full((5, L), L, dtype=np.int64)                                                     # column index out of range
        lk = np.zeros((1, 3)); sl = np.ones((1, 3, L)); out = np.zeros(1)
        assert d.vft_sh_support_batch(ctx.h, 1, 5, api._ptr(col), api._ptr(lk), api._ptr(sl), api._ptr(out)) == -1
        crit = np.zeros(3); ch = np.zeros(1, dtype=np.int32)
        bad = np.array([0, 1, 2, 10 * N], dtype=np.int64)
        assert d.vft_choose_nni_batch(ctx.h, 1, api._ptr(bad), 0.0, 1, api._ptr(crit), api._ptr(ch)) == -1


# ---- the device-resident join loop (csrc/nj_loop_logic.h) through its one-thread CPU double (oracle/nj_loop_cpu.cpp) ----
@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["nt60", "aa60", "c1", "aa300", "nt1000"])
def test_device_loop_logic_gives_the_reference_tree(olib, name, prec):
    """The block-parallel re-expression of topHitNJSearch / topHitJoin / updateVisible ... over flat device state, run with
    one thread: the reference's NJ tree byte for byte, and joins, branch lengths and decision counters identical to the
    host-driven loop."""
    chars, kind = replay.golden_case(name)
    want = open(os.path.join(replay.GOLDEN, "%s_f%d.nj.tree" % (name, prec))).read().strip()
    A = 4 if kind == "nt" else 20
    tables = None
    if kind == "aa":
        z = np.load(os.path.join(replay.GOLDEN, "blosum45_f%d.npz" % prec))
        tables = [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]
    codes = api.encode(chars, kind)
    dev = api.nj_build(codes, A, prec, lib=olib, tables=tables, device_loop=1)
    host = api.nj_build(codes, A, prec, lib=olib, tables=tables, device_loop=0)
    assert dev.newick(["t%d" % i for i in range(chars.shape[0])]) == want
    assert np.array_equal(dev.joins, host.joins) and np.array_equal(dev.parent, host.parent)
    assert dev.branchlength.tobytes() == host.branchlength.tobytes()
    for k in ("nRefreshTopHits", "nVisibleUpdate", "nHillBetter"):
        assert dev.stats[k] == host.stats[k], k


def test_device_loop_logic_on_a_larger_tree(olib):
    chars = synth.make_alignment(2500, 120, "nt", seed=9)
    chars = chars[synth.unique_rows(chars)]
    codes = api.encode(chars, "nt")
    dev = api.nj_build(codes, 4, 32, lib=olib, device_loop=1)
    host = api.nj_build(codes, 4, 32, lib=olib, device_loop=0)
    assert np.array_equal(dev.joins, host.joins) and dev.branchlength.tobytes() == host.branchlength.tobytes()
    assert dev.stats["nRefreshTopHits"] == host.stats["nRefreshTopHits"] > 50
