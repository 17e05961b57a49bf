"""vft_recompute_profiles / vft_profile_average_batch: recomputeProfiles (NeighbourJoining.tcc:3474-3506) as level-synchronous
batches.  averageProfile itself is pinned against the reference by the refdump goldens (test_oracle_golden.py); here
  CPU: the level-synchronous rebuild gives, bit for bit, the profiles of the sequential post-order walk -- with the tables the
       tree was built under (nothing may change) and after a change of basis (other tables: every profile is re-expressed);
  GPU: the device rebuild against the oracle's, every internal profile byte for byte."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import replay  # noqa: E402
from veryfasttree_b200 import api, synth  # noqa: E402


def _tables(prec, swap=False):
    z = np.load(os.path.join(ROOT, "tests", "golden", "blosum45_f%d.npz" % prec))
    t = [z["distances"].copy(), z["eigenval"].copy(), z["eigentot"].copy(), z["codeFreq"].copy()]
    if swap:                     # another basis: any consistent set of tables will do for this test
        t[1] = t[1][::-1].copy(); t[2] = (t[2] * t[2].dtype.type(0.5) + t[2].dtype.type(0.25)); t[3] = t[3][:, ::-1].copy()
    return t


def _build(lib, kind, n, L, prec, seed):
    chars = synth.make_alignment(n, L, kind, seed)
    chars = chars[synth.unique_rows(chars)]
    codes = api.encode(chars, kind)
    A = 4 if kind == "nt" else 20
    tree = api.nj_build(codes, A, prec, lib=lib, tables=_tables(prec) if kind == "aa" else None)
    return codes, A, tree


def _internal_profiles(ctx, tree):
    out = []
    for node in range(tree.n_seqs, tree.maxnode):
        if tree.n_child[node] == 2:
            w, c, v = ctx.get_profile(node)
            out.append((node, w.tobytes(), c.tobytes(), v.tobytes()))
    return out


def _rebuild(lib, codes, A, prec, tree, kind, swap, batch):
    cfg = api.make_config(codes.shape[0], codes.shape[1], A, prec, use_matrix=kind == "aa")
    with api.Context(lib, cfg) as ctx:
        if kind == "aa":
            ctx.upload_tables(*_tables(prec))
        ctx.upload_leaves(codes)
        # the joins of the NJ phase, in order (unweighted averages)
        for k, (i, j) in enumerate(tree.joins):
            ctx.profile_average(tree.n_seqs + k, int(i), int(j))
        before = _internal_profiles(ctx, tree)
        if swap:
            ctx.upload_tables(*_tables(prec, swap=True))
        if batch:
            ctx.recompute_profiles(tree.root, tree.n_child[:tree.maxnode], tree.child[:tree.maxnode])
        else:                    # the reference's serial walk: post-order == join order for the 2-child nodes
            for k, (i, j) in enumerate(tree.joins):
                ctx.profile_average_batch([tree.n_seqs + k], [int(i)], [int(j)])
        return before, _internal_profiles(ctx, tree)


@pytest.mark.parametrize("kind,prec", [("nt", 32), ("aa", 32), ("aa", 64)])
def test_level_synchronous_rebuild_equals_the_serial_walk(kind, prec):
    replay.ensure_oracle_built()
    lib = api.load(replay.ORACLE_LIB)
    codes, A, tree = _build(lib, kind, 150, 80, prec, 4)
    before, same = _rebuild(lib, codes, A, prec, tree, kind, swap=False, batch=True)
    assert before == same                                    # same tables: averageProfile reproduces every join's profile
    if kind == "aa":
        _, serial = _rebuild(lib, codes, A, prec, tree, kind, swap=True, batch=False)
        _, levels = _rebuild(lib, codes, A, prec, tree, kind, swap=True, batch=True)
        assert serial == levels and serial != before


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,L,prec", [("nt", 2000, 200, 32), ("aa", 1200, 700, 32), ("aa", 300, 130, 64)])
def test_device_rebuild_matches_oracle(kind, n, L, prec):
    replay.ensure_oracle_built()
    olib, glib = api.load(replay.ORACLE_LIB), api.load()
    codes, A, tree = _build(olib, kind, n, L, prec, 8)
    for swap in (False, True) if kind == "aa" else (False,):
        _, want = _rebuild(olib, codes, A, prec, tree, kind, swap=swap, batch=True)
        _, got = _rebuild(glib, codes, A, prec, tree, kind, swap=swap, batch=True)
        assert got == want
