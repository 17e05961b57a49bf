"""Shared helpers of the parity tests: refdump reader, script replay through the C-ABI.

`replay(lib, dump)` drives ANY implementation of include/vft_b200.h (the CUDA product library in
the -m gpu tests, the CPU restatement in the CPU tests) through the script that
oracle/refdump.cpp ran on the reference's own templates, and reports every array that is not
bit-identical.
"""
from __future__ import annotations

import os
import struct
import subprocess

import numpy as np

from veryfasttree_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libvftoracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "VeryFastTree")
REFDUMP_BIN = os.path.join(ROOT, "oracle", "_ref", "refdump")

_DT = {b"f": np.float32, b"d": np.float64, b"q": np.int64, b"B": np.uint8}


def read_refdump(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    o = 0
    while o < len(data):
        (nl,) = struct.unpack_from("<I", data, o); o += 4
        name = data[o:o + nl].decode(); o += nl
        dt = _DT[data[o:o + 1]]; o += 1
        (nd,) = struct.unpack_from("<I", data, o); o += 4
        dims = struct.unpack_from("<%dq" % nd, data, o); o += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(data, dtype=dt, count=n, offset=o).reshape(dims)
        o += n * np.dtype(dt).itemsize
        out[name] = arr
    return out


def ensure_oracle_built():
    if not os.path.exists(ORACLE_LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True,
                       stdout=subprocess.DEVNULL)


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return a.tobytes() == b.tobytes()


def golden_case(name: str):
    """(chars, kind) of a named golden alignment; regenerated from its seed, never stored."""
    n, L, kind, seed = CASES[name]
    chars = synth.make_alignment(n, L, kind, seed)
    chars = chars[synth.unique_rows(chars)]
    return chars, kind


# name -> (n, nPos, kind, seed)
CASES = {
    "nt60": (60, 150, "nt", 5),
    "aa60": (60, 120, "aa", 3),
    "c1": (250, 500, "nt", 1),          # BASELINE.json configs[0]
    "nt1000": (1000, 300, "nt", 7),
    "aa300": (300, 200, "aa", 9),
}


def replay(lib: api.Lib, dump: dict, chars: np.ndarray, kind: str, precision: int, device: int = 0):
    """Returns a list of names of arrays that differ (empty = bit-exact)."""
    bad = []

    def cmp(name, got):
        want = dump[name]
        if not bits_equal(np.asarray(got, dtype=want.dtype).reshape(want.shape), want):
            bad.append(name)

    N, L, A = [int(x) for x in dump["shape"]]
    codes = api.encode(chars, kind)
    assert codes.shape == (N, L)
    use_matrix = kind == "aa"
    cfg = api.make_config(N, L, A, precision, use_matrix=use_matrix, reduction=1, device=device)
    with api.Context(lib, cfg) as ctx:
        if use_matrix:
            ctx.upload_tables(dump["tables.distances"], dump["tables.eigenval"], dump["tables.eigentot"],
                              dump["tables.codeFreq"])
        ctx.upload_leaves(codes)
        # constructor tail: outProfile over all leaves + out-distances
        ctx.outprofile_rebuild()
        w, cd, v = ctx.get_profile(-1)
        cmp("ctor.outprofile.weights", w)
        cmp("ctor.outprofile.vectors", v)
        od = ctx.out_distance_all(N, 0.0)
        cmp("ctor.outDistances", od[:N])
        cmp("ctor.selfweight", [ctx.get_self(i)[1] for i in range(N)])
        # also through the batch entry point
        ids = np.arange(0, N, 3)
        cmp_batch = ctx.out_distance_batch(ids, N, 0.0)
        if not bits_equal(cmp_batch, dump["ctor.outDistances"][ids]):
            bad.append("ctor.outDistances(batch)")
        # leaf x leaf
        d, wt = ctx.dist_pairs(dump["seq.i"], dump["seq.j"])
        cmp("seq.dist", d)
        cmp("seq.weight", wt)
        # the join script
        joins = dump["joins"]
        n_active = N
        totdiam = 0.0
        diam = np.zeros(2 * N, dtype=ctx.dt)
        for k in range(joins.shape[0]):
            a, b = int(joins[k, 0]), int(joins[k, 1])
            nw = N + k
            ctx.profile_average(nw, a, b, -1.0, float(dump["diameters"][k]))
            diam[nw] = ctx.dt(dump["diameters"][k])
            ctx.outprofile_update(a, b, nw, n_active)
            # totdiam += diameter[new] - diameter[a] - diameter[b]  (P arithmetic, NJ.tcc:3036)
            totdiam += float(ctx.dt(ctx.dt(diam[nw] - diam[a]) - diam[b]))
            n_active -= 1
            w, cd, v = ctx.get_profile(nw)
            cmp("join%d.profile.weights" % k, w)
            cmp("join%d.profile.codes" % k, cd)
            cmp("join%d.profile.vectors" % k, v)
            cmp("join%d.self" % k, list(ctx.get_self(nw)))
            if ("join%d.outprofile.weights" % k) in dump:
                w, cd, v = ctx.get_profile(-1)
                cmp("join%d.outprofile.weights" % k, w)
                cmp("join%d.outprofile.vectors" % k, v)
        assert n_active == int(dump["nActive"][0])
        if abs(totdiam - float(dump["totdiam"][0])) != 0.0:
            bad.append("totdiam")
        cmp("out.dist", ctx.out_distance_batch(dump["out.ids"], n_active, totdiam))
        d, wt = ctx.dist_pairs(dump["prof.i"], dump["prof.j"], flags=1)
        cmp("prof.dist", d)
        cmp("prof.weight", wt)
        d, wt = ctx.dist_pairs(dump["join.i"], dump["join.j"], flags=0)
        cmp("join.dist", d)
        cmp("join.weight", wt)
        ctx.outprofile_rebuild()
        w, cd, v = ctx.get_profile(-1)
        cmp("rebuild.outprofile.weights", w)
        cmp("rebuild.outprofile.vectors", v)
    return bad


# name -> (n, nPos, kind, seed, model, [(precision, fastexp level), ...])
ML_CASES = {
    "ml_jc": (30, 90, "nt", 6, "jc", [(32, 0), (64, 0)]),
    "ml_gtr": (30, 90, "nt", 6, "gtr", [(32, 3), (64, 2), (64, 0)]),
    "ml_jtt": (30, 90, "aa", 6, "jtt", [(32, 3), (64, 2), (32, 0)]),
}


def ml_case_chars(name):
    n, L, kind, seed, model, _ = ML_CASES[name]
    chars = synth.make_alignment(n, L, kind, seed)
    return chars[synth.unique_rows(chars)], kind, model


def replay_ml(lib: api.Lib, dump: dict, chars: np.ndarray, kind: str, precision: int, exact_log: bool, device: int = 0):
    """Replays oracle/refdump.cpp's "ml" script (posteriorProfile + pairLogLk of the reference) through
    an implementation of the ABI.  Posterior profiles are compared bit for bit when the model has no libm
    call in them (rational fastexp levels 2/3); log-likelihoods within the north_star tolerance (1e-5
    relative fp32, 1e-12 fp64), or bit for bit when `exact_log` (same libm: the CPU oracle).
    Returns (list of mismatching arrays, max relative log-lk error)."""
    import ctypes as C
    bad = []
    N, L, A = [int(x) for x in dump["shape"]]
    codes = api.encode(chars, kind)
    lvl = int(dump["ml.fastexp"][0])
    has_tm = bool(dump["ml.hasTransmat"][0])
    cfg = api.make_config(N, L, A, precision, use_matrix=False, reduction=1, device=device)
    dt = api.np_dtype(precision)
    with api.Context(lib, cfg) as ctx:
        ctx.upload_leaves(codes)
        d = lib.dll
        if has_tm:
            arrs = [np.ascontiguousarray(dump[k], dtype=dt) for k in ("ml.codeFreq", "ml.eigenval", "ml.eigeninv", "ml.eigeninvT", "ml.statinv")]
            lib.check(d.vft_upload_transmat(ctx.h, *[api._ptr(a) for a in arrs]), "vft_upload_transmat")
        else:
            lib.check(d.vft_upload_transmat(ctx.h, None, None, None, None, None), "vft_upload_transmat")
        rates = np.ascontiguousarray(dump["ml.rates"], dtype=dt)
        ratecat = np.ascontiguousarray(dump["ml.ratecat"], dtype=np.int64)
        lib.check(d.vft_sync_rates(ctx.h, api._ptr(rates), len(rates), api._ptr(ratecat), float(dump["ml.minlen"][0]),
                                   float(dump["ml.minlen"][1]), lvl), "vft_sync_rates")
        libm_free = has_tm and lvl >= 2
        for k, (out, a, b) in enumerate(dump["ml.post.script"]):
            l1, l2 = [float(x) for x in dump["ml.post.lens"][k]]
            lib.check(d.vft_posterior_profile(ctx.h, int(out), int(a), int(b), l1, l2), "vft_posterior_profile")
            w, cd, v = ctx.get_profile(int(out))
            want_w, want_c, want_v = dump["post%d.weights" % k], dump["post%d.codes" % k], dump["post%d.vectors" % k]
            if not np.array_equal(cd, want_c):
                bad.append("post%d.codes" % k)
            if libm_free or exact_log:
                if not bits_equal(w, want_w): bad.append("post%d.weights" % k)
                if not bits_equal(v.reshape(want_v.shape), want_v): bad.append("post%d.vectors" % k)
            else:
                tol = 2e-6 if precision == 32 else 1e-13
                if not np.allclose(w, want_w, rtol=tol, atol=tol): bad.append("post%d.weights~" % k)
                if not np.allclose(v.reshape(want_v.shape), want_v, rtol=tol, atol=tol): bad.append("post%d.vectors~" % k)
        pi = np.ascontiguousarray(dump["ml.lk.i"]); pj = np.ascontiguousarray(dump["ml.lk.j"])
        pl = np.ascontiguousarray(dump["ml.lk.len"], dtype=np.float64)
        ll = np.empty(len(pi), dtype=np.float64)
        site = np.empty((3, L), dtype=np.float64)
        lib.check(d.vft_pair_loglk_batch(ctx.h, api._ptr(pi), api._ptr(pj), api._ptr(pl), len(pi), api._ptr(ll), None), "vft_pair_loglk_batch")
        lib.check(d.vft_pair_loglk_batch(ctx.h, api._ptr(pi[:3].copy()), api._ptr(pj[:3].copy()), api._ptr(pl[:3].copy()), 3,
                                         api._ptr(np.empty(3)), api._ptr(site)), "vft_pair_loglk_batch(site)")
        want = dump["ml.lk.loglk"]
        rel = float(np.max(np.abs(ll - want) / np.maximum(1e-300, np.abs(want))))
        if exact_log and (libm_free or True):
            if not bits_equal(ll, want): bad.append("ml.lk.loglk")
            if not bits_equal(site, dump["ml.lk.site"]): bad.append("ml.lk.site")
        else:
            tol = 1e-5 if precision == 32 else 1e-12
            if rel > tol: bad.append("ml.lk.loglk rel=%g" % rel)
            stol = 2e-6 if precision == 32 else 1e-13
            if libm_free:
                if not bits_equal(site, dump["ml.lk.site"]): bad.append("ml.lk.site")
            elif not np.allclose(site, dump["ml.lk.site"], rtol=stol, atol=0): bad.append("ml.lk.site~")
        # whole-tree sweeps: recomputeMLProfiles + treeLogLk of the reference on its own NJ tree (level-synchronous here)
        if "ml.tree.loglk" in dump:
            root = int(dump["ml.tree.root"][0])
            n_child = dump["ml.tree.nChild"]; child = dump["ml.tree.child"]; bl = dump["ml.tree.branchlength"]
            want_lk = dump["ml.tree.loglk"]
            tol = 1e-5 if precision == 32 else 1e-12
            lk_site, site_lk = ctx.tree_loglk(root, n_child, child, bl, recompute=True, leaf_codes=codes, site=True)
            lk_plain, _ = ctx.tree_loglk(root, n_child, child, bl, recompute=False, leaf_codes=codes, site=False)
            for nm, got, want in (("ml.tree.loglk(site)", lk_site, want_lk[0]), ("ml.tree.loglk", lk_plain, want_lk[1])):
                r = abs(got - want) / max(1e-300, abs(want))
                rel = max(rel, r)
                if exact_log:
                    if got != want: bad.append(nm)
                elif r > tol: bad.append("%s rel=%g" % (nm, r))
            want_site = dump["ml.tree.site"]
            if exact_log:
                if not bits_equal(site_lk, want_site): bad.append("ml.tree.site")
            elif not np.allclose(site_lk, want_site, rtol=10 * tol, atol=10 * tol): bad.append("ml.tree.site~")
            w, cd, v = ctx.get_profile(root - 1)             # the last 2-child node rebuilt by the level-synchronous sweep
            if not np.array_equal(cd, dump["ml.tree.lastnode.codes"]): bad.append("ml.tree.lastnode.codes")
            ptol = 2e-5 if precision == 32 else 1e-11
            if not np.allclose(w, dump["ml.tree.lastnode.weights"], rtol=ptol, atol=ptol): bad.append("ml.tree.lastnode.weights")
            if not np.allclose(v.reshape(dump["ml.tree.lastnode.vectors"].shape), dump["ml.tree.lastnode.vectors"], rtol=ptol, atol=ptol):
                bad.append("ml.tree.lastnode.vectors")
            # setMLRates: 6 whole-tree sweeps + the per-site category choice; then the tree likelihood under the chosen rates
            if "ml.cat.rates" in dump:
                want_r, want_c = dump["ml.cat.rates"], dump["ml.cat.ratecat"]
                got_r, got_c, _ = ctx.set_ml_rates(root, n_child, child, bl, len(want_r), float(dump["ml.minlen"][0]),
                                                   float(dump["ml.minlen"][1]), lvl, leaf_codes=codes)
                lk_cat, _ = ctx.tree_loglk(root, n_child, child, bl, recompute=False, leaf_codes=codes)
                # a near-tie between two categories may flip under the GPU's libm: the categories must agree where the
                # CPU double is used, and almost everywhere otherwise
                if exact_log:
                    if not np.array_equal(got_c, want_c): bad.append("ml.cat.ratecat")
                    if not bits_equal(got_r, want_r): bad.append("ml.cat.rates")
                    if lk_cat != float(dump["ml.cat.loglk"][0]): bad.append("ml.cat.loglk")
                else:
                    if np.mean(got_c != want_c) > 0.02: bad.append("ml.cat.ratecat~")
                    if np.array_equal(got_c, want_c):
                        if not np.allclose(got_r, want_r, rtol=10 * tol, atol=0): bad.append("ml.cat.rates~")
                        r = abs(lk_cat - float(dump["ml.cat.loglk"][0])) / abs(float(dump["ml.cat.loglk"][0]))
                        rel = max(rel, r)
                        if r > tol: bad.append("ml.cat.loglk rel=%g" % r)
    return bad, rel
