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
        if "nni.ids" in dump:        # chooseNNI (NJ.tcc:4836-4852): defaults / pseudo-count prior / no log correction
            for variant, (pw, logd) in enumerate(((0.0, True), (1.0, True), (0.0, False))):
                crit, choice = ctx.choose_nni(dump["nni.ids"], pw, logd)
                cmp("nni%d.criteria" % variant, crit)
                cmp("nni%d.choice" % variant, choice.astype(np.int64))
        if "sh.col" in dump:         # SHSupport (NJ.tcc:1126-1165) against the reference's own resampled columns
            cmp("sh.support", ctx.sh_support(dump["sh.col"], dump["sh.loglk"], dump["sh.siteLk"]))
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


def replay_ml_opt(lib: api.Lib, dump: dict, chars: np.ndarray, kind: str, precision: int, exact: bool, device: int = 0):
    """Replays the a15-a17 section of oracle/refdump.cpp's "ml" script -- the reference's own MLPairOptimize,
    MLQuartetNNI (MLQuartetOptimize / onedimenmin / brent), the per-node body of optimizeAllBranchLengths and the whole
    sweep -- through the lock-step batch entry points.  `exact` (the CPU double of the ABI: same libm): every length,
    criterion, choice and the final tree bit-identical.  Otherwise (the device: exp/log differ in the last bits, which
    can move a Brent iterate) lengths within Brent's own fractional tolerance and log-likelihoods within the
    north_star tolerance.  Returns (list of mismatches, info dict)."""
    bad, info = [], {}
    N, L, A = [int(x) for x in dump["shape"]]
    codes = api.encode(chars, kind)
    lvl = int(dump["ml.fastexp"][0])
    has_tm = bool(dump["ml.hasTransmat"][0])
    q_ids = dump["ml.opt.q.ids"]; nQ = len(q_ids)
    s_ids = dump["ml.opt.s.ids"]; nS = len(s_ids)
    n_scratch = max(nQ + 3 * nQ + 2 * nS + N + 8, N + 3 * 128 + 8)      # vft_ml_test_splits: a row per internal node + 3 per split of a chunk
    cfg = api.make_config(N, L, A, precision, use_matrix=False, reduction=1, device=device, n_scratch=n_scratch)
    dt = api.np_dtype(precision)
    lk_tol = 1e-5 if precision == 32 else 1e-10
    len_rtol = 5e-3                      # Brent's fractional tolerance is 1e-3 (Constants.h:27-28)

    def cmp_len(name, got, want, floor):
        if exact:
            if not bits_equal(np.asarray(got, dtype=want.dtype).reshape(want.shape), want): bad.append(name)
        elif not np.allclose(got.reshape(want.shape), want, rtol=len_rtol, atol=2 * floor): bad.append(name + "~")

    def cmp_lk(name, got, want):
        got = np.asarray(got, dtype=np.float64).reshape(np.shape(want)); want = np.asarray(want, dtype=np.float64)
        if exact:
            if not bits_equal(got, want): bad.append(name)
        else:
            rel = float(np.max(np.abs(got - want) / np.maximum(1e-300, np.abs(want))))
            info[name] = rel
            if rel > lk_tol: bad.append("%s rel=%g" % (name, rel))

    with api.Context(lib, cfg) as ctx:
        ctx.upload_leaves(codes)
        d = lib.dll
        if has_tm:
            arrs = [np.ascontiguousarray(dump[k], dtype=dt) for k in ("ml.codeFreq", "ml.eigenval", "ml.eigeninv", "ml.eigeninvT", "ml.statinv")]
            lib.check(d.vft_upload_transmat(ctx.h, *[api._ptr(a) for a in arrs]), "vft_upload_transmat")
        else:
            lib.check(d.vft_upload_transmat(ctx.h, None, None, None, None, None), "vft_upload_transmat")
        rates = np.ascontiguousarray(dump["ml.cat.rates"], dtype=dt)
        ratecat = np.ascontiguousarray(dump["ml.cat.ratecat"], dtype=np.int64)
        lib.check(d.vft_sync_rates(ctx.h, api._ptr(rates), len(rates), api._ptr(ratecat), float(dump["ml.minlen"][0]),
                                   float(dump["ml.minlen"][1]), lvl), "vft_sync_rates")
        root = int(dump["ml.tree.root"][0])
        n_child = dump["ml.tree.nChild"]; child = dump["ml.tree.child"]; bl = np.ascontiguousarray(dump["ml.tree.branchlength"], dtype=dt)
        lk0, _ = ctx.tree_loglk(root, n_child, child, bl, recompute=True, leaf_codes=codes)      # profiles under the CAT rates
        cmp_lk("ml.cat.loglk", lk0, dump["ml.cat.loglk"][0])
        sc = dump["ml.opt.scalars"]; fl = dump["ml.opt.flags"]
        minlen = float(sc[0])
        opt = ctx.ml_options(MLMinBranchLength=minlen, MLFTolBranchLength=float(sc[1]), MLMinBranchLengthTolerance=float(sc[2]),
                             closeLogLkLimit=float(sc[3]), mlAccuracy=int(fl[0]), fastNNI=int(fl[1]))
        # MLPairOptimize
        ln, lk, st = ctx.ml_pair_optimize(opt, dump["ml.opt.pair.a"], dump["ml.opt.pair.b"], dump["ml.opt.pair.len0"])
        cmp_len("ml.opt.pair.len1", ln, dump["ml.opt.pair.len1"], minlen)
        cmp_lk("ml.opt.pair.loglk", lk, dump["ml.opt.pair.loglk"])
        info["pair.stats"] = st
        # MLQuartetNNI: D (an up-profile, or a root sibling) goes into a scratch row; all quartets in one call
        base = 2 * N
        ids = np.array(q_ids, dtype=np.int64)
        for k in range(nQ):
            ctx.put_profile(base + k, dump["ml.opt.q%d.D.weights" % k], dump["ml.opt.q%d.D.codes" % k], dump["ml.opt.q%d.D.vectors" % k])
            ids[k, 3] = base + k
        # (the reference's serial branch ignores bFast, NJ.tcc:4905-4918: its two dumps per quartet are the same run)
        first = None
        for fi in (0, 1):
            opt.fastNNI = 1
            ln, crit, choice, star, st = ctx.ml_quartet_nni(opt, ids, dump["ml.opt.q.len0"][:, fi, :], base + nQ)
            first = first or (ln, crit, choice, star)
            want_choice = dump["ml.opt.q.choice"][:, fi]; want_crit = dump["ml.opt.q.criteria"][:, fi, :]
            info["quartet.stats"] = st
            info["quartet.star"] = int(star.sum())
            if exact:
                if not np.array_equal(choice, want_choice): bad.append("ml.opt.q.choice[%d]" % fi)
                cmp_len("ml.opt.q.len1[%d]" % fi, ln, dump["ml.opt.q.len1"][:, fi, :], minlen)
                if not bits_equal(crit, np.ascontiguousarray(want_crit)): bad.append("ml.opt.q.criteria[%d]" % fi)
                if not np.array_equal(star != 0, want_crit[:, 1] < -1e19): bad.append("ml.opt.q.star[%d]" % fi)
            else:
                # a choice may only differ where the two best criteria are within the log-likelihood tolerance of each other
                for k in np.nonzero(choice != want_choice)[0]:
                    c = np.sort(want_crit[k])[::-1]
                    if abs(c[0] - c[1]) > 10 * lk_tol * abs(c[0]): bad.append("ml.opt.q.choice[%d][%d]" % (fi, k))
                same = choice == want_choice
                real = want_crit > -1e19
                rel = np.abs(crit - want_crit)[real & same[:, None]] / np.abs(want_crit[real & same[:, None]])
                info["quartet.crit.rel"] = float(rel.max()) if rel.size else 0.0
                if rel.size and rel.max() > 10 * lk_tol: bad.append("ml.opt.q.criteria[%d] rel=%g" % (fi, rel.max()))
                if not np.array_equal(star != 0, want_crit[:, 1] < -1e19): bad.append("ml.opt.q.star[%d]" % fi)
                if not np.allclose(ln[same], dump["ml.opt.q.len1"][:, fi, :][same], rtol=len_rtol, atol=2 * minlen): bad.append("ml.opt.q.len1[%d]~" % fi)
        # without the star test (the reference's sections branch): items the test did not stop are the same computation
        opt.fastNNI = 0
        ln, crit, choice, star, st = ctx.ml_quartet_nni(opt, ids, dump["ml.opt.q.len0"][:, 0, :], base + nQ)
        keep = first[3] == 0
        if star.any() or (crit < -1e19).any(): bad.append("fastNNI=0 ran a star test")
        if not (bits_equal(ln[keep], first[0][keep]) and bits_equal(crit[keep], first[1][keep]) and np.array_equal(choice[keep], first[2][keep])):
            bad.append("fastNNI=0 differs where no star test fired")
        opt.fastNNI = int(fl[1])
        # the per-split body of testSplitsML (MLQuartetLogLk + MLQuartetOptimize x2(+1) with per-site likelihoods), then SHSupport
        if "ml.split.loglk" in dump:
            lk3, site3, ch, bd, st = ctx.ml_split_test(opt, ids, dump["ml.opt.q.len0"][:, 0, :], base + nQ)
            info["split.stats"] = st
            want_lk, want_site = dump["ml.split.loglk"], dump["ml.split.site"]
            if exact:
                if not bits_equal(lk3, np.ascontiguousarray(want_lk)): bad.append("ml.split.loglk")
                if not bits_equal(site3, np.ascontiguousarray(want_site)): bad.append("ml.split.site")
                if not np.array_equal(ch, dump["ml.split.choice"]): bad.append("ml.split.choice")
                if not np.array_equal(bd, dump["ml.split.bad"]): bad.append("ml.split.bad")
            else:
                cmp_lk("ml.split.loglk", lk3, want_lk)
                near = np.abs(np.sort(want_lk, axis=1)[:, -1] - np.sort(want_lk, axis=1)[:, -2]) < 10 * lk_tol * np.abs(want_lk[:, 0]) + 0.11
                if np.any((ch != dump["ml.split.choice"]) & ~near): bad.append("ml.split.choice")
                same = ch == dump["ml.split.choice"]
                if not np.allclose(site3[same], want_site[same], rtol=2e-2, atol=0): bad.append("ml.split.site~")
            sup = ctx.sh_support(dump["ml.split.col"], lk3, site3)
            sup = np.where(bd != 0, 0.0, sup)                                  # a bad split gets support 0 (NJ.tcc:6991)
            info["split.support"] = [round(float(x), 2) for x in sup[:8]]
            if exact:
                if not bits_equal(sup, np.ascontiguousarray(dump["ml.split.support"])): bad.append("ml.split.support")
            elif np.max(np.abs(sup - dump["ml.split.support"])[bd == dump["ml.split.bad"]], initial=0.0) > 0.1: bad.append("ml.split.support~")
        # the per-node body of optimizeAllBranchLengths
        sbase = base + 4 * nQ
        sid = np.array(s_ids, dtype=np.int64)
        for k in range(nS):
            ctx.put_profile(sbase + k, dump["ml.opt.s%d.U.weights" % k], dump["ml.opt.s%d.U.codes" % k], dump["ml.opt.s%d.U.vectors" % k])
            sid[k, 2] = sbase + k
        ln, st = ctx.ml_star_optimize(opt, sid, dump["ml.opt.s.len0"], sbase + nS)
        cmp_len("ml.opt.s.len1", ln, dump["ml.opt.s.len1"], minlen)
        info["star.stats"] = st
        # the whole sweep in the reference's order, then the tree likelihood with the new lengths
        bl1, st = ctx.ml_optimize_branch_lengths(opt, root, n_child, child, bl, schedule=0)
        info["tree.stats.reference"] = st
        cmp_len("ml.opt.tree.branchlength", bl1, np.ascontiguousarray(dump["ml.opt.tree.branchlength"], dtype=dt), minlen)
        lk1, _ = ctx.tree_loglk(root, n_child, child, bl1, recompute=False, leaf_codes=codes)
        cmp_lk("ml.opt.tree.loglk", lk1, dump["ml.opt.tree.loglk"][0])
        # testSplitsML over the optimised tree: SH-like support of every internal split.  CPU double of the ABI (same libm):
        # bit-identical.  On the device the whole-tree path runs too (depth-wise up-profile batches, the lock-step split
        # chunks, vft_sh_support_batch), from the REFERENCE's optimised lengths so that both sides test the same tree; the
        # device's exp/log move likelihoods in the last bits, so: supports within 0.1 wherever the bad-split flag (support 0
        # vs > 0) agrees, and the flag itself may differ only for a few splits (a margin within the likelihood tolerance
        # of treeLogLkDelta, NJ.tcc:6947)
        if "ml.splits.support" in dump:
            want_sup = np.ascontiguousarray(dump["ml.splits.support"], dtype=dt)
            if exact:
                sup, nbad, st = ctx.ml_test_splits(opt, root, n_child, child, bl1, dump["ml.splits.col"])
                info["splits.stats"] = st
                if not bits_equal(sup, want_sup): bad.append("ml.splits.support")
                if nbad != int(dump["ml.splits.nBad"][0]): bad.append("ml.splits.nBad")
            else:
                bl_ref = np.ascontiguousarray(dump["ml.opt.tree.branchlength"], dtype=dt)
                ctx.tree_loglk(root, n_child, child, bl_ref, recompute=True, leaf_codes=codes)
                sup, nbad, st = ctx.ml_test_splits(opt, root, n_child, child, bl_ref, dump["ml.splits.col"])
                info["splits.stats"] = st
                if not np.array_equal(sup < 0, want_sup < 0): bad.append("ml.splits.support: which nodes carry a support")
                has = want_sup >= 0
                flag_same = (sup[has] == 0) == (want_sup[has] == 0)
                info["splits.flagDiff"] = int((~flag_same).sum()); info["splits.n"] = int(has.sum())
                err = np.abs(sup[has] - want_sup[has])[flag_same]
                info["splits.maxErr"] = float(err.max()) if err.size else 0.0
                if (~flag_same).sum() > max(1, 0.02 * has.sum()): bad.append("ml.splits.bad flags differ: %d of %d" % ((~flag_same).sum(), has.sum()))
                if err.size and err.max() > 0.1: bad.append("ml.splits.support~ %g" % err.max())
                if abs(nbad - int(dump["ml.splits.nBad"][0])) > max(1, 0.02 * has.sum()): bad.append("ml.splits.nBad %d vs %d" % (nbad, int(dump["ml.splits.nBad"][0])))
                ctx.tree_loglk(root, n_child, child, bl1, recompute=True, leaf_codes=codes)      # back to the state the next section expects
        # the level-synchronous schedule from the same start: a different (Jacobi-style) visiting order, so not the same
        # lengths -- but it must improve the likelihood about as much as the reference's sweep does
        lk0b, _ = ctx.tree_loglk(root, n_child, child, bl, recompute=True, leaf_codes=codes)
        bl2, st = ctx.ml_optimize_branch_lengths(opt, root, n_child, child, bl, schedule=1)
        info["tree.stats.levels"] = st
        lk2, _ = ctx.tree_loglk(root, n_child, child, bl2, recompute=False, leaf_codes=codes)
        lk2r, _ = ctx.tree_loglk(root, n_child, child, bl2, recompute=True, leaf_codes=codes)
        want1 = float(dump["ml.opt.tree.loglk"][0])
        info["tree.loglk"] = {"start": lk0b, "reference": want1, "levels": lk2}
        if abs(lk2 - lk2r) > 10 * lk_tol * abs(lk2r): bad.append("levels: profiles not current after the sweep")
        if not (lk2 > lk0b): bad.append("levels: likelihood did not improve")
        if (want1 - lk2) > 0.25 * (want1 - lk0b) + 1e-6 * abs(want1): bad.append("levels: gain %.4f vs reference %.4f" % (lk2 - lk0b, want1 - lk0b))
    return bad, info
