"""ctypes binding of the C-ABI in include/vft_b200.h (plumbing only; the product is the library).

`load()` opens the CUDA product library (veryfasttree_b200/lib/libvft_b200.so) and fails loudly
when it is missing or when no CUDA device is usable -- there is no CPU fallback on this path.
Tests may pass an explicit path (the test-only CPU double built under oracle/) to exercise the
host logic on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

NOCODE = 127
HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(HERE, "lib", "libvft_b200.so")

CODES_NT = "ACGT"                      # /root/reference/src/Constants.h:51
CODES_AA = "ARNDCQEGHILKMFPSTWYV"      # /root/reference/src/Constants.h:50


class VftConfig(C.Structure):
    _fields_ = [("nSeqs", C.c_int64), ("nPos", C.c_int64), ("nCodes", C.c_int32),
                ("precision", C.c_int32), ("useMatrix", C.c_int32), ("reduction", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32), ("fPostTotalTolerance", C.c_double),
                ("nScratch", C.c_int64)]


class VftCounters(C.Structure):
    _fields_ = [("seqOps", C.c_int64), ("profileOps", C.c_int64), ("outprofileOps", C.c_int64),
                ("profileAvgOps", C.c_int64), ("launches", C.c_int64), ("algoBytes", C.c_int64),
                ("h2dBytes", C.c_int64), ("d2hBytes", C.c_int64), ("msDist", C.c_double), ("msSelect", C.c_double),
                ("msProfile", C.c_double), ("distLaunches", C.c_int64), ("distBytes", C.c_int64),
                ("msKernel", C.c_double * 12), ("nKernel", C.c_int64 * 12), ("bytesKernel", C.c_int64 * 12)]


KERNEL_NAMES = ["k_eval(inline list)", "k_eval(batch)", "k_one_vs_all", "k_out_distance_all", "k_topk_select", "k_merge_prep+finish",
                "k_average", "k_outprofile_update", "k_outprofile_rebuild", "k_pair_loglk", "k_posterior", "k_nj_step"]


class VftMlOptions(C.Structure):
    _fields_ = [("MLMinBranchLength", C.c_double), ("MLFTolBranchLength", C.c_double),
                ("MLMinBranchLengthTolerance", C.c_double), ("closeLogLkLimit", C.c_double),
                ("mlAccuracy", C.c_int32), ("fastNNI", C.c_int32)]


class VftMlStats(C.Structure):
    _fields_ = [("rounds", C.c_int64), ("loglkCalls", C.c_int64), ("loglkItems", C.c_int64),
                ("posteriorCalls", C.c_int64), ("posteriorItems", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class VftNjOptions(C.Structure):
    _fields_ = [("tophitsMult", C.c_double), ("tophitsClose", C.c_double), ("topvisibleMult", C.c_double),
                ("tophitsRefresh", C.c_double), ("staleOutLimit", C.c_double), ("fResetOutProfile", C.c_double),
                ("nResetOutProfile", C.c_int32), ("bionj", C.c_int32), ("prefetch", C.c_int32),
                ("hostThreads", C.c_int32), ("deviceLoop", C.c_int32), ("reserved", C.c_int32)]


class VftNjResult(C.Structure):
    _fields_ = [("parent", C.c_void_p), ("nChild", C.c_void_p), ("child", C.c_void_p),
                ("branchlength", C.c_void_p), ("joins", C.c_void_p), ("leafTopHits", C.c_void_p),
                ("root", C.c_int64), ("maxnode", C.c_int64), ("m", C.c_int64),
                ("nSeeds", C.c_int64), ("nCloseUsed", C.c_int64), ("nRefreshTopHits", C.c_int64),
                ("nVisibleUpdate", C.c_int64), ("nHillBetter", C.c_int64),
                ("nOutPrefetchHit", C.c_int64), ("nOutSingleFetch", C.c_int64),
                ("nPairPrefetchHit", C.c_int64), ("nPairSingleFetch", C.c_int64), ("nDeviceCalls", C.c_int64),
                ("nSpecHit", C.c_int64), ("nSpecMiss", C.c_int64),
                ("secondsLeafTopHits", C.c_double), ("secondsJoins", C.c_double), ("secondsTotal", C.c_double),
                ("deviceMsResident", C.c_double), ("secondsEndToEnd", C.c_double), ("secondsInCalls", C.c_double), ("secondsHost", C.c_double * 8),
                ("counters", VftCounters)]


ABI_SYMBOLS = [
    "vft_ctx_create", "vft_ctx_destroy", "vft_last_error", "vft_backend_name", "vft_upload_tables",
    "vft_upload_leaves", "vft_outprofile_rebuild", "vft_outprofile_update", "vft_profile_average",
    "vft_get_self", "vft_out_distance_batch", "vft_out_distance_all", "vft_dist_pairs",
    "vft_dist_one_vs_all", "vft_get_profile", "vft_get_counters", "vft_nj_default_options", "vft_nj_build",
    "vft_timer_start", "vft_timer_stop", "vft_eval_batch", "vft_profile_average_update",
    "vft_upload_transmat", "vft_sync_rates", "vft_pair_loglk_batch", "vft_posterior_profile",
    "vft_dist_one_vs_all_range", "vft_tophits_merge", "vft_release_cached_memory",
    "vft_posterior_profile_batch", "vft_get_config", "vft_tree_loglk", "vft_set_ml_rates",
    "vft_put_profile", "vft_ml_default_options", "vft_ml_pair_optimize_batch", "vft_ml_quartet_nni_batch",
    "vft_ml_star_optimize_batch", "vft_ml_optimize_branch_lengths", "vft_choose_nni_batch",
    "vft_spec_join_launch", "vft_spec_join_take", "vft_spec_join_discard", "vft_sh_support_batch",
    "vft_ml_split_test_batch", "vft_ml_test_splits",
    "vft_ingest", "vft_profile_average_batch", "vft_recompute_profiles", "vft_dist_unique_id", "vft_dist_init", "vft_dist_init_host", "vft_dist_finalize", "vft_dist_info",
]


class VftError(RuntimeError):
    pass


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class Lib:
    """One loaded implementation of the ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise VftError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                           " (there is no CPU fallback for the CUDA path)")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        d.vft_last_error.restype = C.c_char_p
        d.vft_backend_name.restype = C.c_char_p
        d.vft_ctx_create.argtypes = [C.POINTER(VftConfig), C.POINTER(vp)]
        d.vft_ctx_destroy.argtypes = [vp]
        d.vft_upload_tables.argtypes = [vp, vp, vp, vp, vp]
        d.vft_upload_leaves.argtypes = [vp, vp]
        d.vft_outprofile_rebuild.argtypes = [vp, vp, i64]
        d.vft_outprofile_update.argtypes = [vp, i64, i64, i64, i64]
        d.vft_profile_average.argtypes = [vp, i64, i64, i64, dbl, dbl]
        d.vft_profile_average_update.argtypes = [vp, i64, i64, i64, dbl, dbl, i64]
        if hasattr(d, "vft_upload_transmat"):
            d.vft_upload_transmat.argtypes = [vp, vp, vp, vp, vp, vp]
            d.vft_sync_rates.argtypes = [vp, vp, i64, vp, dbl, dbl, i32]
            d.vft_pair_loglk_batch.argtypes = [vp, vp, vp, vp, i64, vp, vp]
            d.vft_posterior_profile.argtypes = [vp, i64, i64, i64, dbl, dbl]
        d.vft_get_self.argtypes = [vp, i64, C.POINTER(dbl), C.POINTER(dbl)]
        d.vft_out_distance_batch.argtypes = [vp, vp, i64, i64, dbl, vp]
        d.vft_out_distance_all.argtypes = [vp, i64, dbl, vp, i64]
        d.vft_dist_pairs.argtypes = [vp, vp, vp, i64, i32, vp, vp]
        d.vft_eval_batch.argtypes = [vp, vp, i64, i64, dbl, vp, vp, vp, i64, i32, vp, vp]
        d.vft_dist_one_vs_all.argtypes = [vp, i64, i64, i64, vp, vp, vp, vp, C.POINTER(i64)]
        d.vft_dist_one_vs_all_range.argtypes = [vp, i64, i64, i64, i64, i64, vp, vp, vp, vp, C.POINTER(i64)]
        d.vft_get_profile.argtypes = [vp, i64, vp, vp, vp]
        if hasattr(d, "vft_tree_loglk"):
            d.vft_posterior_profile_batch.argtypes = [vp, i64, vp, vp, vp, vp, vp]
            d.vft_tree_loglk.argtypes = [vp, i64, i64, vp, vp, vp, i32, vp, C.POINTER(dbl), vp]
            d.vft_set_ml_rates.argtypes = [vp, i64, i64, vp, vp, vp, i64, dbl, dbl, i32, vp, vp, vp, vp]
        if hasattr(d, "vft_ml_quartet_nni_batch"):
            mo, ms = C.POINTER(VftMlOptions), C.POINTER(VftMlStats)
            d.vft_put_profile.argtypes = [vp, i64, vp, vp, vp]
            d.vft_ml_default_options.argtypes = [i32, mo]
            d.vft_ml_default_options.restype = None
            d.vft_ml_pair_optimize_batch.argtypes = [vp, mo, i64, vp, vp, vp, vp, ms]
            d.vft_ml_quartet_nni_batch.argtypes = [vp, mo, i64, vp, vp, vp, vp, vp, i64, ms]
            d.vft_ml_star_optimize_batch.argtypes = [vp, mo, i64, vp, vp, i64, ms]
            d.vft_ml_optimize_branch_lengths.argtypes = [vp, mo, i64, i64, vp, vp, vp, i32, ms]
            d.vft_choose_nni_batch.argtypes = [vp, i64, vp, dbl, i32, vp, vp]
            d.vft_ml_split_test_batch.argtypes = [vp, mo, i64, vp, vp, vp, vp, vp, vp, i64, ms]
            d.vft_ml_test_splits.argtypes = [vp, mo, i64, i64, vp, vp, vp, i64, vp, vp, C.POINTER(i64), ms]
            d.vft_sh_support_batch.argtypes = [vp, i64, i64, vp, vp, vp, vp]
            d.vft_spec_join_launch.argtypes = [vp, i64, i64, i64, dbl, i64, vp, i64, vp, i64]
            d.vft_spec_join_take.argtypes = [vp, dbl, vp, vp, vp, vp, vp]
            d.vft_spec_join_discard.argtypes = [vp]
        if hasattr(d, "vft_tophits_merge"):
            d.vft_tophits_merge.argtypes = [vp, i64, i64, i64, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
        if hasattr(d, "vft_recompute_profiles"):
            d.vft_profile_average_batch.argtypes = [vp, i64, vp, vp, vp]
            d.vft_recompute_profiles.argtypes = [vp, i64, i64, vp, vp]
        if hasattr(d, "vft_ingest"):
            d.vft_ingest.argtypes = [vp, i64, i64, C.c_char_p, i32, vp, vp, vp, C.POINTER(i64)]
        if hasattr(d, "vft_dist_init"):
            d.vft_dist_unique_id.argtypes = [vp]
            d.vft_dist_init.argtypes = [i32, i32, vp, i32]
            d.vft_dist_init_host.argtypes = [i32, i32, ALLGATHER_FN, vp, i32]
            d.vft_dist_info.argtypes = [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
        d.vft_get_counters.argtypes = [vp, C.POINTER(VftCounters)]
        d.vft_nj_default_options.argtypes = [C.POINTER(VftNjOptions)]
        d.vft_nj_build.argtypes = [C.POINTER(VftConfig), C.POINTER(VftNjOptions), vp, vp, C.POINTER(VftNjResult)]

    @property
    def backend(self) -> str:
        return self.dll.vft_backend_name().decode()

    # ---- one tree sharded over the ranks of a group (include/vft_b200.h, "one tree sharded over the GPUs of a node")
    def dist_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.check(self.dll.vft_dist_unique_id(buf), "vft_dist_unique_id")
        return buf.raw

    def dist_init(self, rank: int, world: int, unique_id: bytes | None, device: int):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self.check(self.dll.vft_dist_init(rank, world, buf, device), "vft_dist_init")

    def dist_init_host(self, rank: int, world: int, allgather, device: int = 0):
        """allgather(send: bytes) -> bytes of world * len(send), rank order (e.g. over torch.distributed / gloo)."""
        def _cb(send, recv, nbytes, _user):
            try:
                out = allgather(C.string_at(send, nbytes))
                assert len(out) == nbytes * world
                C.memmove(recv, out, len(out))
                return 0
            except Exception:           # the C side reports the failure
                import traceback
                traceback.print_exc()
                return 1
        self._allgather_cb = ALLGATHER_FN(_cb)          # keep the trampoline alive
        self.check(self.dll.vft_dist_init_host(rank, world, self._allgather_cb, None, device), "vft_dist_init_host")

    def dist_finalize(self):
        self.check(self.dll.vft_dist_finalize(), "vft_dist_finalize")
        self._allgather_cb = None

    def dist_info(self) -> dict:
        r, w, m = C.c_int32(), C.c_int32(), C.c_int32()
        n, b = C.c_int64(), C.c_int64()
        self.check(self.dll.vft_dist_info(C.byref(r), C.byref(w), C.byref(m), C.byref(n), C.byref(b)), "vft_dist_info")
        return {"rank": r.value, "world": w.value, "mode": ["none", "nccl", "peer", "host"][m.value], "exchanges": n.value, "bytes": b.value}

    def check(self, rc: int, what: str):
        if rc != 0:
            raise VftError(f"{what} failed ({rc}): {self.dll.vft_last_error().decode()}")


_product = None


def load(path: str | None = None) -> Lib:
    """The product library unless a test passes an explicit path."""
    global _product
    if path is not None:
        return Lib(path)
    if _product is None:
        _product = Lib(PRODUCT_LIB)
        if _product.backend != "cuda-sm100a":
            raise VftError("product library reports backend %r" % _product.backend)
    return _product


def encode(chars: np.ndarray, kind: str) -> np.ndarray:
    """ASCII alignment -> codes, as the reference does it: '.'->'-', U->T for nt
    (Alignment.cpp:455-473), letters by position in codesString, everything else NOCODE
    (NeighbourJoining.tcc:415-457)."""
    lut = np.full(256, NOCODE, dtype=np.uint8)
    alphabet = CODES_NT if kind == "nt" else CODES_AA
    for k, ch in enumerate(alphabet):
        lut[ord(ch)] = k
        lut[ord(ch.lower())] = k
    if kind == "nt":
        lut[ord("U")] = lut[ord("T")]
        lut[ord("u")] = lut[ord("T")]
    return np.ascontiguousarray(lut[chars])


def ingest(chars: np.ndarray, kind: str, lib: "Lib | None" = None, device: int = 0):
    """Alignment text [nSeqs][nPos] (uint8 ASCII) -> (codes of the distinct rows, uniqueFirst, alnToUniq) through vft_ingest:
    seqsToProfiles' decoding (NeighbourJoining.tcc:415-457) + Uniquify (Alignment.cpp:494-526)."""
    lib = lib or load()
    text = np.ascontiguousarray(chars, dtype=np.uint8)
    n, L = text.shape
    first = np.empty(n, dtype=np.int64)
    to_uniq = np.empty(n, dtype=np.int64)
    codes = np.empty((n, L), dtype=np.uint8)
    nu = C.c_int64()
    alphabet = (CODES_NT if kind == "nt" else CODES_AA).encode()
    lib.check(lib.dll.vft_ingest(_ptr(text), n, L, alphabet, device, _ptr(first), _ptr(to_uniq), _ptr(codes), C.byref(nu)), "vft_ingest")
    return codes[:nu.value].copy(), first[:nu.value].copy(), to_uniq


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def np_dtype(precision: int):
    return np.float32 if precision == 32 else np.float64


def make_config(n_seqs, n_pos, n_codes, precision, use_matrix=False, reduction=1, device=0, n_scratch=0) -> VftConfig:
    tol = 1.0e-10 if precision == 32 else 1.0e-20     # Constants.h:37-38
    return VftConfig(n_seqs, n_pos, n_codes, precision, int(use_matrix), reduction, device, 0, tol, n_scratch)


class Context:
    """Thin RAII wrapper over vft_ctx for kernel-level calls (tests, smoke, bench)."""

    def __init__(self, lib: Lib, cfg: VftConfig):
        self.lib, self.cfg = lib, cfg
        self.h = C.c_void_p()
        lib.check(lib.dll.vft_ctx_create(C.byref(cfg), C.byref(self.h)), "vft_ctx_create")
        self.dt = np_dtype(cfg.precision)
        self.M = 2 * cfg.nSeqs
        self._maxnode = 0

    def close(self):
        if self.h:
            self.lib.dll.vft_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload_tables(self, distances, eigenval, eigentot, code_freq):
        arrs = [np.ascontiguousarray(x, dtype=self.dt) for x in (distances, eigenval, eigentot, code_freq)]
        self.lib.check(self.lib.dll.vft_upload_tables(self.h, *[_ptr(a) for a in arrs]), "vft_upload_tables")

    def upload_leaves(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        assert codes.shape == (self.cfg.nSeqs, self.cfg.nPos)
        self.lib.check(self.lib.dll.vft_upload_leaves(self.h, _ptr(codes)), "vft_upload_leaves")
        self._maxnode = self.cfg.nSeqs

    def outprofile_rebuild(self, ids=None):
        if ids is None:
            rc = self.lib.dll.vft_outprofile_rebuild(self.h, None, 0)
        else:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            rc = self.lib.dll.vft_outprofile_rebuild(self.h, _ptr(ids), len(ids))
        self.lib.check(rc, "vft_outprofile_rebuild")

    def outprofile_update(self, old1, old2, new, n_active_old):
        self.lib.check(self.lib.dll.vft_outprofile_update(self.h, old1, old2, new, n_active_old),
                       "vft_outprofile_update")

    def profile_average(self, out_id, id1, id2, bionj_weight=-1.0, diameter=0.0):
        self.lib.check(self.lib.dll.vft_profile_average(self.h, out_id, id1, id2, bionj_weight, diameter),
                       "vft_profile_average")
        self._maxnode = max(self._maxnode, out_id + 1)

    def profile_average_update(self, out_id, id1, id2, n_active_old, bionj_weight=-1.0, diameter=0.0):
        self.lib.check(self.lib.dll.vft_profile_average_update(self.h, out_id, id1, id2, bionj_weight, diameter,
                                                               n_active_old), "vft_profile_average_update")
        self._maxnode = max(self._maxnode, out_id + 1)

    def profile_average_batch(self, out_id, id1, id2):
        o, a, b = (np.ascontiguousarray(x, dtype=np.int64) for x in (out_id, id1, id2))
        self.lib.check(self.lib.dll.vft_profile_average_batch(self.h, len(o), _ptr(o), _ptr(a), _ptr(b)), "vft_profile_average_batch")

    def recompute_profiles(self, root, n_child, child):
        nc = np.ascontiguousarray(n_child, dtype=np.int32)
        ch = np.ascontiguousarray(child, dtype=np.int64)
        self.lib.check(self.lib.dll.vft_recompute_profiles(self.h, int(root), len(nc), _ptr(nc), _ptr(ch)), "vft_recompute_profiles")

    def get_self(self, node):
        d, w = C.c_double(), C.c_double()
        self.lib.check(self.lib.dll.vft_get_self(self.h, node, C.byref(d), C.byref(w)), "vft_get_self")
        return d.value, w.value

    def out_distance_batch(self, ids, n_active, totdiam):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.empty(len(ids), dtype=self.dt)
        self.lib.check(self.lib.dll.vft_out_distance_batch(self.h, _ptr(ids), len(ids), n_active, totdiam, _ptr(out)),
                       "vft_out_distance_batch")
        return out

    def out_distance_all(self, n_active, totdiam):
        out = np.zeros(self.M, dtype=self.dt)
        self.lib.check(self.lib.dll.vft_out_distance_all(self.h, n_active, totdiam, _ptr(out), self.M),
                       "vft_out_distance_all")
        return out

    def dist_pairs(self, i, j, flags=0):
        i = np.ascontiguousarray(i, dtype=np.int64)
        j = np.ascontiguousarray(j, dtype=np.int64)
        d = np.empty(len(i), dtype=self.dt)
        w = np.empty(len(i), dtype=self.dt)
        self.lib.check(self.lib.dll.vft_dist_pairs(self.h, _ptr(i), _ptr(j), len(i), flags, _ptr(d), _ptr(w)),
                       "vft_dist_pairs")
        return d, w

    def maxnode(self):
        return self._maxnode

    def dist_one_vs_all(self, query, n_active, k, j_begin=0, j_end=None):
        j = np.empty(k, dtype=np.int64)
        d = np.empty(k, dtype=self.dt)
        w = np.empty(k, dtype=self.dt)
        c = np.empty(k, dtype=self.dt)
        n = C.c_int64()
        if j_end is None:
            j_end = self._maxnode
        self.lib.check(self.lib.dll.vft_dist_one_vs_all_range(self.h, query, n_active, k, j_begin, j_end, _ptr(j), _ptr(d),
                                                              _ptr(w), _ptr(c), C.byref(n)), "vft_dist_one_vs_all_range")
        n = n.value
        return j[:n], d[:n], w[:n], c[:n]

    def tophits_merge(self, newnode, n_active, m, i_nodes, own_offset, own_j, own_dist, all_j, all_dist):
        """vft_tophits_merge: returns (count[nLists], j[nLists, m], dist[nLists, m])."""
        i_nodes = np.ascontiguousarray(i_nodes, dtype=np.int64)
        own_offset = np.ascontiguousarray(own_offset, dtype=np.int64)
        own_j = np.ascontiguousarray(own_j, dtype=np.int64)
        own_dist = np.ascontiguousarray(own_dist, dtype=self.dt)
        all_j = np.ascontiguousarray(all_j, dtype=np.int64)
        all_dist = np.ascontiguousarray(all_dist, dtype=self.dt)
        n_lists = len(i_nodes)
        cnt = np.zeros(n_lists, dtype=np.int64)
        oj = np.full((n_lists, m), -1, dtype=np.int64)
        od = np.zeros((n_lists, m), dtype=self.dt)
        self.lib.check(self.lib.dll.vft_tophits_merge(self.h, newnode, n_active, m, n_lists, _ptr(i_nodes), _ptr(own_offset), _ptr(own_j),
                                                      _ptr(own_dist), len(all_j), _ptr(all_j), _ptr(all_dist),
                                                      _ptr(cnt), _ptr(oj), _ptr(od)), "vft_tophits_merge")
        return cnt, oj, od

    def tree_loglk(self, root, n_child, child, branchlength, recompute=True, leaf_codes=None, site=False):
        """vft_tree_loglk: (loglk, siteLoglk or None)."""
        n_child = np.ascontiguousarray(n_child, dtype=np.int32)
        child = np.ascontiguousarray(child, dtype=np.int64)
        bl = np.ascontiguousarray(branchlength, dtype=self.dt)
        lk = C.c_double()
        sl = np.zeros(self.cfg.nPos, dtype=np.float64) if site else None
        lc = np.ascontiguousarray(leaf_codes, dtype=np.uint8) if leaf_codes is not None else None
        self.lib.check(self.lib.dll.vft_tree_loglk(self.h, int(root), len(n_child), _ptr(n_child), _ptr(child), _ptr(bl),
                                                   1 if recompute else 0, _ptr(lc) if lc is not None else None, C.byref(lk),
                                                   _ptr(sl) if sl is not None else None), "vft_tree_loglk")
        return lk.value, sl

    def set_ml_rates(self, root, n_child, child, branchlength, n_rate_cats, min_rel, min_br, fastexp, leaf_codes=None):
        """vft_set_ml_rates: (rates[nRateCats], ratecat[nPos], siteLoglk[nRateCats, nPos])."""
        n_child = np.ascontiguousarray(n_child, dtype=np.int32)
        child = np.ascontiguousarray(child, dtype=np.int64)
        bl = np.ascontiguousarray(branchlength, dtype=self.dt)
        rates = np.zeros(n_rate_cats, dtype=self.dt)
        ratecat = np.zeros(self.cfg.nPos, dtype=np.int64)
        site = np.zeros((n_rate_cats, self.cfg.nPos), dtype=np.float64)
        lc = np.ascontiguousarray(leaf_codes, dtype=np.uint8) if leaf_codes is not None else None
        self.lib.check(self.lib.dll.vft_set_ml_rates(self.h, int(root), len(n_child), _ptr(n_child), _ptr(child), _ptr(bl), int(n_rate_cats),
                                                     float(min_rel), float(min_br), int(fastexp), _ptr(lc) if lc is not None else None,
                                                     _ptr(rates), _ptr(ratecat), _ptr(site)), "vft_set_ml_rates")
        return rates, ratecat, site

    def get_profile(self, node):
        L, A = self.cfg.nPos, self.cfg.nCodes
        w = np.empty(L, dtype=self.dt)
        cd = np.empty(L, dtype=np.uint8)
        v = np.empty((L, A), dtype=self.dt)
        self.lib.check(self.lib.dll.vft_get_profile(self.h, node, _ptr(w), _ptr(cd), _ptr(v)), "vft_get_profile")
        return w, cd, v

    def put_profile(self, node, weights, codes, vectors):
        w = np.ascontiguousarray(weights, dtype=self.dt)
        cd = np.ascontiguousarray(codes, dtype=np.uint8)
        v = np.ascontiguousarray(vectors, dtype=self.dt)
        self.lib.check(self.lib.dll.vft_put_profile(self.h, int(node), _ptr(w), _ptr(cd), _ptr(v)), "vft_put_profile")

    # ---- branch-length optimisation / ML NNI quartets (lock-step batches, csrc/ml_opt.cpp) ----
    def ml_options(self, **kw) -> VftMlOptions:
        o = VftMlOptions()
        self.lib.dll.vft_ml_default_options(self.cfg.precision, C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def ml_pair_optimize(self, opt, id_a, id_b, length):
        """vft_ml_pair_optimize_batch: (length[n], loglk[n], stats)."""
        a = np.ascontiguousarray(id_a, dtype=np.int64); b = np.ascontiguousarray(id_b, dtype=np.int64)
        ln = np.array(length, dtype=np.float64); lk = np.zeros(len(a), dtype=np.float64)
        st = VftMlStats()
        self.lib.check(self.lib.dll.vft_ml_pair_optimize_batch(self.h, C.byref(opt), len(a), _ptr(a), _ptr(b), _ptr(ln), _ptr(lk),
                                                               C.byref(st)), "vft_ml_pair_optimize_batch")
        return ln, lk, st.as_dict()

    def ml_quartet_nni(self, opt, ids, length, first_scratch_row, criteria=None):
        """vft_ml_quartet_nni_batch: (len[n,5], criteria[n,3], choice[n], star[n], stats)."""
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1, 4)
        n = len(ids)
        ln = np.array(length, dtype=self.dt).reshape(n, 5)
        crit = np.zeros((n, 3), dtype=np.float64) if criteria is None else np.array(criteria, dtype=np.float64).reshape(n, 3)
        choice = np.zeros(n, dtype=np.int32); star = np.zeros(n, dtype=np.int32)
        st = VftMlStats()
        self.lib.check(self.lib.dll.vft_ml_quartet_nni_batch(self.h, C.byref(opt), n, _ptr(ids), _ptr(ln), _ptr(crit), _ptr(choice),
                                                             _ptr(star), int(first_scratch_row), C.byref(st)), "vft_ml_quartet_nni_batch")
        return ln, crit, choice, star, st.as_dict()

    def choose_nni(self, ids, pseudo_weight=0.0, logdist=True):
        """vft_choose_nni_batch: (criteria[n,3], choice[n])."""
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1, 4)
        crit = np.zeros((len(ids), 3), dtype=np.float64); choice = np.zeros(len(ids), dtype=np.int32)
        self.lib.check(self.lib.dll.vft_choose_nni_batch(self.h, len(ids), _ptr(ids), float(pseudo_weight), 1 if logdist else 0,
                                                         _ptr(crit), _ptr(choice)), "vft_choose_nni_batch")
        return crit, choice

    def sh_support(self, col, loglk, site_lk):
        """vft_sh_support_batch: support[n] for col[nBoot, nPos], loglk[n, 3], site_lk[n, 3, nPos]."""
        col = np.ascontiguousarray(col, dtype=np.int64); loglk = np.ascontiguousarray(loglk, dtype=np.float64)
        site_lk = np.ascontiguousarray(site_lk, dtype=np.float64)
        n = loglk.shape[0]
        out = np.zeros(n, dtype=np.float64)
        self.lib.check(self.lib.dll.vft_sh_support_batch(self.h, n, col.shape[0], _ptr(col), _ptr(loglk), _ptr(site_lk), _ptr(out)),
                       "vft_sh_support_batch")
        return out

    def ml_split_test(self, opt, ids, length, first_scratch_row):
        """vft_ml_split_test_batch: (loglk[n,3], siteLk[n,3,nPos], choice[n], bad[n], stats)."""
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1, 4)
        n = len(ids)
        ln = np.ascontiguousarray(length, dtype=self.dt).reshape(n, 5)
        lk = np.zeros((n, 3), dtype=np.float64); site = np.zeros((n, 3, self.cfg.nPos), dtype=np.float64)
        choice = np.zeros(n, dtype=np.int32); bad = np.zeros(n, dtype=np.int32)
        st = VftMlStats()
        self.lib.check(self.lib.dll.vft_ml_split_test_batch(self.h, C.byref(opt), n, _ptr(ids), _ptr(ln), _ptr(lk), _ptr(site), _ptr(choice),
                                                            _ptr(bad), int(first_scratch_row), C.byref(st)), "vft_ml_split_test_batch")
        return lk, site, choice, bad, st.as_dict()

    def ml_test_splits(self, opt, root, n_child, child, branchlength, col):
        """vft_ml_test_splits: (support[maxnode], nBadSplits, stats)."""
        n_child = np.ascontiguousarray(n_child, dtype=np.int32); child = np.ascontiguousarray(child, dtype=np.int64)
        bl = np.ascontiguousarray(branchlength, dtype=self.dt); col = np.ascontiguousarray(col, dtype=np.int64)
        sup = np.zeros(len(n_child), dtype=self.dt)
        nbad = C.c_int64()
        st = VftMlStats()
        self.lib.check(self.lib.dll.vft_ml_test_splits(self.h, C.byref(opt), int(root), len(n_child), _ptr(n_child), _ptr(child), _ptr(bl),
                                                       col.shape[0], _ptr(col), _ptr(sup), C.byref(nbad), C.byref(st)), "vft_ml_test_splits")
        return sup, nbad.value, st.as_dict()

    def ml_star_optimize(self, opt, ids, length, first_scratch_row):
        """vft_ml_star_optimize_batch: (len[n,3], stats)."""
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1, 3)
        n = len(ids)
        ln = np.array(length, dtype=self.dt).reshape(n, 3)
        st = VftMlStats()
        self.lib.check(self.lib.dll.vft_ml_star_optimize_batch(self.h, C.byref(opt), n, _ptr(ids), _ptr(ln), int(first_scratch_row),
                                                               C.byref(st)), "vft_ml_star_optimize_batch")
        return ln, st.as_dict()

    def ml_optimize_branch_lengths(self, opt, root, n_child, child, branchlength, schedule=0):
        """vft_ml_optimize_branch_lengths: (branchlength[maxnode], stats)."""
        n_child = np.ascontiguousarray(n_child, dtype=np.int32)
        child = np.ascontiguousarray(child, dtype=np.int64)
        bl = np.array(branchlength, dtype=self.dt)
        st = VftMlStats()
        self.lib.check(self.lib.dll.vft_ml_optimize_branch_lengths(self.h, C.byref(opt), int(root), len(n_child), _ptr(n_child), _ptr(child),
                                                                   _ptr(bl), int(schedule), C.byref(st)), "vft_ml_optimize_branch_lengths")
        return bl, st.as_dict()

    def counters(self) -> VftCounters:
        c = VftCounters()
        self.lib.check(self.lib.dll.vft_get_counters(self.h, C.byref(c)), "vft_get_counters")
        return c


@dataclass
class NJTree:
    n_seqs: int
    precision: int
    parent: np.ndarray
    n_child: np.ndarray
    child: np.ndarray
    branchlength: np.ndarray
    root: int
    maxnode: int
    m: int
    joins: np.ndarray
    leaf_top_hits: np.ndarray
    stats: dict = field(default_factory=dict)

    def newick(self, names) -> str:
        """printNJ (NeighbourJoining.tcc:2706-2794) for an all-unique alignment."""
        fmt = "%.5f" if self.precision == 32 else "%.9f"
        out = []
        stack = [(self.root, 0)]
        while stack:
            node, end = stack.pop()
            if node < self.n_seqs:
                if self.child[self.parent[node], 0] != node:
                    out.append(",")
                out.append(names[node])
                out.append(":" + fmt % float(self.branchlength[node]))
            elif end:
                if node == self.root:
                    out.append(")")
                else:
                    out.append("):" + fmt % float(self.branchlength[node]))
            else:
                if node != self.root and self.child[self.parent[node], 0] != node:
                    out.append(",")
                out.append("(")
                stack.append((node, 1))
                for k in range(int(self.n_child[node]) - 1, -1, -1):
                    stack.append((int(self.child[node, k]), 0))
        out.append(";")
        return "".join(out)


def nj_build(codes: np.ndarray, n_codes: int, precision: int = 32, lib: Lib | None = None,
             tables=None, device: int = 0, prefetch: bool = True, trace: bool = True,
             reduction: int = 1, profile: bool = False, host_threads: int = 0, bionj: bool = False,
             device_loop: int | None = None) -> NJTree:
    """The metric phase (NJ ctor tail + fastNJ) through vft_nj_build with HOST buffers."""
    lib = lib or load()
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n, L = codes.shape
    dt = np_dtype(precision)
    cfg = make_config(n, L, n_codes, precision, use_matrix=tables is not None, reduction=reduction, device=device)
    cfg.reserved = 1 if profile else 0
    opt = VftNjOptions()
    lib.dll.vft_nj_default_options(C.byref(opt))
    opt.prefetch = int(prefetch)
    opt.hostThreads = host_threads
    opt.bionj = int(bionj)
    if device_loop is not None:
        opt.deviceLoop = int(device_loop)
    M = 2 * n
    parent = np.full(M, -1, dtype=np.int64)
    n_child = np.zeros(M, dtype=np.int32)
    child = np.full((M, 3), -1, dtype=np.int64)
    bl = np.zeros(M, dtype=dt)
    m = int(0.5 + np.sqrt(n))
    joins = np.full((max(n - 3, 0), 2), -1, dtype=np.int64) if trace else None
    lth = np.full((n, max(m, 1)), -1, dtype=np.int64) if trace else None
    res = VftNjResult()
    res.parent, res.nChild, res.child, res.branchlength = _ptr(parent), _ptr(n_child), _ptr(child), _ptr(bl)
    res.joins, res.leafTopHits = _ptr(joins), _ptr(lth)
    tptr = None
    keep = None
    if tables is not None:
        keep = [np.ascontiguousarray(t, dtype=dt) for t in tables]
        arr = (C.c_void_p * 4)(*[t.ctypes.data for t in keep])
        tptr = C.cast(arr, C.c_void_p)
    rc = lib.dll.vft_nj_build(C.byref(cfg), C.byref(opt), _ptr(codes), tptr, C.byref(res))
    lib.check(rc, "vft_nj_build")
    stats = {k: getattr(res, k) for k, _ in VftNjResult._fields_[9:27]}
    stats["secondsHost"] = [float(x) for x in res.secondsHost]
    stats.update({"counters": {k: (list(getattr(res.counters, k)) if k in ("msKernel", "nKernel", "bytesKernel") else getattr(res.counters, k))
                               for k, _ in VftCounters._fields_}})
    return NJTree(n, precision, parent, n_child, child, bl, res.root, res.maxnode, res.m,
                  joins, lth, stats)
