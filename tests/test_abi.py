"""CPU tests of the boundary: the product library loads, exports every symbol declared in
include/vft_b200.h, validates arguments, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from veryfasttree_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vft_b200.h")).read()
    return sorted(set(re.findall(r"\b(vft_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_match_binding():
    assert declared_symbols() == sorted(api.ABI_SYMBOLS)


@pytest.fixture(scope="module")
def product():
    if not os.path.exists(api.PRODUCT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    return api.Lib(api.PRODUCT_LIB)


def test_product_library_exports_every_symbol(product):
    for sym in declared_symbols():
        assert hasattr(product.dll, sym), sym
    assert product.backend == "cuda-sm100a"


def test_no_gpu_means_error_not_fallback(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = api.make_config(8, 16, 4, 32)
    h = C.c_void_p()
    rc = product.dll.vft_ctx_create(C.byref(cfg), C.byref(h))
    assert rc == -2                                   # VFT_ENODEVICE
    assert b"no CPU fallback" in product.dll.vft_last_error()


def test_new_entry_points_refuse_to_run_without_a_device(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    text = np.frombuffer(b"ACGTACGA", dtype=np.uint8).reshape(2, 4).copy()
    out = [np.zeros(2, dtype=np.int64), np.zeros(2, dtype=np.int64), np.zeros((2, 4), dtype=np.uint8)]
    nu = C.c_int64()
    rc = product.dll.vft_ingest(text.ctypes.data_as(C.c_void_p), 2, 4, b"ACGT", 0, *[o.ctypes.data_as(C.c_void_p) for o in out], C.byref(nu))
    assert rc == -2 and b"no CPU fallback" in product.dll.vft_last_error()
    assert product.dll.vft_dist_init(0, 1, None, 0) == -2          # VFT_ENODEVICE: no group without a device either
    info = product.dist_info()
    assert info["world"] == 1 and info["mode"] == "none"


def test_argument_validation(product):
    h = C.c_void_p()
    for bad in (api.make_config(0, 16, 4, 32), api.make_config(8, 16, 5, 32), api.make_config(8, 16, 4, 16)):
        assert product.dll.vft_ctx_create(C.byref(bad), C.byref(h)) == -1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(api.VftError):
        api.Lib(str(tmp_path / "nope.so"))
