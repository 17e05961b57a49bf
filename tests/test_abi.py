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


def test_argument_validation(product):
    h = C.c_void_p()
    for bad in (api.make_config(0, 16, 4, 32), api.make_config(8, 16, 5, 32), api.make_config(8, 16, 4, 16)):
        assert product.dll.vft_ctx_create(C.byref(bad), C.byref(h)) == -1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(api.VftError):
        api.Lib(str(tmp_path / "nope.so"))
