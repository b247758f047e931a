"""CPU: the C-ABI library loads without a GPU and exports exactly what include/spades_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from spades_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "spades_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libspades_b200.so not built: run __graft_entry__.build()")
    L = C.CDLL(_lib.LIB_PATH)
    decl = header_symbols()
    assert decl == sorted(_lib.SYMBOLS)
    for s in decl:
        assert hasattr(L, s), s


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.load()
    h = C.c_void_p()
    cfg = _lib.SgpuConfig(0, 0, 0, 0)
    assert L.sgpu_create(C.byref(cfg), C.byref(h)) == 3      # SGPU_ENODEV
    from spades_b200.kmer_index import Context, SpadesGpuError
    with pytest.raises(SpadesGpuError):
        Context(0)
