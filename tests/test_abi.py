"""The C-ABI library loads (no GPU needed) and exports every symbol include/agcf.h
declares; the ctypes signature table covers exactly that set."""
import os
import re

from arlib_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "agcf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(agcf_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libagcf.so does not export " + s
    assert syms == set(_lib.SIGNATURES), syms ^ set(_lib.SIGNATURES)


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.agcf_abi_version() == 4
    assert lib.agcf_strerror(0) == b"ok"
    assert b"invalid" in lib.agcf_strerror(-1)
    # argument validation happens before any CUDA call, so it is checkable without a GPU
    assert lib.agcf_spmm_csr_f32(None, None, 0, None, None, None, None, None, None, None, None, None, 1.0, None, 0.0, None, None, None, None, 0, None, None, 64, None) == -1
    assert lib.agcf_bpr_ws_bytes(2048) > 0
    assert lib.agcf_score_topk_ws_bytes(10, 100, 48, 5) == -2      # unsupported d
    assert lib.agcf_score_topk_ws_bytes(10, 100, 64, 5) > 0
