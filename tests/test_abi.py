"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly the symbols that
include/nlc_b200.h declares (no compute calls without a GPU), and the product path refuses to run without it."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "nlc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nlc_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from nlc_b200 import _lib
    assert header_functions() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol(lib_path):
    L = ctypes.CDLL(lib_path)
    missing = [n for n in header_functions() if not hasattr(L, n)]
    assert not missing, "libnlc_b200.so lacks %s" % missing


def test_loader_binds_all_signatures(lib_path):
    from nlc_b200 import _lib
    L = _lib.lib()
    assert L.nlc_abi_version() == 1
    assert L.nlc_last_error() is not None
    # pure host-side size queries are safe without a GPU
    assert L.nlc_groupnorm_ws(2, 64, 128, 32) > 0
    assert L.nlc_attention_ws(1, 2, 256, 1, 256) > 0
    assert L.nlc_attention_ws(1, 2, 16, 1, 512) == 0


def test_no_fallback_when_library_is_missing(monkeypatch, tmp_path):
    from nlc_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "absent.so"))
    with pytest.raises(_lib.NlcError):
        _lib.lib()


def test_create_fails_loudly_without_a_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nlc_b200 import _lib
    with pytest.raises(_lib.NlcError):
        _lib.ctx(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diffusion-nlc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
