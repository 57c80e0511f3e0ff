"""The C-ABI shared library loads and exports every symbol include/expb200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from exp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "expb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(expb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 35
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build() first"
    L = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    assert set(_declared()) <= set(_lib.SIGNATURES), sorted(set(_declared()) - set(_lib.SIGNATURES))


def test_error_reporting_without_gpu():
    L = _lib.lib()
    assert L.expb200_version() >= 100
    h = C.c_void_p()
    rc = L.expb200_component_create(0, 0, C.byref(h))        # capacity 0 is invalid everywhere
    assert rc == -1 and b"capacity" in L.expb200_last_error()


def test_no_cpu_fallback_when_library_missing(monkeypatch):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libexpb200.so")
    with pytest.raises(_lib.Expb200Error):
        _lib.lib()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "exp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), os.path.join(dirpath, f)
