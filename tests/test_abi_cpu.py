"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/mlf_b200.h
declares, rejects NULL/foreign handles like the reference (-1 / NULL), and fails loudly without a GPU.
No compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mlfortran_b200 import _lib, mlf_algo_emgmm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    src = open(os.path.join(ROOT, "include", "mlf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(mlf_[A-Za-z0-9_]+)\s*\(", src))
    names -= {"mlf_b200_allreduce_fn"}
    assert len(names) >= 30
    lib = _lib.load()
    for n in sorted(names):
        assert n in _lib.SYMBOLS, f"{n} declared in the header but not bound"
        assert getattr(lib, n) is not None
    assert set(_lib.SYMBOLS) <= names, set(_lib.SYMBOLS) - names


def test_null_and_foreign_handles():
    lib = _lib.load()
    dt = C.c_double(0)
    n = C.c_int64(1)
    assert lib.mlf_dealloc(None) == -1          # src/mlf_intf.f90:261-263
    assert lib.mlf_getnumrsc(None) == -1        # :327-328
    assert lib.mlf_step(None, C.byref(dt), C.byref(n)) == -1  # src/mlf_step_algo.f90:210-213
    assert lib.mlf_getrsc(None, 5, None, None, None, None) is None
    assert lib.mlf_getinfo(None, 1, 1) is None
    assert lib.mlf_updatersc(None, 1, None) == -1
    assert lib.mlf_getNumClasses(None) == -1
    assert lib.mlf_getClass(None, None, None, 2, 0) == -1
    assert lib.mlf_getProba(None, None, None, 2, 0, 1) == -1
    assert lib.mlf_pushState(None, None, 0) == -1
    foreign = (C.c_uint64 * 16)()  # not one of our handles: magic mismatch
    assert lib.mlf_getnumrsc(C.addressof(foreign)) == -1
    assert lib.mlf_step(C.addressof(foreign), C.byref(dt), C.byref(n)) == -1
    assert lib.mlf_dealloc(C.addressof(foreign)) == -1
    assert b"mlfortran_b200" in lib.mlf_b200_version()


def test_fails_loudly_without_gpu():
    lib = _lib.load()
    if _has_cuda():
        assert lib.mlf_init() == 0
        pytest.skip("GPU present: the no-GPU failure path is not reachable")
    assert lib.mlf_init() == -1
    em = mlf_algo_emgmm()
    X = np.zeros((10, 2))
    assert em.init(X, 2, 1.0) != 0  # no CPU fallback: construction fails, nothing is computed
    h = lib.mlf_emgmm_c(X.ctypes.data, 10, 2, 2, 1.0, None)
    assert h is None


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under mlfortran_b200/ may import or link it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mlfortran_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "ORACLEDOC"), f"{f} references the oracle"
