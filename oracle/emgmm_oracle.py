"""ctypes front-end of oracle/emgmm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  PARITY UNPINNED (see the header of emgmm_oracle.c): the reference holds no golden
vectors for the EM path and cannot be compiled in this image.

Array conventions are NumPy views of the reference's Fortran arrays:
    X   (N, d)    C-contiguous  == Fortran X(d,N)
    Mu  (K, d)                   == Mu(d,K)
    Cov (K, d, d)                == Cov(d,d,K)   (symmetric)
    P   (K, N)                   == ProbaC(N,K)
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libemgmm_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, see oracle/Makefile)."""
    src = os.path.join(_HERE, "emgmm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def _find_lapack() -> str | None:
    try:
        import scipy  # the OpenBLAS bundled with SciPy exports scipy_dsytrd_/dorgtr_/dsteqr_/dsyrk_ (LP64)

        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        cands = [c for c in cands if "64_" not in os.path.basename(c)]
        return os.path.abspath(cands[0]) if cands else None
    except Exception:
        return None


def _d(a):
    return a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self, use_lapack: bool = True):
        build()
        self.lib = lib = C.CDLL(_LIB_PATH)
        lib.orc_load_lapack.argtypes = [C.c_char_p]
        lib.orc_inverse_sym_matrix.argtypes = [C.c_int, _dp, _dp, _dp]
        lib.orc_eval_gaussian.argtypes = [C.c_int, C.c_int64, _dp, _dp, _dp, _dp, C.c_double, _dp]
        lib.orc_max_gaussian.argtypes = [C.c_int, C.c_int64, _dp, _dp, _dp, _dp]
        lib.orc_max_gaussian.restype = C.c_double
        lib.orc_exp_step.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        lib.orc_max_step.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double]
        lib.orc_max_step.restype = None
        lib.orc_em_iterations.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int64, _dp]
        lib.orc_get_class.argtypes = [C.c_int64, C.c_int, _dp, _ip]
        lib.orc_get_class.restype = None
        lib.orc_kmeans_evaluate_class.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, _dp, _ip, _dp]
        lib.orc_kmeans_evaluate_class.restype = None
        lib.orc_kmeans_evaluate_centres.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, _dp, _ip]
        lib.orc_set_threads.argtypes = [C.c_int]
        lib.orc_set_cov_diag.argtypes = [C.c_int]
        self.lapack_path = None
        if use_lapack:
            p = _find_lapack()
            if p and lib.orc_load_lapack(p.encode()) == 0:
                self.lapack_path = p
        lib.orc_use_lapack(1 if self.lapack_path else 0)

    # -- knobs -------------------------------------------------------------------------------
    def use_lapack(self, on: bool) -> None:
        self.lib.orc_use_lapack(1 if on else 0)

    @property
    def has_lapack(self) -> bool:
        return bool(self.lib.orc_has_lapack())

    def set_cov_diag(self, on: bool) -> None:
        """EXTENSION (no reference semantics): diagonal covariances, see emgmm_oracle.c."""
        self.lib.orc_set_cov_diag(1 if on else 0)

    def set_threads(self, n: int) -> None:
        self.lib.orc_set_threads(int(n))

    @property
    def max_threads(self) -> int:
        return int(self.lib.orc_max_threads())

    # -- reference routines --------------------------------------------------------------------
    def inverse_sym_matrix(self, Cm):
        Cm = np.ascontiguousarray(Cm, dtype=np.float64)
        d = Cm.shape[0]
        inv = np.empty((d, d))
        ln = C.c_double()
        info = self.lib.orc_inverse_sym_matrix(d, _d(Cm), _d(inv), C.byref(ln))
        return info, inv, ln.value

    def eval_gaussian(self, X, mu, Cm, lam):
        X = np.ascontiguousarray(X, dtype=np.float64)
        N, d = X.shape
        P = np.empty(N)
        ll = C.c_double()
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        Cm = np.ascontiguousarray(Cm, dtype=np.float64)
        info = self.lib.orc_eval_gaussian(d, N, _d(X), _d(P), _d(mu), _d(Cm), float(lam), C.byref(ll))
        return info, P, ll.value

    def max_gaussian(self, X, proba):
        X = np.ascontiguousarray(X, dtype=np.float64)
        N, d = X.shape
        proba = np.ascontiguousarray(proba, dtype=np.float64)
        mu = np.empty(d)
        Cm = np.empty((d, d))
        s = self.lib.orc_max_gaussian(d, N, _d(X), _d(proba), _d(mu), _d(Cm))
        return s, mu, Cm

    def exp_step(self, X, Mu, Cov, lam):
        X = np.ascontiguousarray(X, dtype=np.float64)
        N, d = X.shape
        K = Mu.shape[0]
        P = np.empty((K, N))
        ll = C.c_double()
        Mu = np.ascontiguousarray(Mu, dtype=np.float64)
        Cov = np.ascontiguousarray(Cov, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        info = self.lib.orc_exp_step(d, N, K, _d(X), _d(P), _d(Mu), _d(Cov), _d(lam), C.byref(ll))
        return info, P, ll.value

    def max_step(self, X, P, min_proba=0.0):
        X = np.ascontiguousarray(X, dtype=np.float64)
        N, d = X.shape
        P = np.ascontiguousarray(P, dtype=np.float64)
        K = P.shape[0]
        Mu = np.empty((K, d))
        Cov = np.empty((K, d, d))
        lam = np.empty(K)
        self.lib.orc_max_step(d, N, K, _d(X), _d(P), _d(Mu), _d(Cov), _d(lam), float(min_proba))
        return Mu, Cov, lam

    def em_iterations(self, X, Mu, Cov, lam, min_proba, niter, return_proba=False):
        """niter reference EM iterations from (Mu, Cov, lam); returns new params + per-iteration sumLL."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        N, d = X.shape
        Mu = np.array(Mu, dtype=np.float64, order="C")
        Cov = np.array(Cov, dtype=np.float64, order="C")
        lam = np.array(lam, dtype=np.float64, order="C")
        K = Mu.shape[0]
        P = np.empty((K, N))
        trace = np.empty(max(int(niter), 1))
        info = self.lib.orc_em_iterations(d, N, K, _d(X), _d(P), _d(Mu), _d(Cov), _d(lam), float(min_proba), int(niter), _d(trace))
        if info != 0:
            raise RuntimeError(f"oracle: LAPACK info={info}")
        out = (Mu, Cov, lam, trace[: int(niter)])
        return out + (P,) if return_proba else out

    def get_class(self, P):
        P = np.ascontiguousarray(P, dtype=np.float64)
        K, N = P.shape
        cl = np.empty(N, dtype=np.int32)
        self.lib.orc_get_class(N, K, _d(P), cl.ctypes.data_as(_ip))
        return cl

    def kmeans_evaluate_class(self, X, Mu):
        X = np.ascontiguousarray(X, dtype=np.float64)
        Mu = np.ascontiguousarray(Mu, dtype=np.float64)
        N, d = X.shape
        cl = np.empty(N, dtype=np.int32)
        md = np.empty(N)
        self.lib.orc_kmeans_evaluate_class(d, N, Mu.shape[0], _d(X), _d(Mu), cl.ctypes.data_as(_ip), _d(md))
        return cl, md

    def kmeans_evaluate_centres(self, X, cl, K):
        X = np.ascontiguousarray(X, dtype=np.float64)
        N, d = X.shape
        Mu = np.empty((K, d))
        cl = np.ascontiguousarray(cl, dtype=np.int32)
        empty = self.lib.orc_kmeans_evaluate_centres(d, N, K, _d(X), _d(Mu), cl.ctypes.data_as(_ip))
        return Mu, empty


_singleton = None


def get() -> Oracle:
    global _singleton
    if _singleton is None:
        _singleton = Oracle()
    return _singleton
