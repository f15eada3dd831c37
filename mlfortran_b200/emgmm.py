"""Host-side mirror of the reference's Fortran object API for the EM-GMM path, over the C-ABI.

Names, argument meaning and error behaviour follow `mlf_algo_emgmm` (src/unsupervised/mlf_emgmm.f90:46-153),
`mlf_step_obj` (src/mlf_step_algo.f90:51-201), the model getters (src/mlf_models.f90:239-331) and
`mlf_hdf5_file` (src/mlf_hdf5.f90:153-225,671-737).  Functions return the reference's integer `info` codes
(0 = OK, negative = error, 1 = mlf_STEP_STOP); nothing is computed in Python.

NumPy views of the Fortran arrays: X (N,d) == X(d,N); Mu (K,d) == Mu(d,K); Cov (K,d,d) == Cov(d,d,K);
ProbaC (K,N) == ProbaC(N,K).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

mlf_OK, mlf_UNINIT, mlf_OTHERERROR = 0, -1, -7
mlf_STEP_OK, mlf_STEP_STOP = 0, 1
mlf_NAME, mlf_DESC, mlf_FIELDS = 1, 2, 3
mlf_INT64, mlf_DOUBLE, mlf_DIRECT = 3, 6, 1
X_ON_DEVICE = 1
COV_DIAG = 2


def mlf_init() -> int:
    return _lib.load().mlf_init()


def mlf_quit() -> int:
    return _lib.load().mlf_quit()


class mlf_hdf5_file:
    """`mlf_hdf5_file` (src/mlf_hdf5.f90:46-82): createFile / openFile / pushState / finalize."""

    def __init__(self):
        self._h = None

    def createFile(self, fname: str, trunk: bool = True) -> int:
        self.finalize()
        self._h = _lib.load().mlf_hdf5_createFile(fname.encode(), 1 if trunk else 0)
        return 0 if self._h else -1

    def openFile(self, fname: str, rw: bool = True) -> int:
        self.finalize()
        self._h = _lib.load().mlf_hdf5_openFile(fname.encode(), 1 if rw else 0)
        return 0 if self._h else -1

    def open_group(self, groupName: str, create: bool = False) -> "mlf_hdf5_file | None":
        """handler => this%open_group(groupName, create) (src/mlf_hdf5.f90:106-131); None on failure."""
        return self._group(groupName, 1 if create else 0)

    def getSubObject(self, obj_name: str) -> "mlf_hdf5_file | None":
        """ptr => this%getSubObject(obj_name): open the group, else create it (src/mlf_hdf5.f90:133-151)."""
        return self._group(obj_name, 2)

    def _group(self, name: str, mode: int):
        h = _lib.load().mlf_b200_h5_open_group(self._h, name.encode(), mode)
        if not h:
            return None
        g = mlf_hdf5_file()
        g._h = h
        return g

    def unsupported(self) -> int:
        return _lib.load().mlf_b200_h5_unsupported(self._h)

    def pushState(self, obj: "mlf_algo_emgmm", override: bool = False) -> int:
        return _lib.load().mlf_pushState(self._h, obj._h, 1 if override else 0)

    def finalize(self) -> None:
        if self._h:
            _lib.load().mlf_dealloc(self._h)
            self._h = None

    # generic dataset access (extension; dims in file/C order)
    def put(self, name: str, arr: np.ndarray, override: bool = False) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.int64:
            dt = mlf_INT64
        elif arr.dtype == np.float64:
            dt = mlf_DOUBLE
        else:
            return -3
        dims = (C.c_int64 * arr.ndim)(*arr.shape)
        return _lib.load().mlf_b200_h5_put(self._h, name.encode(), dt, arr.ndim, dims, arr.ctypes.data, 1 if override else 0)

    def get(self, name: str):
        lib = _lib.load()
        dt, rank = C.c_int(), C.c_int()
        dims = (C.c_int64 * 15)()
        if lib.mlf_b200_h5_get(self._h, name.encode(), C.byref(dt), C.byref(rank), dims, None, 0) != 0:
            return None
        shape = tuple(dims[i] for i in range(rank.value))
        out = np.empty(shape, dtype=np.int64 if dt.value == mlf_INT64 else np.float64)
        if lib.mlf_b200_h5_get(self._h, name.encode(), None, None, None, out.ctypes.data, out.nbytes) != 0:
            return None
        return out

    def names(self):
        lib = _lib.load()
        return [lib.mlf_b200_h5_name(self._h, i).decode() for i in range(max(lib.mlf_b200_h5_count(self._h), 0))]

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass


class mlf_algo_emgmm:
    """`mlf_algo_emgmm` step object.  Public components of the Fortran type are properties here."""

    def __init__(self):
        self._h = None
        self._lib = None
        self._X = None
        self._cb = None

    # -- init / lifetime ------------------------------------------------------------------------------------------
    def init(self, X, nC: int, minProba: float | None = None, data_handler: mlf_hdf5_file | None = None,
             device: int = -1, cov_type: str = "full", devices=None) -> int:
        """info = this%init(X, nC, minProba, data_handler) (src/unsupervised/mlf_emgmm.f90:76-112).

        X: NumPy (N,d) float64 C-contiguous (borrowed, copied once to HBM) or a CUDA torch tensor of that shape
        (borrowed on the device).  With a data_handler the number of components comes from the checkpoint."""
        self.finalize()
        self._lib = lib = _lib.load()
        flags = COV_DIAG if cov_type == "diag" else 0   # extension; "full" is the reference behaviour
        if isinstance(X, np.ndarray):
            if X.dtype != np.float64 or X.ndim != 2 or not X.flags.c_contiguous:
                return mlf_OTHERERROR
            ptr = X.ctypes.data
        else:  # torch tensor on the device
            if X.dim() != 2 or not X.is_contiguous() or str(X.dtype) != "torch.float64" or not X.is_cuda:
                return mlf_OTHERERROR
            ptr = X.data_ptr()
            flags |= X_ON_DEVICE
            if device < 0:
                device = X.device.index
        self._X = X
        n, d = int(X.shape[0]), int(X.shape[1])
        h = data_handler._h if data_handler is not None else None
        if devices is not None:   # single-process multi-GPU object: one shard per listed device ([] = every visible device)
            if flags & X_ON_DEVICE:
                return mlf_OTHERERROR
            dv = (C.c_int * max(len(devices), 1))(*devices)
            self._h = lib.mlf_b200_emgmm_create_multi(ptr, n, d, int(nC), 0.0 if minProba is None else float(minProba), h, dv, len(devices), flags)
        else:
            self._h = lib.mlf_b200_emgmm_create(ptr, n, d, int(nC), 0.0 if minProba is None else float(minProba), h, device, flags)
        if not self._h:
            return mlf_OTHERERROR
        self._bind()
        return mlf_OK

    def _bind(self):
        lib = self._lib
        dt = _lib.MLF_DT()
        rank = C.c_int()
        dims = (C.c_int * 15)()

        def view(idx, ctype, np_dtype):
            p = lib.mlf_getrsc(self._h, idx, C.byref(dt), C.byref(rank), dims, None)
            shape = tuple(reversed([dims[i] for i in range(rank.value)]))  # Fortran dims -> C-order view
            n = int(np.prod(shape))
            if idx == 2:
                n = 3  # rPar: cpuMax, realMax (+ minProba, bound past the exposed extent; see mlf_capi.cu)
                shape = (3,)
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,))
            return arr.reshape(shape).view(np_dtype)

        self._iPar = view(1, C.c_int64, np.int64)
        self._rPar = view(2, C.c_double, np.float64)
        self._iVar = view(3, C.c_int64, np.int64)
        self._rVar = view(4, C.c_double, np.float64)
        self._Mu = view(5, C.c_double, np.float64)
        self._Cov = view(6, C.c_double, np.float64)
        self._lambda = view(7, C.c_double, np.float64)

    def finalize(self) -> None:
        if self._h:
            self._lib.mlf_dealloc(self._h)
        self._h = None
        self._X = None

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass

    # -- public components (host mirrors; direct pointers into the object's arrays) --------------------------------
    Mu = property(lambda self: self._Mu)
    Cov = property(lambda self: self._Cov)
    # `lambda` is a Python keyword
    lambda_ = property(lambda self: self._lambda)
    niter = property(lambda self: int(self._iVar[0]))
    cpuTime = property(lambda self: float(self._rVar[0]))
    realTime = property(lambda self: float(self._rVar[1]))
    sumLL = property(lambda self: float(self._rVar[2]))

    @property
    def minProba(self) -> float:
        return float(self._rPar[2])

    @minProba.setter
    def minProba(self, v: float) -> None:
        self._rPar[2] = v

    @property
    def initialized(self) -> bool:
        return bool(self._lib.mlf_b200_get_initialized(self._h))

    @initialized.setter
    def initialized(self, v: bool) -> None:
        self._lib.mlf_b200_set_initialized(self._h, 1 if v else 0)

    @property
    def ProbaC(self) -> np.ndarray:
        """Responsibilities of the resident points after the last E-step, (K,N) == ProbaC(N,K)."""
        K, n = self._Mu.shape[0], int(self._X.shape[0])
        P = np.empty((K, n))
        info = self._lib.mlf_b200_resident_proba(self._h, P.ctypes.data)
        if info != 0:
            raise RuntimeError(self.last_error())
        return P

    # -- step-object API ----------------------------------------------------------------------------------------------
    def reinit(self) -> int:
        """src/unsupervised/mlf_emgmm.f90:115-122 + src/mlf_step_algo.f90:144-153."""
        self._iVar[0] = 0
        self._iPar[0] = np.iinfo(np.int64).max
        self._rVar[0] = 0.0
        self._rVar[1] = 0.0
        self._rPar[0] = np.finfo(np.float64).max
        self._rPar[1] = np.finfo(np.float64).max
        self._Mu[...] = 0.0
        self._Cov[...] = 0.0
        self.initialized = False
        return 0

    def constraints_step(self, niterMax=None, realMax=None, cpuMax=None) -> None:
        """src/mlf_step_algo.f90:155-162."""
        if niterMax is not None:
            self._iPar[0] = niterMax
        if realMax is not None:
            self._rPar[1] = realMax
        if cpuMax is not None:
            self._rPar[0] = cpuMax

    def step(self, niter: int = 1):
        """info = this%step(dt, niter) (src/mlf_step_algo.f90:187-201).  Returns (info, dt)."""
        dt = C.c_double(0.0)
        n = C.c_int64(int(niter))
        info = self._lib.mlf_step(self._h, C.byref(dt), C.byref(n))
        return info, dt.value

    def stepF(self, niter: int = 1) -> int:
        return self.step(niter)[0]

    # -- model API (getProba / getClass / getNumClasses) ---------------------------------------------------------------
    def getNumClasses(self) -> int:
        return self._lib.mlf_getNumClasses(self._h)

    def getProba(self, X: np.ndarray):
        """info, Proba(K,N') = ExpStep on caller data (src/unsupervised/mlf_emgmm.f90:161-171)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, d = X.shape
        K = self.getNumClasses()
        P = np.empty((K, n))
        info = self._lib.mlf_getProba(self._h, X.ctypes.data, P.ctypes.data, d, n, K)
        return info, P

    def getClass(self, X: np.ndarray):
        """info, cl = MAXLOC(Proba, DIM=2), 1-based (src/mlf_models.f90:239-249)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, d = X.shape
        cl = np.zeros(n, dtype=np.int32)
        info = self._lib.mlf_getClass(self._h, X.ctypes.data, cl.ctypes.data, d, n)
        return info, cl

    # -- B200 extensions -------------------------------------------------------------------------------------------------
    def inject(self, Mu, Cov, lam) -> None:
        """Assign the public Mu/Cov/lambda components and mark the object initialised (what a Fortran caller does by
        assigning em%Mu etc.; bypasses the RNG-driven k-means initialiser)."""
        self._Mu[...] = Mu
        self._Cov[...] = Cov
        self._lambda[...] = lam
        self.initialized = True

    def refresh_x(self, X: np.ndarray) -> int:
        """Tell the object that the caller's X changed (the reference reads `this%X => X` on every step; here X lives in
        HBM).  The upload is streamed underneath the next step's first sweep."""
        if X.dtype != np.float64 or X.ndim != 2 or not X.flags.c_contiguous or X.shape != tuple(self._X.shape):
            return mlf_OTHERERROR
        self._X = X
        return self._lib.mlf_b200_refresh_x(self._h, X.ctypes.data)

    def set_cov_type(self, cov_type: str) -> int:
        """EXTENSION: 'full' (reference behaviour, default) or 'diag' (diagonal covariances, BASELINE config 5)."""
        return self._lib.mlf_b200_set_cov_type(self._h, {"full": 0, "diag": 1}[cov_type])

    def set_seed(self, seed: int) -> None:
        """Seed of the k-means initialiser's random stream (reference: unseeded RANDOM_NUMBER)."""
        self._lib.mlf_b200_set_seed(self._h, int(seed) & (2 ** 64 - 1))

    def labels(self) -> np.ndarray:
        cl = np.zeros(int(self._X.shape[0]), dtype=np.int32)
        info = self._lib.mlf_b200_resident_labels(self._h, cl.ctypes.data)
        if info != 0:
            raise RuntimeError(self.last_error())
        return cl

    def sumLL_trace(self) -> np.ndarray:
        n = self._lib.mlf_b200_sumll_trace(self._h, None, 0)
        out = np.empty(max(int(n), 0))
        if n > 0:
            self._lib.mlf_b200_sumll_trace(self._h, out.ctypes.data_as(_lib.c_double_p), n)
        return out

    def kernel_plan(self) -> str:
        return (self._lib.mlf_b200_kernel_plan(self._h) or b"").decode()

    def last_error(self) -> str:
        return (self._lib.mlf_b200_last_error(self._h) or b"").decode()

    def set_profile(self, on: bool) -> None:
        self._lib.mlf_b200_set_profile(self._h, 1 if on else 0)

    def times(self) -> dict:
        out = (C.c_double * 8)()
        self._lib.mlf_b200_get_times(self._h, out)
        return dict(zip(["refresh_ms", "estep_ms", "mid_ms", "mstep_ms", "finalize_ms", "iters", "sweeps", "side_list_pairs"], list(out)))

    def counters(self) -> dict:
        out = (C.c_int64 * 4)()
        self._lib.mlf_b200_get_counters(self._h, out)
        return dict(zip(["iters", "sweeps", "side_list_pairs", "kept_pairs"], [int(v) for v in out]))

    def qf_stats(self) -> dict:
        """Sweeps that evaluated the log-densities as the expanded quadratic form, and the magnitude bound of the last sweep."""
        n, b = C.c_int64(0), C.c_double(0.0)
        self._lib.mlf_b200_get_qf_stats(self._h, C.byref(n), C.byref(b))
        return {"qf_sweeps": int(n.value), "bound": float(b.value), "scatter_passes": int(self._lib.mlf_b200_scatter_passes(self._h))}

    def launch_count(self) -> int:
        return int(self._lib.mlf_b200_launch_count(self._h))

    def stream(self) -> int:
        return int(self._lib.mlf_b200_stream(self._h) or 0)

    def global_points(self) -> int:
        return int(self._lib.mlf_b200_global_points(self._h))

    def set_allreduce(self, fn, rank: int, world: int) -> None:
        """fn(buf_ptr:int, count:int, stream:int) -> int performs an in-place sum all-reduce over doubles."""

        def _tramp(_ctx, buf, count, stream):
            try:
                return int(fn(int(buf), int(count), int(stream or 0)) or 0)
            except Exception as exc:  # never let an exception cross the C boundary
                import sys
                print(f"mlfortran_b200 all-reduce callback failed: {exc!r}", file=sys.stderr)
                return -1

        self._cb = _lib.ALLREDUCE_FN(_tramp)
        self._lib.mlf_b200_set_allreduce(self._h, self._cb, None, rank, world)


def mlf_evalGaussian(X, mu, C, lam):
    """info, P, sumLL = mlf_evalGaussian(ND, NX, X, P, mu, C, lambda, sumLL) (src/unsupervised/mlf_gaussian.f90:81-88)."""
    import ctypes as _C
    lib = _lib.load()
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d = X.shape
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    Cm = np.ascontiguousarray(C, dtype=np.float64)
    P = np.empty(n)
    ll = _C.c_double()
    info = lib.mlf_evalGaussian(d, n, X.ctypes.data, P.ctypes.data, mu.ctypes.data, Cm.ctypes.data, float(lam), _C.byref(ll))
    return info, P, ll.value


def mlf_maxGaussian(X, Proba):
    """sumPi, mu, C = mlf_maxGaussian(ND, NX, X, Proba, mu, C) (src/unsupervised/mlf_gaussian.f90:53-59)."""
    lib = _lib.load()
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d = X.shape
    Proba = np.ascontiguousarray(Proba, dtype=np.float64)
    mu = np.empty(d)
    Cm = np.empty((d, d))
    s = lib.mlf_maxGaussian(d, n, X.ctypes.data, Proba.ctypes.data, mu.ctypes.data, Cm.ctypes.data)
    return s, mu, Cm
