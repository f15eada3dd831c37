"""ctypes binding of libmlfortran_b200.so (the C-ABI declared in include/mlf_b200.h).

The library is the product: there is no Python or CPU fallback.  A missing or unloadable library raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmlfortran_b200.so")
CSRC = os.path.join(_HERE, "csrc")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int64_p = C.POINTER(C.c_int64)


class MLF_DT(C.Structure):
    _fields_ = [("dt", C.c_short), ("access", C.c_short)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

# every symbol include/mlf_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mlf_init": (C.c_int, []),
    "mlf_quit": (C.c_int, []),
    "mlf_dealloc": (C.c_int, [C.c_void_p]),
    "mlf_getnumrsc": (C.c_int, [C.c_void_p]),
    "mlf_getrsc": (C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(MLF_DT), c_int_p, c_int_p, C.c_void_p]),
    "mlf_getinfo": (C.c_char_p, [C.c_void_p, C.c_int, C.c_int]),
    "mlf_updatersc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mlf_pushState": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "mlf_step": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
    "mlf_getClass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mlf_getProba": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mlf_getNumClasses": (C.c_int, [C.c_void_p]),
    "mlf_hdf5_createFile": (C.c_void_p, [C.c_char_p, C.c_int]),
    "mlf_hdf5_openFile": (C.c_void_p, [C.c_char_p, C.c_int]),
    "mlf_evalGaussian": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, c_double_p]),
    "mlf_maxGaussian": (C.c_double, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlf_emgmm_c": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]),
    "mlf_b200_emgmm_create": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_uint]),
    "mlf_b200_emgmm_create_multi": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p, c_int_p, C.c_int, C.c_uint]),
    "mlf_b200_set_allreduce": (C.c_int, [C.c_void_p, ALLREDUCE_FN, C.c_void_p, C.c_int, C.c_int]),
    "mlf_b200_set_nccl_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mlf_b200_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "mlf_b200_nccl_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mlf_b200_set_initialized": (C.c_int, [C.c_void_p, C.c_int]),
    "mlf_b200_get_initialized": (C.c_int, [C.c_void_p]),
    "mlf_b200_set_seed": (C.c_int, [C.c_void_p, C.c_uint64]),
    "mlf_b200_set_cov_type": (C.c_int, [C.c_void_p, C.c_int]),
    "mlf_b200_refresh_x": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mlf_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "mlf_b200_kernel_plan": (C.c_char_p, [C.c_void_p]),
    "mlf_b200_sumll_trace": (C.c_int64, [C.c_void_p, c_double_p, C.c_int64]),
    "mlf_b200_resident_proba": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mlf_b200_resident_labels": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mlf_b200_resident_proba_device": (C.c_void_p, [C.c_void_p, c_int64_p]),
    "mlf_b200_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "mlf_b200_get_times": (C.c_int, [C.c_void_p, c_double_p]),
    "mlf_b200_get_counters": (C.c_int, [C.c_void_p, c_int64_p]),
    "mlf_b200_get_qf_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "mlf_b200_scatter_passes": (C.c_int64, [C.c_void_p]),
    "mlf_b200_host_alloc": (C.c_void_p, [C.c_size_t, C.c_int]),
    "mlf_b200_host_free": (None, [C.c_void_p]),
    "mlf_b200_launch_count": (C.c_int64, [C.c_void_p]),
    "mlf_b200_stream": (C.c_void_p, [C.c_void_p]),
    "mlf_b200_global_points": (C.c_int64, [C.c_void_p]),
    "mlf_b200_version": (C.c_char_p, []),
    "mlf_b200_selftest_exp": (C.c_int, [c_double_p, c_double_p, C.c_int64, C.c_int]),
    "mlf_b200_selftest_exp128": (C.c_int, [c_double_p, c_double_p, C.c_int64, C.c_int]),
    "mlf_b200_synth_gmm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint64,
                                     C.c_int, C.c_uint, C.c_void_p, C.c_void_p]),
    "mlf_b200_engine_create": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint]),
    "mlf_b200_engine_create_multi": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_int_p, C.c_int, C.c_uint]),
    "mlf_b200_engine_destroy": (None, [C.c_void_p]),
    "mlf_b200_engine_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, c_double_p]),
    "mlf_b200_engine_kmeans_init": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlf_b200_engine_expectation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mlf_b200_engine_proba": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mlf_b200_engine_set_allreduce": (C.c_int, [C.c_void_p, ALLREDUCE_FN, C.c_void_p, C.c_int, C.c_int]),
    "mlf_b200_engine_nccl_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mlf_b200_engine_last_error": (C.c_char_p, [C.c_void_p]),
    "mlf_b200_h5_put": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, c_int64_p, C.c_void_p, C.c_int]),
    "mlf_b200_h5_get": (C.c_int, [C.c_void_p, C.c_char_p, c_int_p, c_int_p, c_int64_p, C.c_void_p, C.c_int64]),
    "mlf_b200_h5_count": (C.c_int, [C.c_void_p]),
    "mlf_b200_h5_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "mlf_b200_h5_unsupported": (C.c_int, [C.c_void_p]),
    "mlf_b200_h5_open_group": (C.c_void_p, [C.c_void_p, C.c_char_p, C.c_int]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(mlfortran_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
