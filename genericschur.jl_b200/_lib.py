"""ctypes loader for libgschur_cuda.so (the C ABI declared in include/gschur_cuda.h).

The library is built in-tree (genericschur.jl_b200/libgschur_cuda.so) by csrc/Makefile.  Loading fails loudly:
there is no Python or CPU fallback for any entry point.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSCHUR_CUDA_LIB", os.path.join(_HERE, "libgschur_cuda.so"))   # override: development builds

F64, C64, DD, CDD = 0, 1, 2, 3
FLAG_DEVICE_PTRS = 0x1
FLAG_HESS_INPUT = 0x2
FLAG_CHECK_SUBDIAG = 0x4
ERR_ARG, ERR_CUDA, ERR_SIZE, ERR_SUBDIAG = -1, -2, -3, -4
STATS_PER_MATRIX = 4

EXPORTS = [
    "gschur_cuda_version",
    "gschur_cuda_device_count",
    "gschur_cuda_last_error",
    "gschur_cuda_launch_count",
    "gschur_cuda_max_batched_n",
    "gschur_cuda_batched",
    "gschur_cuda_batched_async",
    "gschur_cuda_hessenberg_batched",
    "gschur_cuda_measure_fp64_peak",
    "gschur_cuda_measure_dmma_peak",
    "gschur_cuda_measure_l2_bandwidth",
    "gschur_cuda_stage_timing",
    "gschur_cuda_stage_timing3",
    "gschur_cuda_release_workspace",
    "gschur_cuda_eigvecs_batched",
    "gschur_cuda_eigvecs_last_error",
    "gschur_cuda_balance_batched",
    "gschur_cuda_balance_apply_batched",
    "gschur_cuda_triangularize_batched",
    "gschur_cuda_balance_last_error",
    "gschur_cuda_hessenberg_large",
    "gschur_cuda_large",
    "gschur_cuda_dgemm",
    "gschur_cuda_large_last_error",
]

_lib = None


def build(jobs=8):
    """Compile the CUDA library for sm_100a with nvcc (works without a GPU)."""
    subprocess.check_call(["make", "-s", "-j", str(jobs), "-C", os.path.join(_HERE, "csrc")])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C genericschur.jl_b200/csrc -j` "
                "(or __graft_entry__.build()). There is no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, i64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32
        L.gschur_cuda_version.restype = ci
        L.gschur_cuda_device_count.restype = ci
        L.gschur_cuda_last_error.restype = ctypes.c_char_p
        L.gschur_cuda_launch_count.restype = ctypes.c_uint64
        L.gschur_cuda_max_batched_n.argtypes = [ci]
        L.gschur_cuda_max_batched_n.restype = ci
        L.gschur_cuda_batched.argtypes = [ci, ci, i64, vp, ci, i64, vp, ci, i64, vp, ci, ci, vp, vp, vp, ci, u32]
        L.gschur_cuda_batched.restype = ci
        L.gschur_cuda_batched_async.argtypes = [ci, ci, i64, vp, ci, i64, vp, ci, i64, vp, ci, ci, vp, vp, vp, u32]
        L.gschur_cuda_batched_async.restype = ci
        L.gschur_cuda_hessenberg_batched.argtypes = [ci, ci, i64, vp, ci, i64, vp, vp, ci, i64, vp, ci, u32]
        L.gschur_cuda_hessenberg_batched.restype = ci
        L.gschur_cuda_measure_fp64_peak.argtypes = [vp, vp]
        L.gschur_cuda_measure_fp64_peak.restype = ci
        L.gschur_cuda_measure_l2_bandwidth.argtypes = [vp, vp]
        L.gschur_cuda_measure_l2_bandwidth.restype = ci
        cd = ctypes.c_double
        L.gschur_cuda_eigvecs_batched.argtypes = [ci, ci, i64, vp, ci, i64, vp, ci, i64, vp, ci, i64, ci, u32]
        L.gschur_cuda_eigvecs_batched.restype = ci
        L.gschur_cuda_eigvecs_last_error.restype = ctypes.c_char_p
        L.gschur_cuda_balance_batched.argtypes = [ci, ci, i64, vp, ci, i64, vp, vp, vp, vp, ci, ci, u32]
        L.gschur_cuda_balance_batched.restype = ci
        L.gschur_cuda_balance_apply_batched.argtypes = [ci, ci, i64, vp, ci, i64, vp, vp, vp, ci, u32]
        L.gschur_cuda_balance_apply_batched.restype = ci
        L.gschur_cuda_triangularize_batched.argtypes = [ci, i64, vp, ci, i64, vp, ci, i64, vp, vp, vp, u32]
        L.gschur_cuda_triangularize_batched.restype = ci
        L.gschur_cuda_balance_last_error.restype = ctypes.c_char_p
        L.gschur_cuda_hessenberg_large.argtypes = [ci, vp, ci, vp, vp, ci, u32]
        L.gschur_cuda_hessenberg_large.restype = ci
        L.gschur_cuda_stage_timing.argtypes = [ci, vp, vp]
        L.gschur_cuda_stage_timing.restype = ci
        L.gschur_cuda_stage_timing3.argtypes = [ci, vp, vp, vp]
        L.gschur_cuda_stage_timing3.restype = ci
        L.gschur_cuda_large.argtypes = [ci, vp, ci, vp, ci, vp, ci, vp, vp, u32]
        L.gschur_cuda_large.restype = ci
        L.gschur_cuda_dgemm.argtypes = [ci, ci, ci, ci, ci, cd, vp, ci, vp, ci, cd, vp, ci]
        L.gschur_cuda_dgemm.restype = ci
        L.gschur_cuda_large_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def last_error():
    return lib().gschur_cuda_last_error().decode()
