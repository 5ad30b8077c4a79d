"""ctypes binding of libicp_b200.so (C ABI in include/icp_b200.h).

There is NO CPU fallback: if the CUDA library is missing or cannot be loaded this module
raises, and every product entry point fails loudly with it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ICP_B200_LIB points at another build of the same library (A/B experiments); there is still no fallback
LIB_PATH = os.environ.get("ICP_B200_LIB") or os.path.join(_HERE, "libicp_b200.so")

ICP_OK = 0
ICP_ERR_BAD_ARG = 1
ICP_ERR_TOO_FEW_PAIRS = 2
ICP_ERR_CUDA = 3
ICP_ERR_NCCL = 4
ICP_ERR_OOM = 5
ICP_ERR_STATE = 6
ICP_ERR_SINGULAR = 7

NN_BRUTEFORCE = 0
NN_BVH = 1
NN_SURFACE = 2

SHARD_BASE = 0
SHARD_QUERIES = 1


class IcpParams(C.Structure):
    """Mirror of `icp_params` (include/icp_b200.h)."""

    _fields_ = [
        ("struct_size", C.c_int32),
        ("device", C.c_int32),
        ("thresh", C.c_double),
        ("target_d", C.c_double),
        ("reject_fraction", C.c_double),
        ("stride", C.c_int32),
        ("use_target", C.c_int32),
        ("align_meth", C.c_int32),
        ("nn_mode", C.c_int32),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("nn_variant", C.c_int32),
        ("shard_mode", C.c_int32),
    ]


class IcpIterStats(C.Structure):
    """Mirror of `icp_iter_stats` (include/icp_b200.h)."""

    _fields_ = [
        ("pairs", C.c_int64),
        ("mean_dist", C.c_double),
        ("std_dist", C.c_double),
        ("trans", C.c_double),
        ("angle", C.c_double),
        ("M", C.c_double * 16),
    ]


# name -> (restype, argtypes); every symbol declared in include/icp_b200.h
SIGNATURES = {
    "icp_version": (C.c_int, []),
    "icp_last_error": (C.c_char_p, [C.c_void_p]),
    "icp_device_count": (C.c_int, []),
    "icp_create": (C.c_int, [C.POINTER(IcpParams), C.POINTER(C.c_void_p)]),
    "icp_destroy": (C.c_int, [C.c_void_p]),
    "icp_set_params": (C.c_int, [C.c_void_p, C.POINTER(IcpParams)]),
    "icp_set_base": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "icp_set_base_triangles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "icp_set_moving": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "icp_set_moving_collective": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "icp_set_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "icp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "icp_comm_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "icp_query_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "icp_nn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "icp_make_pairs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.POINTER(C.c_int64), C.c_void_p]),
    "icp_solve_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                  C.c_void_p]),
    "icp_step": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(IcpIterStats)]),
    "icp_run": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                          C.c_void_p]),
    "icp_reset_convergence": (C.c_int, [C.c_void_p]),
    "icp_get_matrix_world": (C.c_int, [C.c_void_p, C.c_void_p]),
    "icp_last_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "icp_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "icp_kernel_time_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double),
                                     C.POINTER(C.c_int64)]),
    "icp_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "icp_synchronize": (C.c_int, [C.c_void_p]),
    "icp_measure_fp32_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "icp_measure_copy_bandwidth": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "icp_flush_l2": (C.c_int, [C.c_void_p]),
    "icp_debug_counters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "icp_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "icp_free_pinned": (C.c_int, [C.c_void_p]),
}

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def lib():
    """Load libicp_b200.so once; raise NativeLibraryMissing (never fall back) if that fails."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C object_alignment_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        try:
            handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        except OSError as e:  # missing libcudart / driver
            raise NativeLibraryMissing("cannot load %s: %s" % (LIB_PATH, e)) from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the ABI lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib
