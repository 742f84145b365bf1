"""ctypes binding of libzelll_b200.so -- the C ABI declared in include/zelll_b200.h.

This is the same binding a Rust `zelll-b200-sys` crate or the PyO3 module would make
(INTEGRATION.md).  There is no fallback: if the CUDA library is missing the import of the
product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZB_LIB") or os.path.join(HERE, "libzelll_b200.so")  # ZB_LIB: experiment builds

F32, F64 = 0, 1
CMP_NONE, CMP_LT, CMP_LE = 0, 1, 2
OK, ERR_BAD_ARG, ERR_CUDA, ERR_CAPACITY, ERR_TOO_MANY, ERR_GRID_TOO_LARGE, ERR_NOT_BUILT, ERR_OUT_OF_WINDOW = range(8)
STATUS_NAMES = ["ZB_OK", "ZB_ERR_BAD_ARG", "ZB_ERR_CUDA", "ZB_ERR_CAPACITY", "ZB_ERR_TOO_MANY",
                "ZB_ERR_GRID_TOO_LARGE", "ZB_ERR_NOT_BUILT", "ZB_ERR_OUT_OF_WINDOW"]
ABI_VERSION = 1
STAGES = ["bbox", "count", "scan", "scatter", "pair_count", "pair_emit", "pair_lj", "other"]


class ZbInfo(C.Structure):
    _fields_ = [
        ("inf", C.c_double * 3),
        ("sup", C.c_double * 3),
        ("cutoff", C.c_double),
        ("shape", C.c_int32 * 3),
        ("strides", C.c_int32 * 3),
        ("n", C.c_uint64),
        ("n_cells", C.c_uint64),
        ("ndim", C.c_int32),
        ("dtype", C.c_int32),
        ("keys_changed", C.c_int32),
        ("reserved", C.c_int32),
    ]


class ZbSlabInfo(C.Structure):
    _fields_ = [
        ("inf", C.c_double * 3),
        ("sup", C.c_double * 3),
        ("shape", C.c_int32 * 3),
        ("reserved", C.c_int32),
        ("z_begin", C.c_int64),
        ("z_end", C.c_int64),
        ("n_local", C.c_uint64),
        ("n_halo", C.c_uint64),
    ]


# every symbol include/zelll_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _i64, _int, _dbl = C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_double
_dp, _u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
SIGNATURES = {
    "zb_abi_version": (_int, []),
    "zb_build_id": (C.c_char_p, []),
    "zb_grid_create": (_int, [_int, _int, _int, C.POINTER(_vp)]),
    "zb_grid_destroy": (None, [_vp]),
    "zb_grid_set_stream": (_int, [_vp, _vp]),
    "zb_grid_track_key_changes": (_int, [_vp, _int]),
    "zb_grid_set_stable": (_int, [_vp, _int]),
    "zb_last_error": (C.c_char_p, [_vp]),
    "zb_grid_rebuild": (_int, [_vp, _vp, _u64, _dp]),
    "zb_grid_prefetch": (_int, [_vp, _vp, _u64]),
    "zb_grid_prefetch_wait": (_int, [_vp]),
    "zb_grid_rebuild_sharded": (_int, [_vp, _vp, _u64, _vp, _dp, _dp, _dp, _i64, _i64]),
    "zb_aabb": (_int, [_vp, _vp, _u64, _dp]),
    "zb_layer_of": (_int, [_vp, _vp, _u64, _dbl, _dbl, _int, _vp]),
    "zb_slab_top_layer": (_int, [_vp, _vp, _u64, _dbl, _dbl, _i64, _i64, C.c_uint32, _vp, _u64, _u64p,
                                 C.POINTER(C.c_int)]),
    "zb_comm_unique_id": (_int, [C.c_char_p, _vp]),
    "zb_comm_init": (_int, [_vp, C.c_char_p, _vp, _int, _int]),
    "zb_grid_rebuild_slab_local": (_int, [_vp, _vp, _u64, _u64, _dp, C.c_uint32, _u64, _vp]),
    "zb_grid_lj_energy_allreduce": (_int, [_vp, _int, _dbl, _dp, _u64p]),
    "zb_grid_info": (_int, [_vp, C.POINTER(ZbInfo)]),
    "zb_grid_keys": (_int, [_vp, _vp]),
    "zb_grid_neighbor_indices": (_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "zb_grid_cells": (_int, [_vp, _vp, _vp, _vp, _u64, _u64p]),
    "zb_grid_cell_storage": (_int, [_vp, _vp, _vp]),
    "zb_grid_pair_count": (_int, [_vp, _int, _dbl, _vp]),
    "zb_grid_pairs": (_int, [_vp, _int, _dbl, _vp, _u64, _u64p]),
    "zb_grid_lj_energy": (_int, [_vp, _int, _dbl, _vp, _vp]),
    "zb_grid_query_neighbors": (_int, [_vp, _vp, _u64, _int, _dbl, _vp, _vp, _vp, _u64, _u64p]),
    "zb_grid_profile": (_int, [_vp, _int]),
    "zb_grid_profile_read": (_int, [_vp, _dp, _u64p]),
    "zb_grid_launch_count": (_u64, [_vp]),
}

_lib = None


class ZelllB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        self.status = status
        name = STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else str(status)
        super().__init__(f"{name}: {message}")


def load() -> C.CDLL:
    """Load libzelll_b200.so and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA engine first (python -m zelll_b200.build). "
            "zelll_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = header / library mismatch, by design
        fn.restype = res
        fn.argtypes = args
    if lib.zb_abi_version() != ABI_VERSION:
        raise ImportError(f"libzelll_b200.so has ABI {lib.zb_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib
