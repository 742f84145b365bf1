"""ctypes front end of the CPU oracle (oracle/zelll_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  Nothing under zelll_b200/ imports this package.

The class mirrors the reference's `CellGrid<(usize, [T; N]), N, T>` (src/cellgrid.rs:112-126).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

CMP_NONE, CMP_LT, CMP_LE = 0, 1, 2


class _Info(C.Structure):
    _fields_ = [
        ("inf", C.c_double * 3),
        ("sup", C.c_double * 3),
        ("cutoff", C.c_double),
        ("shape", C.c_int32 * 3),
        ("strides", C.c_int32 * 3),
        ("n", C.c_uint64),
        ("n_cells", C.c_uint64),
        ("buffer_len", C.c_uint64),
    ]


def _native_name() -> str:
    """-march=native code only runs on the CPU it was built for: key the file by the CPU's flags."""
    import hashlib

    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("flags", "model name")):
                    flags += line
                    if line.startswith("flags"):
                        break
    except OSError:
        pass
    return "libzelll_oracle_native_%s.so" % hashlib.sha1(flags.encode()).hexdigest()[:10]


def build(native: bool = False, quiet: bool = True) -> str:
    """Compile the oracle with the committed Makefile; returns the path of the .so."""
    target = _native_name() if native else "libzelll_oracle.so"
    subprocess.run(
        ["make", "-C", _HERE, target],
        check=True,
        stdout=subprocess.DEVNULL if quiet else None,
        stderr=subprocess.STDOUT if quiet else None,
    )
    return os.path.join(_HERE, target)


_LIBS: dict[str, C.CDLL] = {}


def load(native: bool = False) -> C.CDLL:
    name = _native_name() if native else "libzelll_oracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build(native=native)
    lib = C.CDLL(path)
    vp, u64, i32p, dp = C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.zo_grid_create.restype = vp
    lib.zo_grid_create.argtypes = [C.c_int, C.c_int]
    lib.zo_grid_destroy.argtypes = [vp]
    lib.zo_grid_rebuild.restype = C.c_int
    lib.zo_grid_rebuild.argtypes = [vp, vp, u64, dp]
    lib.zo_grid_rebuild_mut.restype = C.c_int
    lib.zo_grid_rebuild_mut.argtypes = [vp, vp, u64, dp, C.POINTER(C.c_int)]
    lib.zo_grid_info.argtypes = [vp, C.POINTER(_Info)]
    lib.zo_grid_keys.argtypes = [vp, vp]
    lib.zo_grid_neighbor_indices.restype = C.c_int
    lib.zo_grid_neighbor_indices.argtypes = [vp, vp]
    lib.zo_grid_cells.argtypes = [vp, vp, vp, vp]
    lib.zo_grid_cell_storage.argtypes = [vp, vp, vp]
    lib.zo_flatten_index.restype = C.c_int32
    lib.zo_flatten_index.argtypes = [vp, i32p]
    lib.zo_try_cell_index.restype = C.c_int
    lib.zo_try_cell_index.argtypes = [vp, dp, i32p]
    lib.zo_flat_cell_index.restype = C.c_int32
    lib.zo_flat_cell_index.argtypes = [vp, dp]
    lib.zo_grid_pair_count.restype = u64
    lib.zo_grid_pair_count.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    lib.zo_grid_pairs.restype = u64
    lib.zo_grid_pairs.argtypes = [vp, C.c_int, C.c_double, vp, u64]
    lib.zo_grid_lj_energy.argtypes = [vp, C.c_int, C.c_double, C.c_int, dp]
    lib.zo_grid_query_neighbors.restype = C.c_int64
    lib.zo_grid_query_neighbors.argtypes = [vp, dp, C.c_int, C.c_double, vp, u64]
    lib.zo_generate_pointcloud.restype = u64
    lib.zo_generate_pointcloud.argtypes = [C.POINTER(u64), C.c_double, dp, vp, u64]
    lib.zo_max_threads.restype = C.c_int
    _LIBS[name] = lib
    return lib


def _dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float32:
        return 0
    if dt == np.float64:
        return 1
    raise TypeError(f"oracle supports float32/float64, got {dt}")


def generate_pointcloud(shape, cutoff: float, origin) -> np.ndarray:
    """`zelll::cellgrid::util::generate_pointcloud` (util.rs:317-340)."""
    lib = load()
    shp = (C.c_uint64 * 3)(*[int(s) for s in shape])
    org = (C.c_double * 3)(*[float(o) for o in origin])
    n = lib.zo_generate_pointcloud(shp, float(cutoff), org, None, 0)
    out = np.empty((n, 3), dtype=np.float64)
    lib.zo_generate_pointcloud(shp, float(cutoff), org, out.ctypes.data, n)
    return out


class OracleCellGrid:
    """CPU restatement of `zelll::CellGrid` over enumerated particles."""

    def __init__(self, points=None, cutoff: float = 1.0, dtype=np.float64, ndim: int = 3, native: bool = False):
        self._lib = load(native=native)
        self.dtype = np.dtype(dtype)
        self.ndim = int(ndim)
        self._h = self._lib.zo_grid_create(_dtype_code(self.dtype), self.ndim)
        if not self._h:
            raise ValueError("bad dtype/ndim")
        if points is not None:
            self.rebuild(points, cutoff)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.zo_grid_destroy(h)

    def _pts(self, points) -> np.ndarray:
        a = np.ascontiguousarray(points, dtype=self.dtype).reshape(-1, self.ndim)
        return a

    @staticmethod
    def _opt(cutoff):
        return None if cutoff is None else C.byref(C.c_double(float(cutoff)))

    # CellGrid::new / rebuild (cellgrid.rs:166-238)
    def rebuild(self, points, cutoff=None) -> None:
        a = self._pts(points)
        rc = self._lib.zo_grid_rebuild(self._h, a.ctypes.data, a.shape[0], self._opt(cutoff))
        if rc != 0:
            raise RuntimeError(f"reference would panic here (rc={rc})")

    # CellGrid::rebuild_mut (cellgrid.rs:264-312); returns FlatIndex::rebuild_mut's flag
    def rebuild_mut(self, points, cutoff=None) -> bool:
        a = self._pts(points)
        changed = C.c_int(0)
        rc = self._lib.zo_grid_rebuild_mut(self._h, a.ctypes.data, a.shape[0], self._opt(cutoff), C.byref(changed))
        if rc != 0:
            raise RuntimeError(f"reference would panic here (rc={rc})")
        return bool(changed.value)

    def info(self) -> dict:
        i = _Info()
        self._lib.zo_grid_info(self._h, C.byref(i))
        nd = self.ndim
        return {
            "inf": [i.inf[d] for d in range(nd)],
            "sup": [i.sup[d] for d in range(nd)],
            "cutoff": i.cutoff,
            "shape": [i.shape[d] for d in range(nd)],
            "strides": [i.strides[d] for d in range(nd)],
            "n": int(i.n),
            "n_cells": int(i.n_cells),
            "buffer_len": int(i.buffer_len),
        }

    def keys(self) -> np.ndarray:
        out = np.empty(self.info()["n"], dtype=np.int32)
        self._lib.zo_grid_keys(self._h, out.ctypes.data)
        return out

    def neighbor_indices(self) -> np.ndarray:
        out = np.empty(3**self.ndim, dtype=np.int32)
        cnt = self._lib.zo_grid_neighbor_indices(self._h, out.ctypes.data)
        return out[:cnt].copy()

    def cells(self):
        """(keys, begin, length) of the non-empty cells in map-iteration order."""
        nc = self.info()["n_cells"]
        keys = np.empty(nc, dtype=np.int32)
        begin = np.empty(nc, dtype=np.uint64)
        length = np.empty(nc, dtype=np.uint64)
        self._lib.zo_grid_cells(self._h, keys.ctypes.data, begin.ctypes.data, length.ctypes.data)
        return keys, begin, length

    def cell_storage(self):
        """(labels, coords) in buffer order (cellgrid.rs:412-414)."""
        m = self.info()["buffer_len"]
        labels = np.empty(m, dtype=np.uint64)
        xyz = np.empty((m, self.ndim), dtype=self.dtype)
        self._lib.zo_grid_cell_storage(self._h, labels.ctypes.data, xyz.ctypes.data)
        return labels, xyz

    def flatten_index(self, idx) -> int:
        a = (C.c_int32 * 3)(*([int(v) for v in idx] + [0] * (3 - len(idx))))
        return int(self._lib.zo_flatten_index(self._h, a))

    def try_cell_index(self, p):
        q = (C.c_double * 3)(*([float(v) for v in p] + [0.0] * (3 - len(p))))
        out = (C.c_int32 * 3)()
        ok = self._lib.zo_try_cell_index(self._h, q, out)
        return [out[d] for d in range(self.ndim)] if ok else None

    def flat_cell_index(self, p) -> int:
        q = (C.c_double * 3)(*([float(v) for v in p] + [0.0] * (3 - len(p))))
        return int(self._lib.zo_flat_cell_index(self._h, q))

    def pair_count(self, cmp: int = CMP_NONE, cutoff: float = 0.0, part: int = 0, full: bool = False,
                   nthreads: int = 1) -> int:
        return int(self._lib.zo_grid_pair_count(self._h, part, int(full), cmp, float(cutoff), nthreads))

    def pairs(self, cmp: int = CMP_NONE, cutoff: float = 0.0) -> np.ndarray:
        """(m, 2) uint32 labels as particle_pairs() yields them (home particle first)."""
        m = int(self._lib.zo_grid_pairs(self._h, cmp, float(cutoff), None, 0))
        out = np.empty((m, 2), dtype=np.uint32)
        self._lib.zo_grid_pairs(self._h, cmp, float(cutoff), out.ctypes.data, m)
        return out

    def pairs_canonical(self, cmp: int = CMP_NONE, cutoff: float = 0.0) -> np.ndarray:
        return canonical_pairs(self.pairs(cmp, cutoff))

    def lj_energy(self, cmp: int = CMP_LT, cutoff: float | None = None, nthreads: int = 1):
        """(sum in T, sum in f64, pairs kept) -- benches/lj.rs:81-92."""
        if cutoff is None:
            cutoff = self.info()["cutoff"]
        out = (C.c_double * 3)()
        self._lib.zo_grid_lj_energy(self._h, cmp, float(cutoff), nthreads, out)
        return out[0], out[1], int(out[2])

    def query_neighbors(self, point, cmp: int = CMP_NONE, cutoff: float = 0.0):
        """Labels query_neighbors() yields (cellgrid.rs:391-401), or None."""
        q = (C.c_double * 3)(*([float(v) for v in point] + [0.0] * (3 - len(point))))
        m = int(self._lib.zo_grid_query_neighbors(self._h, q, cmp, float(cutoff), None, 0))
        if m < 0:
            return None
        out = np.empty(m, dtype=np.uint64)
        self._lib.zo_grid_query_neighbors(self._h, q, cmp, float(cutoff), out.ctypes.data, m)
        return out


def canonical_pairs(pairs: np.ndarray) -> np.ndarray:
    """Sorted (i<j) rows; the form in which pair sets are compared bit-exactly."""
    p = np.asarray(pairs).reshape(-1, 2).astype(np.uint64)
    lo = np.minimum(p[:, 0], p[:, 1])
    hi = np.maximum(p[:, 0], p[:, 1])
    key = (lo << np.uint64(32)) | hi
    key.sort()
    out = np.empty((key.shape[0], 2), dtype=np.uint32)
    out[:, 0] = (key >> np.uint64(32)).astype(np.uint32)
    out[:, 1] = (key & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return out


def max_threads() -> int:
    return int(load().zo_max_threads())
