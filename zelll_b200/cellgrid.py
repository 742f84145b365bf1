"""Host-side mirror of zelll's `CellGrid` over the C ABI of libzelll_b200.so.

Two faces, both thin:

* the Python binding's entry points (python/src/lib.rs:96-260, stub python/zelll.pyi:4-21):
  `CellGrid(particles=None, /, cutoff=1.0)`, `.rebuild`, `__iter__`, `.aabb()`, `.cutoff()`,
  `.query_neighbors`, `.neighbors`, pickling;
* the Rust API's names (src/cellgrid.rs:166-451) at array granularity: `rebuild_mut`,
  `particle_pairs`, `par_particle_pairs`, `info`, `cell_storage`, plus the fused consumers
  `pair_count` / `lj_energy` of the benches.

All computation happens in the CUDA library; this module only marshals pointers.  Inputs may be
numpy arrays (host) or torch CUDA tensors (device, zero-copy).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, Optional

import numpy as np

from . import _ffi

CMP = {None: _ffi.CMP_NONE, "none": _ffi.CMP_NONE, "lt": _ffi.CMP_LT, "<": _ffi.CMP_LT, "le": _ffi.CMP_LE,
       "<=": _ffi.CMP_LE, 0: 0, 1: 1, 2: 2}


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class GridInfo:
    """`zelll::cellgrid::GridInfo` (src/cellgrid/util.rs:81-297), host-side integer/float helpers."""

    def __init__(self, raw: _ffi.ZbInfo):
        nd = raw.ndim
        self.ndim = nd
        self.dtype = np.dtype(np.float32 if raw.dtype == _ffi.F32 else np.float64)
        T = self.dtype.type
        self._inf = np.array([raw.inf[d] for d in range(nd)], dtype=self.dtype)
        self._sup = np.array([raw.sup[d] for d in range(nd)], dtype=self.dtype)
        self._cutoff = T(raw.cutoff)
        self._shape = np.array([raw.shape[d] for d in range(nd)], dtype=np.int32)
        self._strides = np.array([raw.strides[d] for d in range(nd)], dtype=np.int32)
        self.n = int(raw.n)
        self.n_cells = int(raw.n_cells)
        self.keys_changed = None if raw.keys_changed < 0 else bool(raw.keys_changed)

    def origin(self):  # util.rs:139-141
        return self._inf.copy()

    def shape(self):  # util.rs:144-146
        return self._shape.copy()

    def strides(self):  # util.rs:149-151
        return self._strides.copy()

    def bounding_box(self):  # util.rs:154-156
        return self._inf.copy(), self._sup.copy()

    def cutoff(self):  # util.rs:179-181
        return float(self._cutoff)

    def flatten_index(self, idx) -> int:  # util.rs:171-176 (wrapping i32 dot product)
        acc = 0
        for d in range(self.ndim):
            acc += int(idx[d]) * int(self._strides[d])
        return ((acc + 2**31) % 2**32) - 2**31

    def _raw_index(self, coords):
        p = np.asarray(coords, dtype=self.dtype).reshape(self.ndim)
        with np.errstate(all="ignore"):
            q = np.floor((p - self._inf) / self._cutoff)
        out = []
        for v in q:  # Rust `as i32`: saturating, NaN -> 0
            if np.isnan(v):
                out.append(0)
            else:
                out.append(int(min(max(float(v), -2147483648.0), 2147483647.0)))
        return out

    def try_cell_index(self, coords):  # util.rs:245-256
        idx = self._raw_index(coords)
        if all(-1 <= idx[d] <= int(self._shape[d]) for d in range(self.ndim)):
            return idx
        return None

    def cell_index(self, coords):  # util.rs:229-232 (the reference panics)
        idx = self.try_cell_index(coords)
        if idx is None:
            raise IndexError("cell index is out of bounds")
        return idx

    def flat_cell_index(self, coords) -> int:  # util.rs:291-297 (no bounds check)
        return self.flatten_index(self._raw_index(coords))


class CellGridIter:
    """`zelll.CellGridIter` (python/src/lib.rs:273-345): yields ((i, [x,y,z]), (j, [x,y,z]))."""

    def __init__(self, pairs: np.ndarray, coords_by_label, owner=None):
        self._pairs = pairs
        self._coords = coords_by_label
        self._k = 0
        self._owner = owner  # a live iterator borrows the grid (python/src/lib.rs:138-143)
        if owner is not None:
            owner._live_iters += 1

    def __del__(self):
        owner, self._owner = getattr(self, "_owner", None), None
        if owner is not None:
            owner._live_iters -= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self._k >= self._pairs.shape[0]:
            raise StopIteration
        i, j = self._pairs[self._k]
        self._k += 1
        i, j = int(i), int(j)
        return (i, self._coords(i)), (j, self._coords(j))

    def __len__(self):
        return self._pairs.shape[0] - self._k


class CellQueryIter:
    """`zelll.CellQueryIter` (python/src/lib.rs:357-394): yields (j, [x,y,z])."""

    def __init__(self, labels: np.ndarray, coords_by_label):
        self._labels = labels
        self._coords = coords_by_label
        self._k = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self._k >= self._labels.shape[0]:
            raise StopIteration
        j = int(self._labels[self._k])
        self._k += 1
        return j, self._coords(j)


class GridCell:
    """`zelll::cellgrid::GridCell` (src/cellgrid/iters.rs:121-241): a view of one (possibly empty) cell
    over the cell-sorted storage fetched once by `CellGrid.iter_cells()` / `.query()`."""

    def __init__(self, view: "_CellView", index: int):
        self._v = view
        self.index = int(index)  # flat cell key (iters.rs:143-146)

    def iter(self):
        """Particles of this cell as (label, [coords]) (iters.rs:154-168); empty for an absent cell."""
        k = self._v.slot.get(self.index)
        if k is None:
            return iter(())
        b, c = int(self._v.begin[k]), int(self._v.count[k])
        return ((int(self._v.labels[p]), [float(x) for x in self._v.xyz[p]]) for p in range(b, b + c))

    __iter__ = iter

    def __len__(self):
        k = self._v.slot.get(self.index)
        return 0 if k is None else int(self._v.count[k])

    def neighbors(self, full: bool = False):
        """Non-empty neighbour cells (iters.rs:197-214): Half = the first half of
        `FlatIndex::neighbor_indices` (iters.rs:58-63), Full = all of them (iters.rs:44-56)."""
        rel = self._v.neighbor_indices
        if not full:
            rel = rel[: len(rel) // 2]
        out = []
        for r in rel:
            key = _wrap_i32(self.index + int(r))
            if key in self._v.slot:
                out.append(GridCell(self._v, key))
        return out

    def intra_cell_pairs(self, full: bool = False):
        """iters.rs:29-36 (Half: j after i) / :48-55 (Full: all ordered pairs i != j)."""
        ps = list(self.iter())
        if full:
            return [(a, b) for i, a in enumerate(ps) for j, b in enumerate(ps) if i != j]
        return [(a, b) for i, a in enumerate(ps) for b in ps[i + 1:]]

    def inter_cell_pairs(self, full: bool = False):
        """iters.rs:228-231: this cell x the particles of its neighbour cells."""
        others = [q for cell in self.neighbors(full) for q in cell.iter()]
        return [(p, q) for p in self.iter() for q in others]

    def particle_pairs(self):
        """iters.rs:238-241: intra::<Half> ++ inter::<Half>."""
        return self.intra_cell_pairs(False) + self.inter_cell_pairs(False)


def _wrap_i32(v: int) -> int:
    return ((v + 2**31) % 2**32) - 2**31


class _CellView:
    """Host snapshot of the grid's CSR arrays shared by the GridCell views of one rebuild."""

    def __init__(self, grid: "CellGrid"):
        self.keys, self.begin, self.count = grid.cells()
        self.labels, self.xyz = grid.cell_storage()
        self.neighbor_indices = [int(v) for v in grid.neighbor_indices()]
        self.slot = {int(k): i for i, k in enumerate(self.keys)}


class CellGrid:
    """`zelll.CellGrid` on a B200.

    `CellGrid(particles=None, /, cutoff=1.0)` as in python/src/lib.rs:111-131; the keyword-only
    extras choose what the Rust generics choose at compile time (`T`, `N`) and the CUDA device.
    """

    def __init__(self, particles=None, /, cutoff: float = 1.0, *, dtype=np.float64, ndim: int = 3, device: int = 0):
        self._lib = _ffi.load()
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("dtype must be float32 or float64")
        self.ndim = int(ndim)
        self.device = int(device)
        self._h = C.c_void_p()
        self._live_iters = 0
        self._points = None      # host copy / device tensor of the last input (for iteration & pickling)
        self._label_map = None   # position -> original enumerate index when items were skipped
        self._cutoff = float(self.dtype.type(cutoff))
        rc = self._lib.zb_grid_create(_ffi.F32 if self.dtype == np.float32 else _ffi.F64, self.ndim, self.device,
                                      C.byref(self._h))
        if rc != _ffi.OK:
            self._h = C.c_void_p()
            raise _ffi.ZelllB200Error(rc, "zb_grid_create failed (no usable CUDA device? there is no CPU fallback)")
        # CellGrid::default() is an empty grid (cellgrid.rs:112); build it so queries are defined
        self.rebuild(particles if particles is not None else np.empty((0, self.ndim), dtype=self.dtype), cutoff)

    # -- plumbing ---------------------------------------------------------------------------
    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.zb_grid_destroy(h)
            except Exception:
                pass

    def close(self):
        self.__del__()

    def _check(self, rc: int):
        if rc != _ffi.OK:
            msg = self._lib.zb_last_error(self._h)
            raise _ffi.ZelllB200Error(rc, msg.decode() if msg else "")

    def _marshal(self, particles):
        """-> (pointer, n, keepalive, label_map)"""
        if _is_torch(particles):
            import torch

            t = particles
            want = torch.float32 if self.dtype == np.float32 else torch.float64
            if t.dtype != want:
                t = t.to(want)
            t = t.reshape(-1, self.ndim).contiguous()
            if t.is_cuda and t.device.index != self.device:
                raise ValueError(f"tensor lives on {t.device}, grid on cuda:{self.device}")
            if t.is_cuda:
                # device tensors are produced on torch's current stream: enqueue behind them
                self.use_stream(torch.cuda.current_stream(t.device).cuda_stream)
            return t.data_ptr(), t.shape[0], t, None
        if isinstance(particles, np.ndarray) and particles.dtype.kind in "fiu":
            a = np.ascontiguousarray(particles, dtype=self.dtype).reshape(-1, self.ndim)
            return a.ctypes.data, a.shape[0], a, None
        # arbitrary re-iterable of sequences (python/src/lib.rs:15-58): items that do not convert to
        # [float; ndim] are skipped but still consume an enumerate index
        rows, labels = [], []
        for i, item in enumerate(particles):
            try:
                row = [float(v) for v in item]
            except (TypeError, ValueError):
                continue
            if len(row) != self.ndim:
                continue
            rows.append(row)
            labels.append(i)
        a = np.asarray(rows, dtype=self.dtype).reshape(-1, self.ndim)
        lm = None
        if labels and labels[-1] != len(labels) - 1:
            lm = np.asarray(labels, dtype=np.uint32)
        return a.ctypes.data, a.shape[0], a, lm

    @staticmethod
    def _optional(cutoff):
        return None if cutoff is None else C.byref(C.c_double(float(cutoff)))

    # -- construction (cellgrid.rs:166-312; python/src/lib.rs:155-166) ------------------------
    def rebuild(self, particles, /, cutoff: Optional[float] = None) -> None:
        if self._live_iters:
            # PyO3's runtime borrow check (python/src/lib.rs:138-143)
            raise RuntimeError("Already borrowed")
        ptr, n, keep, lm = self._marshal(particles)
        self._check(self._lib.zb_grid_rebuild(self._h, ptr, n, self._optional(cutoff)))
        self._points, self._label_map = keep, lm
        self._view = None
        if cutoff is not None:
            self._cutoff = float(self.dtype.type(cutoff))

    def prefetch(self, particles) -> None:
        """Start copying the NEXT rebuild's (pinned) host input to the device while this grid is still being
        consumed; `rebuild(particles)` with the same array then finds it resident (zb_grid_prefetch)."""
        if isinstance(particles, np.ndarray):
            a = np.ascontiguousarray(particles, dtype=self.dtype).reshape(-1, self.ndim)
            if a.ctypes.data != particles.ctypes.data:
                return  # a converted copy would not be the array rebuild() sees
            if a.shape[0]:
                self._check(self._lib.zb_grid_prefetch(self._h, a.ctypes.data, a.shape[0]))
                # the copy is asynchronous and the slot is matched by address: keep the arrays of the two
                # staging slots alive so neither the DMA nor a later match can see reused memory.  The
                # caller must not modify `particles` until the rebuild that consumes it.
                self._prefetched = (getattr(self, "_prefetched", ()) + (particles,))[-2:]

    def prefetch_wait(self) -> None:
        """Order this grid's stream behind the copies `prefetch` has started (zb_grid_prefetch_wait)."""
        self._check(self._lib.zb_grid_prefetch_wait(self._h))

    def rebuild_mut(self, particles, cutoff: Optional[float] = None) -> None:
        """`CellGrid::rebuild_mut(&mut self, particles, Option<T>)` (cellgrid.rs:264-312)."""
        self.rebuild(particles, cutoff)

    def set_stable(self, enable: bool = True) -> None:
        """Keep the particles of a cell in input order from the next rebuild on (storage.rs:77-81)."""
        self._check(self._lib.zb_grid_set_stable(self._h, int(enable)))

    def track_key_changes(self, enable: bool = True) -> None:
        self._check(self._lib.zb_grid_track_key_changes(self._h, int(enable)))

    # -- inspection -------------------------------------------------------------------------
    def info(self) -> GridInfo:
        raw = _ffi.ZbInfo()
        self._check(self._lib.zb_grid_info(self._h, C.byref(raw)))
        return GridInfo(raw)

    def aabb(self):  # python/src/lib.rs:174-180
        inf, sup = self.info().bounding_box()
        return [float(v) for v in inf], [float(v) for v in sup]

    def cutoff(self) -> float:  # python/src/lib.rs:183-185
        return self.info().cutoff()

    def __len__(self):
        return self.info().n

    def keys(self) -> np.ndarray:
        out = np.empty(self.info().n, dtype=np.int32)
        self._check(self._lib.zb_grid_keys(self._h, out.ctypes.data))
        return out

    def neighbor_indices(self) -> np.ndarray:
        out = (C.c_int32 * 26)()
        cnt = C.c_int32(0)
        self._check(self._lib.zb_grid_neighbor_indices(self._h, out, C.byref(cnt)))
        return np.array(out[: cnt.value], dtype=np.int32)

    def cells(self):
        """Non-empty cells in ascending key order: (reference flat key, begin, count)."""
        nc = self.info().n_cells
        keys = np.empty(nc, dtype=np.int32)
        begin = np.empty(nc, dtype=np.uint32)
        count = np.empty(nc, dtype=np.uint32)
        n_out = C.c_uint64(0)
        self._check(self._lib.zb_grid_cells(self._h, keys.ctypes.data, begin.ctypes.data, count.ctypes.data, nc,
                                            C.byref(n_out)))
        return keys, begin, count

    def cell_storage(self):
        """`cell_storage()` (cellgrid.rs:412-414): (labels, coords) in cell-sorted buffer order."""
        n = self.info().n
        labels = np.empty(n, dtype=np.uint32)
        xyz = np.empty((n, self.ndim), dtype=self.dtype)
        self._check(self._lib.zb_grid_cell_storage(self._h, labels.ctypes.data, xyz.ctypes.data))
        return self._map_labels(labels), xyz

    def _map_labels(self, labels: np.ndarray) -> np.ndarray:
        return labels if self._label_map is None else self._label_map[labels]

    def _host_points(self) -> np.ndarray:
        p = self._points
        if _is_torch(p):
            p = p.detach().cpu().numpy()
            self._points = p
        return p

    def _coords_by_label(self):
        pts = self._host_points()
        if self._label_map is None:
            return lambda i: [float(v) for v in pts[i]]
        pos = {int(l): k for k, l in enumerate(self._label_map)}
        return lambda i: [float(v) for v in pts[pos[i]]]

    # -- pair enumeration and consumers ---------------------------------------------------------
    def _filter(self, cutoff, cmp):
        code = CMP[cmp]
        if cutoff is None:
            cutoff = self._cutoff
        return code, float(cutoff)

    def pair_count(self, cutoff: Optional[float] = None, cmp="none") -> int:
        """`particle_pairs().filter(..).count()` (benches/cellgrid.rs:84-88)."""
        code, fc = self._filter(cutoff, cmp)
        out = C.c_uint64(0)
        self._check(self._lib.zb_grid_pair_count(self._h, code, fc, C.addressof(out)))
        return int(out.value)

    def particle_pairs(self, cutoff: Optional[float] = None, cmp="none") -> np.ndarray:
        """`CellGrid::particle_pairs()` (cellgrid.rs:338-340), materialised: (m, 2) uint32 labels,
        home particle first; unfiltered candidates by default, distance-filtered with cmp='lt'/'le'."""
        code, fc = self._filter(cutoff, cmp)
        n_out = C.c_uint64(0)
        rc = self._lib.zb_grid_pairs(self._h, code, fc, None, 0, C.byref(n_out))
        if rc not in (_ffi.OK, _ffi.ERR_CAPACITY):
            self._check(rc)
        m = int(n_out.value)
        out = np.empty((m, 2), dtype=np.uint32)
        if m:
            self._check(self._lib.zb_grid_pairs(self._h, code, fc, out.ctypes.data, m, C.byref(n_out)))
        return self._map_labels(out)

    def particle_pairs_device(self, cutoff: Optional[float] = None, cmp="none", capacity: Optional[int] = None):
        """`particle_pairs()` materialised in HBM: a torch CUDA tensor (m, 2) int32 (the bits are uint32
        labels).  `capacity` (rows) skips the sizing call when the caller knows an upper bound."""
        import torch

        code, fc = self._filter(cutoff, cmp)
        n_out = C.c_uint64(0)
        dev = torch.device("cuda", self.device)
        self.use_stream(torch.cuda.current_stream(dev).cuda_stream)
        if capacity is None:
            rc = self._lib.zb_grid_pairs(self._h, code, fc, None, 0, C.byref(n_out))
            if rc not in (_ffi.OK, _ffi.ERR_CAPACITY):
                self._check(rc)
            capacity = int(n_out.value)
        out = torch.empty((max(int(capacity), 1), 2), dtype=torch.int32, device=dev)
        self._check(self._lib.zb_grid_pairs(self._h, code, fc, out.data_ptr(), int(capacity), C.byref(n_out)))
        return out[: int(n_out.value)]

    def par_particle_pairs(self, cutoff: Optional[float] = None, cmp="none", chunks: int = 16):
        """`par_particle_pairs()` (cellgrid.rs:447-451): the enumeration itself is parallel on the
        device; callers get the materialised list split into `chunks` slices to fan out over."""
        pairs = self.particle_pairs(cutoff, cmp)
        return np.array_split(pairs, max(1, int(chunks)))

    def lj_energy(self, cutoff: Optional[float] = None, cmp="lt", return_pairs: bool = False):
        """Fused `filter(dsq < c^2).map(lj).sum()` of benches/lj.rs:81-92."""
        code, fc = self._filter(cutoff, cmp)
        e = C.c_double(0.0)
        npairs = C.c_uint64(0)
        self._check(self._lib.zb_grid_lj_energy(self._h, code, fc, C.addressof(e), C.addressof(npairs)))
        return (e.value, int(npairs.value)) if return_pairs else e.value

    def __iter__(self) -> Iterator:  # python/src/lib.rs:168-170
        pairs = self.particle_pairs()
        return CellGridIter(pairs, self._coords_by_label(), self)

    # -- cell-level views (src/cellgrid/iters.rs:121-290) ---------------------------------------
    def _cell_view(self) -> _CellView:
        v = getattr(self, "_view", None)
        if v is None:
            v = self._view = _CellView(self)
        return v

    def iter_cells(self):
        """`CellGrid::iter()` (iters.rs:261-266): the non-empty cells (here in ascending key order)."""
        v = self._cell_view()
        return [GridCell(v, int(k)) for k in v.keys]

    def par_iter_cells(self, chunks: int = 16):
        """`CellGrid::par_iter()` (iters.rs:282-290): the same cells split into chunks to fan out over."""
        cells = self.iter_cells()
        k = max(1, int(chunks))
        return [cells[i::k] for i in range(k)]

    def query(self, coordinates):
        """`CellGrid::query()` (cellgrid.rs:360-365): the (possibly empty) cell of a point, or None."""
        info = self.info()
        idx = info.try_cell_index(coordinates)
        if idx is None:
            return None
        return GridCell(self._cell_view(), info.flatten_index(idx))

    # -- point queries (cellgrid.rs:360-401; python/src/lib.rs:204-241) ------------------------
    def query_neighbors_batch(self, queries, cutoff: Optional[float] = None, cmp="none"):
        """Batched `query_neighbors`: (offsets[nq+1], valid[nq], labels)."""
        code, fc = self._filter(cutoff, cmp)
        q = np.ascontiguousarray(queries, dtype=self.dtype).reshape(-1, self.ndim)
        nq = q.shape[0]
        offsets = np.zeros(nq + 1, dtype=np.uint64)
        valid = np.zeros(nq, dtype=np.uint8)
        n_out = C.c_uint64(0)
        rc = self._lib.zb_grid_query_neighbors(self._h, q.ctypes.data, nq, code, fc, offsets.ctypes.data,
                                               valid.ctypes.data, None, 0, C.byref(n_out))
        if rc not in (_ffi.OK, _ffi.ERR_CAPACITY):
            self._check(rc)
        m = int(n_out.value)
        labels = np.empty(m, dtype=np.uint32)
        if m:
            self._check(self._lib.zb_grid_query_neighbors(self._h, q.ctypes.data, nq, code, fc, offsets.ctypes.data,
                                                          valid.ctypes.data, labels.ctypes.data, m, C.byref(n_out)))
        return offsets, valid.astype(bool), self._map_labels(labels)

    def query_neighbors(self, coordinates):
        offsets, valid, labels = self.query_neighbors_batch([coordinates])
        if not valid[0]:
            return None
        return CellQueryIter(labels, self._coords_by_label())

    def neighbors(self, coordinates):
        """Filtered by `<= cutoff^2` like python/src/lib.rs:229-241."""
        offsets, valid, labels = self.query_neighbors_batch([coordinates], cmp="le")
        if not valid[0]:
            return None
        coords = self._coords_by_label()
        return [(int(j), coords(int(j))) for j in labels]

    # -- pickling (python/src/lib.rs:243-259): snapshot = (points, cutoff) + device rebuild -------
    def __getstate__(self):
        pts = self._host_points()
        return {"points": np.array(pts, copy=True), "cutoff": self._cutoff, "dtype": self.dtype.str,
                "ndim": self.ndim, "device": self.device, "label_map": self._label_map}

    def __setstate__(self, state):
        if not isinstance(state, dict) or "points" not in state:
            raise TypeError("invalid CellGrid state")
        self.__init__(state["points"], state["cutoff"], dtype=np.dtype(state["dtype"]), ndim=state["ndim"],
                      device=state["device"])
        self._label_map = state.get("label_map")

    # -- streams and device timing -------------------------------------------------------------
    def use_stream(self, cuda_stream: int) -> None:
        """Enqueue all later work of this grid on `cuda_stream` (a cudaStream_t as int)."""
        if getattr(self, "_stream", None) != cuda_stream:
            self._check(self._lib.zb_grid_set_stream(self._h, C.c_void_p(cuda_stream)))
            self._stream = cuda_stream

    def profile(self, enable: bool = True, stages=None) -> None:
        """Per-stage device timing on / off; `stages` (names from _ffi.STAGES) restricts the recording."""
        code = int(bool(enable))
        if enable and stages is not None:
            code = 0
            for name in stages:
                code |= 1 << (_ffi.STAGES.index(name) + 1)
        self._check(self._lib.zb_grid_profile(self._h, code))

    def profile_read(self) -> dict:
        """{stage: (summed device ms, launches)} since profile(True)."""
        ms = (C.c_double * 8)()
        cnt = (C.c_uint64 * 8)()
        self._check(self._lib.zb_grid_profile_read(self._h, ms, cnt))
        return {name: (ms[k], int(cnt[k])) for k, name in enumerate(_ffi.STAGES)}

    @property
    def launch_count(self) -> int:
        return int(self._lib.zb_grid_launch_count(self._h))
