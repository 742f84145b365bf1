"""Slab-decomposed CellGrid across the GPUs of one node.

The reference has nothing distributed (its only parallelism is rayon over cells,
src/cellgrid/iters.rs:282-290); this is new work behind the same API.  One process per GPU:

1. local `Aabb::from_particles` (util.rs:35-52) then all-reduce(min, max): every rank derives the
   SAME GridInfo (util.rs:191-220), hence the same keys and the same pair set as one big grid;
2. z-layers (the slowest axis, largest stride: util.rs:200-212) are split into `world` slabs;
3. one all-to-all-v routes each particle to the rank owning its layer AND to the rank whose lower
   halo layer it is (the z-major half shell only looks at layers z-1 and z); slab-local input
   (presorted or generated per slab) only moves its top layer: a single send/recv over NVLink;
4. each unordered pair is owned by the rank owning its home cell -- no double counting;
5. all-reduce(sum) of the f64 energy and of the pair count.  Pair lists stay sharded.

`ShardedCellGrid` is the per-rank engine (zb_grid_rebuild_sharded); `DistributedCellGrid` is the
torch.distributed orchestration and takes the engine as a parameter so its routing logic is
testable on CPU (gloo) without a GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _ffi
from .cellgrid import CMP, CellGrid, _is_torch


def slab_bounds(n_layers: int, world: int, rank: int):
    """Layers [begin, end) owned by `rank`: equal layer counts (the benchmark box is uniform)."""
    return (rank * n_layers) // world, ((rank + 1) * n_layers) // world


def grid_shape(inf, sup, cutoff, dtype) -> list:
    """`GridInfo::new` shape (util.rs:198) in the arithmetic of `dtype`."""
    T = np.dtype(dtype).type
    inf = np.asarray(inf, dtype=dtype)
    sup = np.asarray(sup, dtype=dtype)
    return [int(np.floor((s - i) / T(cutoff))) + 1 for i, s in zip(inf, sup)]


class ShardedCellGrid(CellGrid):
    """One rank's slab: home layers [z_begin, z_end) plus the lower halo layer."""

    def __init__(self, *, dtype=np.float64, ndim: int = 3, device: int = 0, cutoff: float = 1.0):
        super().__init__(None, cutoff, dtype=dtype, ndim=ndim, device=device)
        self._keep = None

    # -- engine interface used by DistributedCellGrid ---------------------------------------
    def local_aabb(self, points):
        ptr, n, keep, _ = self._marshal(points)
        out = (C.c_double * 6)()
        self._check(self._lib.zb_aabb(self._h, ptr, n, out))
        nd = self.ndim
        if n == 0:
            return np.full(nd, np.inf), np.full(nd, -np.inf)
        return np.array(out[0:nd]), np.array(out[3:3 + nd])

    def local_aabb_into(self, points, out6) -> None:
        """Asynchronous: (inf xyz, sup xyz) as 6 doubles into the CUDA tensor `out6` (+-inf when empty)."""
        ptr, n, keep, _ = self._marshal(points)
        if n == 0:
            out6[:3] = float("inf")
            out6[3:] = float("-inf")
            return
        self._check(self._lib.zb_aabb(self._h, ptr, n, C.cast(out6.data_ptr(), C.POINTER(C.c_double))))
        if self.ndim < 3:
            out6[self.ndim:3] = float("inf")
            out6[3 + self.ndim:] = float("-inf")

    def lj_energy_into(self, cutoff: float, cmp, energy_out, pairs_out) -> None:
        """Asynchronous fused LJ pass: energy (f64[1]) and kept pairs (int64[1]) stay on the device."""
        self._check(self._lib.zb_grid_lj_energy(self._h, CMP[cmp], float(cutoff), energy_out.data_ptr(), pairs_out.data_ptr()))

    def layer_of(self, points, inf_axis: float, cutoff: float, axis: Optional[int] = None):
        axis = self.ndim - 1 if axis is None else axis
        ptr, n, keep, _ = self._marshal(points)
        if _is_torch(points) and points.is_cuda:
            import torch

            out = torch.empty(n, dtype=torch.int32, device=points.device)
            self._check(self._lib.zb_layer_of(self._h, ptr, n, float(inf_axis), float(cutoff), axis, out.data_ptr()))
            return out
        out = np.empty(n, dtype=np.int32)
        self._check(self._lib.zb_layer_of(self._h, ptr, n, float(inf_axis), float(cutoff), axis, out.ctypes.data))
        if _is_torch(points):
            import torch

            return torch.from_numpy(out)
        return out

    def slab_top_layer(self, points, inf_axis: float, cutoff: float, z_begin: int, z_end: int, label_offset: int,
                       halo_rows, cap_rows: int) -> int:
        """Check slab-locality of `points` and compact their top layer into halo_rows[1:] (device tensor of
        shape [cap_rows + 1, 4]); returns the number of rows written."""
        ptr, n, keep, _ = self._marshal(points)
        if halo_rows.is_cuda:
            # asynchronous: the row count travels in the block header, the slab check is reported by
            # the rebuild that follows
            self._check(self._lib.zb_slab_top_layer(self._h, ptr, n, float(inf_axis), float(cutoff), int(z_begin),
                                                    int(z_end), int(label_offset) & 0xFFFFFFFF, halo_rows.data_ptr(),
                                                    int(cap_rows), None, None))
            return -1
        raise ValueError("halo_rows must be a CUDA tensor")

    def rebuild_local(self, points, labels, cutoff, inf, sup, z_begin: int, z_end: int) -> None:
        ptr, n, keep, _ = self._marshal(points)
        lab_keep, lptr = None, None
        if labels is not None:
            if _is_torch(labels):
                import torch

                lab_keep = labels.to(torch.int32).contiguous() if labels.dtype not in (torch.int32, torch.uint32) else labels.contiguous()
                lptr = lab_keep.data_ptr()
            else:
                lab_keep = np.ascontiguousarray(labels, dtype=np.uint32)
                lptr = lab_keep.ctypes.data
        a_inf = (C.c_double * 3)(*([float(v) for v in inf] + [0.0] * (3 - len(inf))))
        a_sup = (C.c_double * 3)(*([float(v) for v in sup] + [0.0] * (3 - len(sup))))
        self._check(self._lib.zb_grid_rebuild_sharded(self._h, ptr, n, lptr, self._optional(cutoff), a_inf, a_sup,
                                                      int(z_begin), int(z_end)))
        self._points, self._label_map, self._keep = keep, None, lab_keep
        self._view = None
        if cutoff is not None:
            self._cutoff = float(self.dtype.type(cutoff))


def nccl_library_path() -> str:
    """The NCCL shared library this process already uses (torch's bundled one), for zb_comm_init."""
    import glob
    import os

    import torch

    site = os.path.dirname(os.path.dirname(torch.__file__))
    for pat in (os.path.join(site, "nvidia", "nccl", "lib", "libnccl.so*"),
                os.path.join(os.path.dirname(torch.__file__), "lib", "libnccl.so*")):
        hits = sorted(glob.glob(pat))
        if hits:
            return hits[0]
    return "libnccl.so.2"


class NativeSlabGrid(ShardedCellGrid):
    """Slab-local multi-GPU step driven by the library itself (zb_comm_init, zb_grid_rebuild_slab_local,
    zb_grid_lj_energy_allreduce): NCCL is called from C on the handle's stream, three host round trips
    per step.  torch.distributed is used once, to hand rank 0's NCCL unique id to the other ranks."""

    def __init__(self, *, dtype=np.float64, ndim: int = 3, device: int = 0, group=None):
        import torch.distributed as dist

        super().__init__(dtype=dtype, ndim=ndim, device=device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        path = nccl_library_path().encode()
        uid = C.create_string_buffer(128)
        if self.rank == 0:
            rc = self._lib.zb_comm_unique_id(path, uid)
            if rc != _ffi.OK:
                raise _ffi.ZelllB200Error(rc, "ncclGetUniqueId failed")
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        uid = C.create_string_buffer(box[0], 128)
        self._check(self._lib.zb_comm_init(self._h, path, uid, self.world, self.rank))
        self.slab = None

    def rebuild_slab_local(self, buf, n_local: int, cutoff: Optional[float], label_offset: int, halo_cap: int = 8192):
        """buf: CUDA tensor [cap_rows, ndim]; rows [0, n_local) are this rank's own layers, the halo is
        appended behind them.  Returns the zb_slab_info of the step."""
        if not (_is_torch(buf) and buf.is_cuda):
            raise ValueError("buf must be a CUDA tensor")
        import torch

        self.use_stream(torch.cuda.current_stream(buf.device).cuda_stream)
        info = _ffi.ZbSlabInfo()
        self._check(self._lib.zb_grid_rebuild_slab_local(self._h, buf.data_ptr(), int(n_local), int(buf.shape[0]),
                                                         self._optional(cutoff), int(label_offset) & 0xFFFFFFFF,
                                                         int(halo_cap), C.byref(info)))
        self._points, self._label_map = buf, None
        self._view = None
        if cutoff is not None:
            self._cutoff = float(self.dtype.type(cutoff))
        self.slab = info
        return info

    @property
    def n_halo(self) -> int:
        """Halo rows of the last step (counted on the device; asking synchronises the step)."""
        return int(self.info().n) - int(self.slab.n_local) if self.slab is not None else 0

    def lj_energy_allreduce(self, cutoff: Optional[float] = None, cmp="lt", return_pairs: bool = False):
        code, fc = self._filter(cutoff, cmp)
        e, m = C.c_double(0.0), C.c_uint64(0)
        self._check(self._lib.zb_grid_lj_energy_allreduce(self._h, code, fc, C.byref(e), C.byref(m)))
        return (e.value, int(m.value)) if return_pairs else e.value


class DistributedCellGrid:
    """torch.distributed orchestration of one ShardedCellGrid per rank."""

    def __init__(self, engine=None, *, dtype=np.float64, ndim: int = 3, device: Optional[int] = None, group=None):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dtype = np.dtype(dtype)
        self.ndim = ndim
        self.backend = dist.get_backend(group)
        if engine is None:
            engine = ShardedCellGrid(dtype=dtype, ndim=ndim, device=0 if device is None else device)
        self.engine = engine
        self.comm_device = torch.device("cuda", 0 if device is None else device) if self.backend == "nccl" else torch.device("cpu")
        self.inf = self.sup = None
        self.shape = None
        self.cutoff = None
        self.z_begin = self.z_end = 0
        self.n_home = 0
        self.n_halo = 0
        self._halo_send = self._halo_recv = None
        self._labels, self._labels_key = None, None
        self._box6 = None
        self._red = self._red_pairs = None

    # -- helpers ---------------------------------------------------------------------------------
    def _tdtype(self):
        return self._torch.float32 if self.dtype == np.float32 else self._torch.float64

    def _global_box(self, points, cutoff):
        torch, dist = self._torch, self._dist
        nd = self.ndim
        if hasattr(self.engine, "local_aabb_into") and getattr(points, "is_cuda", False):
            # device path: bbox kernel -> negate inf -> all-reduce(max) -> ONE host read
            if self._box6 is None:
                self._box6 = torch.empty(6, dtype=torch.float64, device=points.device)
            self.engine.local_aabb_into(points, self._box6)
            self._box6[:3].neg_()
            dist.all_reduce(self._box6, op=dist.ReduceOp.MAX, group=self.group)
            b = self._box6.cpu().numpy()
            inf, sup = -b[:nd], b[3:3 + nd]
        else:
            lo, hi = self.engine.local_aabb(points)
            buf = torch.tensor(np.concatenate([-np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)]),
                               dtype=torch.float64, device=self.comm_device)
            dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=self.group)  # max(-inf_d) = -min(inf_d)
            buf = buf.cpu().numpy()
            inf, sup = -buf[:nd], buf[nd:]
        if not np.all(np.isfinite(inf)):  # no particles anywhere: Aabb of an empty set is zeros (util.rs:41)
            inf, sup = np.zeros(nd), np.zeros(nd)
        self.inf, self.sup, self.cutoff = inf, sup, float(self.dtype.type(cutoff))
        self.shape = grid_shape(inf, sup, cutoff, self.dtype)
        self.z_begin, self.z_end = slab_bounds(self.shape[-1], self.world, self.rank)

    def _bounds(self):
        nz = self.shape[-1]
        return [slab_bounds(nz, self.world, r) for r in range(self.world)]

    # -- construction ------------------------------------------------------------------------------
    def rebuild(self, points, cutoff: float, labels=None, label_offset: int = 0):
        """General input: every rank holds an arbitrary subset of the particles (torch tensor on the
        communication device).  `labels` (or label_offset + local index) are the global labels."""
        torch, dist = self._torch, self._dist
        points = points.reshape(-1, self.ndim).to(self._tdtype()).contiguous()
        n = points.shape[0]
        if labels is None:
            labels = torch.arange(label_offset, label_offset + n, dtype=torch.int64, device=points.device)
        labels = labels.to(torch.int64)
        self._global_box(points, cutoff)
        layer = self.engine.layer_of(points, float(self.inf[-1]), self.cutoff)
        layer = torch.as_tensor(layer).to(points.device).to(torch.int64)
        # destination lists: rank q receives its home layers and its lower halo layer
        send_idx, counts = [], []
        for (zb_, ze_) in self._bounds():
            if ze_ <= zb_:
                sel = torch.zeros(0, dtype=torch.int64, device=points.device)
            else:
                sel = torch.nonzero((layer >= zb_ - 1) & (layer < ze_)).reshape(-1)
            send_idx.append(sel)
            counts.append(int(sel.numel()))
        order = torch.cat(send_idx) if send_idx else torch.zeros(0, dtype=torch.int64)
        send_pts = points[order].contiguous()
        send_lab = labels[order].contiguous()
        cnt = torch.tensor(counts, dtype=torch.int64, device=self.comm_device)
        rcnt = torch.empty_like(cnt)
        dist.all_to_all_single(rcnt, cnt, group=self.group)
        rcounts = [int(v) for v in rcnt.cpu().tolist()]
        recv_pts = torch.empty((sum(rcounts), self.ndim), dtype=points.dtype, device=points.device)
        recv_lab = torch.empty(sum(rcounts), dtype=torch.int64, device=points.device)
        dist.all_to_all_single(recv_pts, send_pts, output_split_sizes=rcounts, input_split_sizes=counts, group=self.group)
        dist.all_to_all_single(recv_lab, send_lab, output_split_sizes=rcounts, input_split_sizes=counts, group=self.group)
        self._finish(recv_pts, recv_lab)

    def rebuild_slab_local(self, buf, n_local: int, cutoff: float, label_offset: int, box=None, halo_cap: int = 8192):
        """Fast path: rank r already holds exactly the particles of its own layers in buf[:n_local]
        (generated per slab, or presorted and split at layer boundaries); buf has spare rows at the
        tail that receive the lower halo layer from rank r-1 -- the only NVLink traffic: ONE
        fixed-size send/recv of {count; x, y, z, label} rows per neighbour pair.
        `box` = (inf, sup) skips the bounding-box all-reduce when the global box is known."""
        torch, dist = self._torch, self._dist
        points = buf[:n_local]
        if box is None:
            self._global_box(points, cutoff)
        else:
            self.inf, self.sup = np.asarray(box[0], dtype=np.float64), np.asarray(box[1], dtype=np.float64)
            self.cutoff = float(self.dtype.type(cutoff))
            self.shape = grid_shape(self.inf, self.sup, cutoff, self.dtype)
            self.z_begin, self.z_end = slab_bounds(self.shape[-1], self.world, self.rank)
        if self.shape[-1] < self.world:
            # the lower halo comes from rank - 1's top layer only: an empty slab would cut the chain
            # (the general all-to-all path, rebuild(), has no such limit)
            raise ValueError(f"slab-local input needs at least one layer per rank: {self.shape[-1]} layers, {self.world} ranks")
        up, down = self.rank + 1, self.rank - 1
        if self._halo_send is None or self._halo_send.shape[0] != halo_cap + 1 or self._halo_send.device != buf.device:
            self._halo_send = torch.zeros((halo_cap + 1, 4), dtype=buf.dtype, device=buf.device)
            self._halo_recv = torch.zeros((halo_cap + 1, 4), dtype=buf.dtype, device=buf.device)
        # top layer of this slab -> halo block (one engine kernel; also verifies slab-locality)
        n_top = self.engine.slab_top_layer(points, float(self.inf[-1]), self.cutoff, self.z_begin, self.z_end,
                                           label_offset, self._halo_send, halo_cap)
        if n_top >= 0:  # engines without an on-device header write the count here
            self._halo_send[0, 0] = float(n_top)
        ops = []
        if up < self.world:
            ops.append(dist.P2POp(dist.isend, self._halo_send, up, group=self.group))
        if down >= 0:
            ops.append(dist.P2POp(dist.irecv, self._halo_recv, down, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        n_halo = int(self._halo_recv[0, 0].item()) if down >= 0 else 0
        if n_halo > halo_cap:
            raise ValueError(f"the neighbour's top layer exceeds halo_cap={halo_cap} rows")
        if n_local + n_halo > buf.shape[0]:
            raise ValueError(f"halo of {n_halo} rows does not fit the {buf.shape[0] - n_local} spare rows of buf")
        # persistent label array: local labels are written once, only the halo tail changes per step
        key = (int(buf.shape[0]), int(n_local), int(label_offset), str(buf.device))
        if self._labels_key != key:
            self._labels = torch.empty(buf.shape[0], dtype=torch.int32, device=buf.device)
            self._labels[:n_local] = (torch.arange(n_local, dtype=torch.int64, device=buf.device) + label_offset).to(torch.int32)
            self._labels_key = key
        if n_halo:
            rows = self._halo_recv[1:1 + n_halo]
            buf[n_local:n_local + n_halo] = rows[:, :3]
            bits = rows[:, 3].contiguous().view(torch.int64 if buf.dtype == torch.float64 else torch.int32)
            self._labels[n_local:n_local + n_halo] = bits.to(torch.int32)
        self.n_total_local = n_local + n_halo
        self.engine.rebuild_local(buf[:n_local + n_halo], self._labels[:n_local + n_halo], self.cutoff, self.inf, self.sup,
                                  self.z_begin, self.z_end)

    def _finish(self, pts, labels):
        torch = self._torch
        self.n_total_local = int(pts.shape[0])
        lab32 = labels.to(torch.int64).cpu().numpy().astype(np.uint32) if not pts.is_cuda else labels.to(torch.int32)
        # int64 -> int32 keeps the low 32 bits: labels are u32 on the device
        self.engine.rebuild_local(pts, lab32, self.cutoff, self.inf, self.sup, self.z_begin, self.z_end)

    # -- consumers -----------------------------------------------------------------------------------
    def _allreduce_sum(self, values, dtype):
        torch, dist = self._torch, self._dist
        t = torch.tensor(values, dtype=dtype, device=self.comm_device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().tolist()

    def pair_count(self, cutoff: Optional[float] = None, cmp="none") -> int:
        local = self.engine.pair_count(self.cutoff if cutoff is None else cutoff, cmp)
        return int(self._allreduce_sum([local], self._torch.int64)[0])

    def lj_energy(self, cutoff: Optional[float] = None, cmp="lt", return_pairs: bool = False):
        torch, dist = self._torch, self._dist
        fc = self.cutoff if cutoff is None else cutoff
        if hasattr(self.engine, "lj_energy_into") and self.backend == "nccl":
            # device path: the kernel's results are all-reduced where they are; ONE host read.
            # (the pair count is exact in f64 below 2^53)
            if self._red is None:
                self._red = torch.zeros(2, dtype=torch.float64, device=self.comm_device)
                self._red_pairs = torch.zeros(1, dtype=torch.int64, device=self.comm_device)
            self.engine.lj_energy_into(fc, cmp, self._red[0:1], self._red_pairs)
            self._red[1] = self._red_pairs[0].to(torch.float64)
            dist.all_reduce(self._red, op=dist.ReduceOp.SUM, group=self.group)
            e_all, m_all = self._red.cpu().tolist()
        else:
            e, m = self.engine.lj_energy(fc, cmp, return_pairs=True)
            e_all, m_all = self._allreduce_sum([e, float(m)], torch.float64)
        return (e_all, int(m_all)) if return_pairs else e_all

    def local_particle_pairs(self, cutoff: Optional[float] = None, cmp="none") -> np.ndarray:
        """This rank's shard of the pair list (global labels); the list stays sharded."""
        return self.engine.particle_pairs(self.cutoff if cutoff is None else cutoff, cmp)
