"""World-size-2/3 gloo tests (CPU) of the multi-GPU host logic in zelll_b200/sharded.py: global
bounding box all-reduce, slab ownership, all-to-all-v routing incl. the halo layer, the slab-local
fast path (halo send/recv only) and the energy / count all-reduce.  The per-rank engine is replaced
by a numpy brute-force stand-in that applies the same ownership rule as the CUDA engine (a pair is
owned by the rank owning the HIGHER z-layer of the two particles = its home cell's layer), so the
union of the ranks' pair lists must equal the oracle's pair set of the whole cloud."""
import os
import socket
import tempfile

import numpy as np
import pytest

from zelll_b200 import workload

CUTOFF = 10.0


class BruteEngine:
    """Stand-in for ShardedCellGrid with the same interface, O(n^2) numpy."""

    def __init__(self):
        self.pts = self.labels = None

    def local_aabb(self, points):
        p = points.numpy()
        if len(p) == 0:
            return np.full(3, np.inf), np.full(3, -np.inf)
        return p.min(0), p.max(0)

    def layer_of(self, points, inf_axis, cutoff, axis=None):
        p = points.numpy()
        return np.floor((p[:, 2] - inf_axis) / cutoff).astype(np.int32)

    def slab_top_layer(self, points, inf_axis, cutoff, z_begin, z_end, label_offset, halo_rows, cap_rows):
        import torch

        p = points.numpy()
        layer = np.floor((p[:, 2] - inf_axis) / cutoff).astype(np.int64)
        if len(p) and (layer.min() < z_begin or layer.max() >= z_end):
            raise ValueError("not slab-local")
        idx = np.nonzero(layer == z_end - 1)[0]
        assert len(idx) <= cap_rows
        halo_rows[1:1 + len(idx), :3] = torch.from_numpy(p[idx])
        halo_rows[1:1 + len(idx), 3] = torch.from_numpy((idx + label_offset).astype(np.int64).view(np.float64))
        return len(idx)

    def rebuild_local(self, points, labels, cutoff, inf, sup, z_begin, z_end):
        self.pts = np.asarray(points.numpy() if hasattr(points, "numpy") else points, dtype=np.float64)
        self.labels = np.asarray(labels, dtype=np.uint32)
        self.cutoff, self.inf, self.zb, self.ze = cutoff, np.asarray(inf), z_begin, z_end
        cell = np.floor((self.pts - self.inf) / cutoff).astype(np.int64)
        self.cell = cell
        assert len(cell) == 0 or (cell[:, 2].min() >= max(z_begin - 1, 0) and cell[:, 2].max() < z_end)

    def _pairs(self, cutoff, cmp):
        p, c = self.pts, self.cell
        n = len(p)
        if n < 2:
            return np.zeros((0, 2), np.uint32), np.zeros(0)
        i, j = np.triu_indices(n, 1)
        adjacent = np.all(np.abs(c[i] - c[j]) <= 1, axis=1)   # candidate pairs: neighbouring cells
        d = p[i] - p[j]
        dsq = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        keep = adjacent & (dsq < cutoff * cutoff if cmp == "lt" else dsq <= cutoff * cutoff)
        home_layer = np.maximum(c[i, 2], c[j, 2])
        keep &= (home_layer >= self.zb) & (home_layer < self.ze)
        return np.stack([self.labels[i[keep]], self.labels[j[keep]]], 1), dsq[keep]

    def particle_pairs(self, cutoff, cmp="lt"):
        return self._pairs(cutoff, cmp)[0]

    def pair_count(self, cutoff, cmp="lt"):
        return len(self._pairs(cutoff, cmp)[0])

    def lj_energy(self, cutoff, cmp="lt", return_pairs=False):
        pairs, dsq = self._pairs(cutoff, cmp)
        t = (1.0 / dsq) ** 3
        e = float(np.sum(4.0 * t * (t - 1.0)))
        return (e, len(pairs)) if return_pairs else e


def _worker(rank, world, port, mode, n, outdir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zelll_b200.sharded import DistributedCellGrid, slab_bounds

        pts = workload.generate_points_random(n)
        dg = DistributedCellGrid(engine=BruteEngine(), dtype=np.float64)
        if mode == "few_layers":
            # fewer z layers than ranks: the slab-local fast path must refuse on EVERY rank (its halo
            # chain would skip the empty slabs); the general path below handles empty slabs
            order = np.argsort(pts[:, 2], kind="stable")
            mine = np.array_split(np.arange(n), world)[rank]
            buf = torch.zeros((len(mine) + 64, 3), dtype=torch.float64)
            buf[: len(mine)] = torch.from_numpy(pts[order][mine])
            try:
                dg.rebuild_slab_local(buf, len(mine), CUTOFF, label_offset=int(mine[0]))
                raise AssertionError("slab-local path accepted fewer layers than ranks")
            except ValueError as e:
                assert "at least one layer per rank" in str(e)
        if mode in ("general", "few_layers"):
            # every rank holds an arbitrary (interleaved) subset with its global labels
            mine = np.arange(rank, n, world)
            dg.rebuild(torch.from_numpy(pts[mine]), CUTOFF, labels=torch.from_numpy(mine.astype(np.int64)))
        else:
            # slab-local: sort by z, split at layer boundaries, labels = position in the sorted cloud
            order = np.argsort(pts[:, 2], kind="stable")
            spts = pts[order]
            inf_z = spts[0, 2]
            nz = int(np.floor((spts[-1, 2] - inf_z) / CUTOFF)) + 1
            layer = np.floor((spts[:, 2] - inf_z) / CUTOFF).astype(np.int64)
            zb, ze = slab_bounds(nz, world, rank)
            sel = np.nonzero((layer >= zb) & (layer < ze))[0]
            buf = torch.zeros((len(sel) + 512, 3), dtype=torch.float64)
            buf[: len(sel)] = torch.from_numpy(spts[sel])
            dg.rebuild_slab_local(buf, len(sel), CUTOFF, label_offset=int(sel[0]) if len(sel) else 0)
        e, m = dg.lj_energy(CUTOFF, "lt", return_pairs=True)
        cnt = dg.pair_count(CUTOFF, "le")
        pairs = dg.local_particle_pairs(CUTOFF, "lt")
        np.savez(os.path.join(outdir, f"r{rank}.npz"), pairs=pairs, e=e, m=m, cnt=cnt, shape=np.array(dg.shape),
                 inf=dg.inf, sup=dg.sup)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["general", "slab_local"])
@pytest.mark.parametrize("world", [2, 3])
def test_distributed_host_logic_gloo(mode, world):
    import torch.multiprocessing as mp

    import oracle

    n = 1500
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), mode, n, d), nprocs=world, join=True)
        res = [np.load(os.path.join(d, f"r{r}.npz")) for r in range(world)]
    pts = workload.generate_points_random(n)
    if mode == "slab_local":
        pts = pts[np.argsort(pts[:, 2], kind="stable")]
    og = oracle.OracleCellGrid(pts, CUTOFF)
    info = og.info()
    for r in res:  # every rank derived the single-grid GridInfo
        assert r["shape"].tolist() == info["shape"]
        assert r["inf"].tolist() == info["inf"] and r["sup"].tolist() == info["sup"]
    want = og.pairs_canonical(oracle.CMP_LT, CUTOFF)
    got = oracle.canonical_pairs(np.concatenate([r["pairs"] for r in res]))
    assert np.array_equal(got, want)  # no pair lost, none owned twice
    _, e64, m = og.lj_energy(oracle.CMP_LT, CUTOFF)
    for r in res:  # all-reduced values agree on every rank
        assert int(r["m"]) == m
        assert abs(float(r["e"]) - e64) <= 1e-10 * abs(e64)
        assert int(r["cnt"]) == og.pair_count(oracle.CMP_LE, CUTOFF)


def test_fewer_layers_than_ranks_gloo():
    """shape[z] < world.  slab-local refuses everywhere; the general path stays exact."""
    import torch.multiprocessing as mp

    import oracle

    n, world = 150, 3  # box 30 x 30 x 16.7: two layers, three ranks
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), "few_layers", n, d), nprocs=world, join=True)
        res = [np.load(os.path.join(d, f"r{r}.npz")) for r in range(world)]
    pts = workload.generate_points_random(n)
    og = oracle.OracleCellGrid(pts, CUTOFF)
    assert og.info()["shape"][2] < world
    got = oracle.canonical_pairs(np.concatenate([r["pairs"] for r in res]))
    assert np.array_equal(got, og.pairs_canonical(oracle.CMP_LT, CUTOFF))
    _, e64, m = og.lj_energy(oracle.CMP_LT, CUTOFF)
    for r in res:
        assert int(r["m"]) == m and abs(float(r["e"]) - e64) <= 1e-10 * abs(e64)
