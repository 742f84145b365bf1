"""torchrun parity check of the slab-decomposed engine on real GPUs (NCCL): every rank builds its
slab (general all-to-all path AND slab-local fast path); the all-reduced pair count / LJ energy and
the union of the sharded pair lists must equal the single-GPU grid of the whole cloud.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py [n]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zelll_b200  # noqa: E402
from zelll_b200 import workload  # noqa: E402
from zelll_b200.sharded import DistributedCellGrid, NativeSlabGrid, slab_bounds  # noqa: E402


def canonical(p):
    p = np.asarray(p).reshape(-1, 2).astype(np.uint64)
    key = (np.minimum(p[:, 0], p[:, 1]) << np.uint64(32)) | np.maximum(p[:, 0], p[:, 1])
    key.sort()
    return key


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cutoff = 10.0
    pts = workload.generate_points_random(n)
    single = zelll_b200.CellGrid(pts, cutoff, device=local)
    e_ref, m_ref = single.lj_energy(cutoff, "lt", return_pairs=True)
    c_ref = single.pair_count(cutoff, "le")
    want = canonical(single.particle_pairs(cutoff, "lt"))
    dg = DistributedCellGrid(dtype=np.float64, device=local)
    ng = NativeSlabGrid(dtype=np.float64, device=local)
    ok = True
    for mode in ("general", "slab_local", "native"):
        if mode == "general":
            mine = np.arange(rank, n, world)
            dg.rebuild(torch.from_numpy(pts[mine]).to(dev), cutoff, labels=torch.from_numpy(mine.astype(np.int64)).to(dev))
            to_global = None
        else:
            order = np.argsort(pts[:, 2], kind="stable")
            spts = pts[order]
            inf_z = spts[0, 2]
            nz = int(np.floor((spts[-1, 2] - inf_z) / cutoff)) + 1
            layer = np.floor((spts[:, 2] - inf_z) / cutoff).astype(np.int64)
            zb, ze = slab_bounds(nz, world, rank)
            sel = np.nonzero((layer >= zb) & (layer < ze))[0]
            buf = torch.zeros((len(sel) + 4096, 3), dtype=torch.float64, device=dev)
            buf[: len(sel)] = torch.from_numpy(spts[sel]).to(dev)
            off = int(sel[0]) if len(sel) else 0
            if mode == "native":
                ng.rebuild_slab_local(buf, len(sel), cutoff, label_offset=off)
            else:
                dg.rebuild_slab_local(buf, len(sel), cutoff, label_offset=off)
            to_global = order.astype(np.uint64)
        if mode == "native":
            e, m = ng.lj_energy_allreduce(cutoff, "lt", return_pairs=True)
            ct = torch.tensor([ng.pair_count(cutoff, "le")], dtype=torch.int64, device=dev)
            dist.all_reduce(ct)
            c = int(ct.item())
            local_pairs = ng.particle_pairs(cutoff, "lt")
        else:
            e, m = dg.lj_energy(cutoff, "lt", return_pairs=True)
            c = dg.pair_count(cutoff, "le")
            local_pairs = dg.local_particle_pairs(cutoff, "lt")
        if to_global is not None:
            local_pairs = to_global[local_pairs.astype(np.int64)]
        gathered = [None] * world
        dist.all_gather_object(gathered, np.asarray(local_pairs, dtype=np.uint64))
        got = canonical(np.concatenate(gathered))
        good = (m == m_ref and c == c_ref and abs(e - e_ref) <= 1e-10 * abs(e_ref) and np.array_equal(got, want))
        ok &= bool(good)
        if rank == 0:
            print(f"[dist_check] world={world} n={n} mode={mode}: pairs {m} vs {m_ref}, le-count {c} vs {c_ref}, "
                  f"energy rel diff {abs(e - e_ref) / abs(e_ref):.2e}, pair-set equal {np.array_equal(got, want)} -> "
                  f"{'OK' if good else 'MISMATCH'}")
    # native path over a SEQUENCE of frames: a repeated frame runs speculatively (box of the last step), a
    # moved frame changes the box under the speculation and must be repeated transparently on every rank
    frames = [pts, pts, workload.perturb(pts, 0, 0.3 * cutoff), None, workload.perturb(pts, 1, 0.2 * cutoff)]
    frames[3] = frames[2]
    for k, f in enumerate(frames):
        ref = zelll_b200.CellGrid(f, cutoff, device=local)
        e_ref, m_ref = ref.lj_energy(cutoff, "lt", return_pairs=True)
        order = np.argsort(f[:, 2], kind="stable")
        spts = f[order]
        inf_z = spts[0, 2]
        nz = int(np.floor((spts[-1, 2] - inf_z) / cutoff)) + 1
        layer = np.floor((spts[:, 2] - inf_z) / cutoff).astype(np.int64)
        zb, ze = slab_bounds(nz, world, rank)
        sel = np.nonzero((layer >= zb) & (layer < ze))[0]
        buf = torch.zeros((len(sel) + 4096, 3), dtype=torch.float64, device=dev)
        buf[: len(sel)] = torch.from_numpy(spts[sel]).to(dev)
        ng.rebuild_slab_local(buf, len(sel), cutoff, label_offset=int(sel[0]) if len(sel) else 0)
        e, m = ng.lj_energy_allreduce(cutoff, "lt", return_pairs=True)
        n_here = int(ng.info().n)
        good = m == m_ref and abs(e - e_ref) <= 1e-10 * abs(e_ref) and n_here >= len(sel)
        ok &= bool(good)
        if rank == 0:
            print(f"[dist_check] world={world} frame {k}: pairs {m} vs {m_ref}, energy rel diff {abs(e - e_ref) / abs(e_ref):.2e}, "
                  f"rows incl. halo {n_here} -> {'OK' if good else 'MISMATCH'}")
    # f32 grids through the native path (halo rows travel as 4 x f32): counts exact, energy to 1e-5
    n32 = min(n, 100_000)  # f32 coordinates collide in longer boxes (|z| reaches n / 18)
    p32 = workload.generate_points_random(n32, dtype=np.float32)
    ref = zelll_b200.CellGrid(p32, cutoff, dtype=np.float32, device=local)
    e_ref, m_ref = ref.lj_energy(cutoff, "lt", return_pairs=True)
    c_ref = ref.pair_count(cutoff, "le")
    ng32 = NativeSlabGrid(dtype=np.float32, device=local)
    order = np.argsort(p32[:, 2], kind="stable")
    spts = p32[order]
    inf_z = spts[0, 2]
    layer = np.floor((spts[:, 2] - inf_z) / np.float32(cutoff)).astype(np.int64)
    nz = int(layer.max()) + 1
    zb, ze = slab_bounds(nz, world, rank)
    sel = np.nonzero((layer >= zb) & (layer < ze))[0]
    buf = torch.zeros((len(sel) + 4096, 3), dtype=torch.float32, device=dev)
    buf[: len(sel)] = torch.from_numpy(spts[sel]).to(dev)
    for k in range(2):  # the second pass runs speculatively
        ng32.rebuild_slab_local(buf, len(sel), cutoff, label_offset=int(sel[0]) if len(sel) else 0)
        e, m = ng32.lj_energy_allreduce(cutoff, "lt", return_pairs=True)
        ct = torch.tensor([ng32.pair_count(cutoff, "le")], dtype=torch.int64, device=dev)
        dist.all_reduce(ct)
        good = m == m_ref and int(ct.item()) == c_ref and abs(e - e_ref) <= 1e-5 * abs(e_ref)
        ok &= bool(good)
        if rank == 0:
            print(f"[dist_check] world={world} f32 pass {k}: pairs {m} vs {m_ref}, le-count {int(ct.item())} vs {c_ref}, "
                  f"energy rel diff {abs(e - e_ref) / abs(e_ref):.2e} -> {'OK' if good else 'MISMATCH'}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


main()
