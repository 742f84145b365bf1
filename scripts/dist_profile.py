"""Host-side timeline of one distributed step (torchrun, 2+ GPUs): where the per-step overhead goes."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from zelll_b200.sharded import DistributedCellGrid

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_per = 10_000_000
buf = bench.slab_points(torch, rank, world, n_per, dev, 8192)
dg = DistributedCellGrid(dtype=np.float64, device=local)
dg.engine.use_stream(torch.cuda.current_stream(dev).cuda_stream)

def T():
    torch.cuda.synchronize()
    return time.perf_counter()

for it in range(6):
    marks = []
    t0 = T()
    dg._global_box(buf[:n_per], 10.0); marks.append(("global_box", T()))
    n_top = dg.engine.slab_top_layer(buf[:n_per], float(dg.inf[-1]), dg.cutoff, dg.z_begin, dg.z_end, rank * n_per, dg._halo_send if dg._halo_send is not None else torch.zeros((8193, 4), dtype=torch.float64, device=dev), 8192); marks.append(("top_layer", T()))
    dg.rebuild_slab_local(buf, n_per, 10.0, rank * n_per, box=(dg.inf, dg.sup)); marks.append(("slab_local(box given)", T()))
    e = dg.engine.lj_energy(10.0, "lt", return_pairs=True); marks.append(("engine.lj", T()))
    e = dg._allreduce_sum([e[0], float(e[1])], torch.float64); marks.append(("allreduce", T()))
    t1 = T()
    dg.rebuild_slab_local(buf, n_per, 10.0, rank * n_per); full = T() - t1
    if rank == 0 and it >= 3:
        prev = t0
        s = []
        for name, t in marks:
            s.append(f"{name}={1e3 * (t - prev):.3f}")
            prev = t
        print("ms:", " ".join(s), f"| full rebuild_slab_local={1e3 * full:.3f}")
dist.destroy_process_group()
