"""Wall-clock split of the native multi-GPU step (run under torch.distributed.run): time spent in
zb_grid_rebuild_slab_local vs zb_grid_lj_energy_allreduce per rank, next to the kernels' own time."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from bench import slab_points, CUTOFF
from zelll_b200.sharded import NativeSlabGrid

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
buf = slab_points(torch, rank, world, n, dev, 8192)
dg = NativeSlabGrid(dtype=np.float64, device=lr)
dg.use_stream(torch.cuda.current_stream(dev).cuda_stream)
for _ in range(5):
    dg.rebuild_slab_local(buf, n, CUTOFF, label_offset=rank * n)
    dg.lj_energy_allreduce(CUTOFF, "lt")
dist.barrier(); torch.cuda.synchronize()
K = 30
tr = tl = 0.0
dg.profile(True)
t00 = time.perf_counter()
for _ in range(K):
    t0 = time.perf_counter()
    dg.rebuild_slab_local(buf, n, CUTOFF, label_offset=rank * n)
    t1 = time.perf_counter()
    dg.lj_energy_allreduce(CUTOFF, "lt")
    t2 = time.perf_counter()
    tr += t1 - t0; tl += t2 - t1
tot = time.perf_counter() - t00
st = {k: round(v[0] / v[1], 4) for k, v in dg.profile_read().items() if v[1]}
print(f"rank {rank}: step {tot/K*1e3:.3f} ms  rebuild_slab_local {tr/K*1e3:.3f}  lj_allreduce {tl/K*1e3:.3f}  kernels {st} sum {sum(st.values()):.3f}", flush=True)
dist.destroy_process_group()
