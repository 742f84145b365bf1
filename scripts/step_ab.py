"""A/B of the whole step (rebuild_mut + lj_energy) with and without per-stage event profiling."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zelll_b200
from zelll_b200 import workload

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
t = torch.from_numpy(workload.generate_points_random(n)).cuda()
cg = zelll_b200.CellGrid(t, 10.0)
def run(k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        cg.rebuild_mut(t, None)
        cg.lj_energy(10.0, "lt")
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
run(5)
for prof in (False, True, False, True):
    cg.profile(prof)
    print(f"n={n:.0e} profile={prof}: {run(30):.4f} ms/step", flush=True)
    if prof: cg.profile_read()
cg.profile(False)
