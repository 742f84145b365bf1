"""configs[3]: z-presorted cloud, repeated rebuild_mut(None) + LJ over perturbed steps (device-resident)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zelll_b200
from zelll_b200 import workload

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pts = workload.presort_by_z(workload.generate_points_random(n))
t = torch.from_numpy(pts).cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
cg = zelll_b200.CellGrid(t, 10.0)
cg.track_key_changes(True)
cg.rebuild(t)
times, unchanged = [], 0
for s in range(steps):
    t = t + (torch.rand(t.shape, generator=g, device="cuda", dtype=torch.float64) * 2 - 1) * 1.0   # +-0.1 c
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cg.rebuild_mut(t, None)
    e = cg.lj_energy(10.0, "lt")
    times.append(time.perf_counter() - t0)
    unchanged += cg.info().keys_changed is False
times = np.array(times[5:]) * 1e3
cg.profile(True)
cg.rebuild_mut(t, None); cg.lj_energy(10.0, "lt")
st = {k: round(v[0], 4) for k, v in cg.profile_read().items() if v[1]}
print(f"presorted n={n:.0e}: {len(times)} steps, mean {times.mean():.3f} ms, median {np.median(times):.3f} ms "
      f"(incl. key tracking), keys unchanged in {unchanged}/{steps} steps; stages(ms)={st}")
