"""Ad-hoc timing of rebuild / pair_count / lj_energy with device-resident input (stage profile)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zelll_b200
from zelll_b200 import workload

def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    dtype = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
    pts = workload.generate_points_random(n, dtype=dtype)
    t = torch.from_numpy(pts).cuda()
    cg = zelll_b200.CellGrid(t, 10.0, dtype=dtype)
    res = {}
    for name, fn in [("rebuild", lambda: cg.rebuild(t)), ("pair_count_le", lambda: cg.pair_count(10.0, "le")),
                     ("pair_count_none", lambda: cg.pair_count()),
                     ("lj_energy", lambda: cg.lj_energy(10.0, "lt")),
                     ("pairs_device_lt", lambda: cg.particle_pairs_device(10.0, "lt", capacity=170_000_000 * n // 10_000_000).shape),
                     ("rebuild+lj", lambda: (cg.rebuild(t), cg.lj_energy(10.0, "lt")))]:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        cg.profile(True)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps): r = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        st = {k: round(v[0] / v[1], 4) for k, v in cg.profile_read().items() if v[1]}
        cg.profile(False)
        print(f"{name:16s} n={n:.0e} {dtype.__name__}: {dt*1e3:8.3f} ms  stages(ms)={st} result={r}")
main()
