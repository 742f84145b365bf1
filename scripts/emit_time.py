"""Ad-hoc: time the single-pass pair list (device buffer with head-room) and the sized two-call path."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zelll_b200
from zelll_b200 import workload

def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dtype = np.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else np.float64
    pts = workload.generate_points_random(n, dtype=dtype)
    t = torch.from_numpy(pts).cuda()
    cg = zelll_b200.CellGrid(t, 10.0, dtype=dtype)
    cap = 17 * n
    def one_pass():
        cg.rebuild(t)  # new build: nothing known about the list
        return cg.particle_pairs_device(10.0, "lt", capacity=cap).shape[0]
    def sized():
        cg.rebuild(t)
        return cg.particle_pairs_device(10.0, "lt").shape[0]
    for name, fn in (("rebuild+pairs(one pass)", one_pass), ("rebuild+pairs(size, fetch)", sized)):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        cg.profile(True)
        t0 = time.perf_counter()
        for _ in range(reps): r = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        st = {k: round(v[0] / reps, 4) for k, v in cg.profile_read().items() if v[1]}
        cg.profile(False)
        print(f"{name:28s} n={n:.0e} {dtype.__name__}: {dt*1e3:8.3f} ms  stage totals per rep (ms)={st} rows={r}")
main()
