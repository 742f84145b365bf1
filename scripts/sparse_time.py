"""Timing of the compact-cell (sparse) build and pair pass: a diffuse cloud in a huge box, and the benchmark box
forced through the sparse path (ZB_SPARSE=2) for comparison with the dense table."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zelll_b200
from zelll_b200 import workload

def bench(name, fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    print(f"{name:34s} {(time.perf_counter() - t0) / reps * 1e3:8.3f} ms  {r}")

rng = np.random.default_rng(3)
for n, side in ((100_000, 1000.0), (1_000_000, 3000.0)):
    pts = rng.random((n, 3)) * side
    t = torch.from_numpy(pts).cuda()
    cg = zelll_b200.CellGrid(t, 1.0)
    shape = cg.info().shape().astype(np.int64)
    print(f"n={n} box {shape.tolist()} = {float(np.prod(shape.astype(float))):.2e} cells, non-empty {cg.info().n_cells}")
    bench(f"  rebuild (n={n:.0e})", lambda: cg.rebuild(t))
    bench(f"  pair_count le", lambda: cg.pair_count(1.0, "le"))
    bench(f"  lj_energy", lambda: cg.lj_energy(1.0, "lt"))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pts = workload.generate_points_random(n)
t = torch.from_numpy(pts).cuda()
cg = zelll_b200.CellGrid(t, 10.0)
print(f"benchmark box n={n}, ZB_SPARSE={os.environ.get('ZB_SPARSE')}, non-empty cells {cg.info().n_cells}")
bench("  rebuild", lambda: cg.rebuild(t))
bench("  pair_count le", lambda: cg.pair_count(10.0, "le"))
bench("  lj_energy", lambda: cg.lj_energy(10.0, "lt"))
