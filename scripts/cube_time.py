"""Timing on a CUBIC box (the common MD geometry): n particles at 10 per cutoff^3 in a cube; ZB_ROW_TILES=0
reads records through L1/L2 (round-1 behaviour for wide grids), default stages five row segments per tile."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zelll_b200
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
side = (n / 10.0) ** (1.0 / 3.0) * 10.0
pts = np.random.default_rng(5).random((n, 3)) * side
t = torch.from_numpy(pts).cuda()
cg = zelll_b200.CellGrid(t, 10.0)
print(f"cube n={n}, shape {cg.info().shape().tolist()}, ZB_ROW_TILES={os.environ.get('ZB_ROW_TILES')} ZB_PREFILTER={os.environ.get('ZB_PREFILTER')}")
for name, fn in (("rebuild", lambda: cg.rebuild(t)), ("pair_count le", lambda: cg.pair_count(10.0, "le")),
                 ("lj_energy", lambda: cg.lj_energy(10.0, "lt")),
                 ("pairs (count+emit)", lambda: cg.particle_pairs_device(10.0, "lt", capacity=22 * n).shape[0])):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): r = fn()
    torch.cuda.synchronize()
    print(f"  {name:20s} {(time.perf_counter() - t0) / 10 * 1e3:8.3f} ms  {r}")
