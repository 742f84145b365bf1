"""Small driver for ncu captures: build the n-particle benchmark grid, run one op a few times."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import zelll_b200
from zelll_b200 import workload

op = sys.argv[1] if len(sys.argv) > 1 else "lj"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
dtype = np.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else np.float64
pts = workload.generate_points_random(n, dtype=dtype)
t = torch.from_numpy(pts).cuda()
cg = zelll_b200.CellGrid(t, 10.0, dtype=dtype)
for _ in range(3):
    if op == "lj":
        r = cg.lj_energy(10.0, "lt")
    elif op == "count":
        r = cg.pair_count(10.0, "le")
    elif op == "rebuild":
        r = cg.rebuild(t)
    elif op == "pairs":
        r = len(cg.particle_pairs(10.0, "lt"))
torch.cuda.synchronize()
print(op, n, r)
