"""Small end-to-end run touching every kernel of the engine (smoke; compute-sanitizer is closed on this pool)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import zelll_b200
from zelll_b200 import workload, sharded

for dtype in (np.float64, np.float32):
    for n, kind in ((3000, "lj"), (2500, "dense"), (2000, "flat"), (33, "tiny")):
        rng = np.random.default_rng(n)
        if kind == "lj":
            pts, c = workload.generate_points_random(n, dtype=dtype), 10.0
        elif kind == "dense":
            pts, c = (rng.random((n, 3)) * 2.0).astype(dtype), 1.0
        elif kind == "flat":
            pts, c = (rng.random((n, 3)) * [40.0, 40.0, 1.5]).astype(dtype), 1.0
        else:
            pts, c = (rng.random((n, 3)) * 3.0).astype(dtype), 1.0
        cg = zelll_b200.CellGrid(pts, c, dtype=dtype)
        cg.track_key_changes(True)
        cg.rebuild(pts)
        cg.keys(); cg.cells(); cg.cell_storage(); cg.neighbor_indices()
        for cmp in ("none", "lt", "le"):
            cg.pair_count(c, cmp)
            cg.particle_pairs(c, cmp)
        cg.lj_energy(c, "lt")
        cg.query_neighbors_batch(pts[:50] + 0.3, c, "le")
        inf, sup = cg.info().bounding_box()
        nz = int(cg.info().shape()[2])
        layer = np.floor((pts[:, 2] - inf[2]) / dtype(c)).astype(np.int64)
        for r in range(2):
            zb, ze = sharded.slab_bounds(nz, 2, r)
            sel = (layer >= max(zb - 1, 0)) & (layer < ze)
            g = sharded.ShardedCellGrid(dtype=dtype)
            g.rebuild_local(pts[sel], np.nonzero(sel)[0].astype(np.uint32), c, inf, sup, zb, ze)
            g.pair_count(c, "le"); g.lj_energy(c, "lt")
pts2 = np.random.default_rng(0).random((500, 2))
cg = zelll_b200.CellGrid(pts2, 0.1, ndim=2)
cg.particle_pairs(0.1, "le"); cg.lj_energy(0.1, "lt")
print("small e2e done")
