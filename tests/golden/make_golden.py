"""Writes tests/golden/reference_known_answers.json.

The reference (zelll v0.5.0, Rust) cannot be built or imported in this image, so these goldens
are TRANSCRIBED from the reference's own unit tests and doctests -- every literal below cites
the reference file:line it was copied from.  Nothing here is computed by our code.

    python tests/golden/make_golden.py
"""
import json
import os

G = {}

# src/cellgrid/util.rs:346-379  test_generate_pointcloud
G["generate_pointcloud_3x3x3_unit_origin0"] = {
    "cite": "src/cellgrid/util.rs:346-379",
    "shape": [3, 3, 3], "cutoff": 1.0, "origin": [0.0, 0.0, 0.0],
    "points": [
        [0.0, 0.0, 0.0], [0.5, 0.5, 0.5], [0.0, 0.0, 2.0], [0.5, 0.5, 2.5],
        [0.0, 1.0, 1.0], [0.5, 1.5, 1.5], [0.0, 2.0, 0.0], [0.5, 2.5, 0.5],
        [0.0, 2.0, 2.0], [0.5, 2.5, 2.5], [1.0, 0.0, 1.0], [1.5, 0.5, 1.5],
        [1.0, 1.0, 0.0], [1.5, 1.5, 0.5], [1.0, 1.0, 2.0], [1.5, 1.5, 2.5],
        [1.0, 2.0, 1.0], [1.5, 2.5, 1.5], [2.0, 0.0, 0.0], [2.5, 0.5, 0.5],
        [2.0, 0.0, 2.0], [2.5, 0.5, 2.5], [2.0, 1.0, 1.0], [2.5, 1.5, 1.5],
        [2.0, 2.0, 0.0], [2.5, 2.5, 0.5], [2.0, 2.0, 2.0], [2.5, 2.5, 2.5],
    ],
}

# src/cellgrid/util.rs:381-430  test_utils
G["test_utils"] = {
    "cite": "src/cellgrid/util.rs:381-430",
    "shape": [3, 3, 3], "cutoff": 1.0, "origin": [0.2, 0.25, 0.3],
    "n_points": 28,
    "aabb_inf": [0.2, 0.25, 0.3],
    "aabb_sup": [2.7, 2.75, 2.8],
    "grid_shape": [3, 3, 3],
    "grid_strides": [1, 7, 49],
    "cell_index_cases": [
        # 2.3 - 0.3 = 1.9999999999999998 -> z cell 1, not 2 (util.rs:406-414)
        {"p": [2.7, 2.75, 2.3], "cell": [2, 2, 1], "flat": 65},
        {"p": [2.7, 2.75, 2.8], "cell": [2, 2, 2], "flat": 114},
    ],
}

# src/cellgrid/flatindex.rs:161-171  test_neighbor_indices (2-D, 8x8 padded chessboard)
G["test_neighbor_indices_2d"] = {
    "cite": "src/cellgrid/flatindex.rs:161-171",
    "points": [[0.0, 0.0], [3.0, 3.0]], "cutoff": 1.0,
    "neighbor_indices": [-9, -1, 7, -8, 8, -7, 1, 9],
}

# src/cellgrid/flatindex.rs:173-192  test_flatindex: keys == flatten_index([x,y,z]) twice per
# even cell, x slowest / z fastest, on the 3x3x3 chessboard at origin 0
G["test_flatindex"] = {
    "cite": "src/cellgrid/flatindex.rs:173-192",
    "shape": [3, 3, 3], "cutoff": 1.0, "origin": [0.0, 0.0, 0.0],
    "rule": "for x,y,z in 0..3 (z fastest) if (x+y+z)%2==0: push flatten([x,y,z]) twice",
}

# src/cellgrid/iters.rs:298-331  test_cellgrid_iter / test_gridcell_iter
G["test_cellgrid_iter"] = {
    "cite": "src/cellgrid/iters.rs:298-331",
    "shape": [3, 3, 3], "cutoff": 1.0, "origin": [0.0, 0.0, 0.0],
    "nonempty_cells": 14,
    "sum_cell_sizes_equals_n": True,
}

# src/cellgrid/iters.rs:333-356  test_neighborcell_particle_pairs
G["test_neighborcell_particle_pairs"] = {
    "cite": "src/cellgrid/iters.rs:333-356",
    "shape": [2, 2, 2], "cutoff": 1.0, "origin": [0.0, 0.0, 0.0],
    "intra_half": 4,
    "inter_half": 24,
}

# src/cellgrid/iters.rs:358-387  test_half_full_space_particle_pairs
G["test_half_full_space_particle_pairs"] = {
    "cite": "src/cellgrid/iters.rs:358-387",
    "shape": [2, 2, 2], "cutoff": 1.0, "origin": [0.0, 0.0, 0.0],
    "full_over_half_intra": 2,
    "full_over_half_inter": 2,
}

# src/cellgrid/util.rs:268-286 doctests of flat_cell_index / cell_index
G["doctest_flat_cell_index"] = {
    "cite": "src/cellgrid/util.rs:268-286",
    "points": [[0.0, 0.0, 0.0], [1.0, 2.0, 0.0], [0.0, 0.1, 0.2]], "cutoff": 1.0,
    "p_ok": [-1.0, -1.0, -1.0],       # flat_cell_index(p) == flatten_index(cell_index(p))
    "p_panics": [-2.0, -2.0, -2.0],   # cell_index(p) panics == try_cell_index(p) is None
}

# src/cellgrid.rs:72-111, :322-336, :374-389 doctests: the 3-point cloud every doctest uses;
# src/cellgrid/iters.rs:274-280: iter().count() == par_iter().count(); :256-259: sum of cell sizes == n
G["doctest_three_points"] = {
    "cite": "src/cellgrid.rs:72-111,322-336,374-389; src/cellgrid/iters.rs:256-259,274-280",
    "points": [[0.0, 0.0, 0.0], [1.0, 2.0, 0.0], [0.0, 0.1, 0.2]], "cutoff": 1.0,
    "query_point": [0.5, 1.0, 0.1],   # query_neighbors(p) is Some (cellgrid.rs:381-383)
    "points_2d": [[0.0, 0.0], [1.0, 2.0], [0.0, 0.1]],
}

# surface-sampling/src/sdf/numdual.rs:107-192  test_sdf_autodiff: the one reference-held FLOATING-POINT
# known answer on this path.  sdf(x) = -sigma * ln(sum exp(-d/r)) over query_neighbors(x) filtered by
# d <= cutoff (numdual.rs:17-58), sigma = sum(r exp(-d)) / sum(exp(-d)); every atom is
# Element::default() = Carbon, radius 1.70 (atom.rs:3-6, 17-21); CellGrid cutoff 1.0 (numdual.rs:176).
G["test_sdf_autodiff"] = {
    "cite": "surface-sampling/src/sdf/numdual.rs:107-192; surface-sampling/src/atom.rs:3-28",
    "cutoff": 1.0, "radius": 1.70,
    "points": [
        [0.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [1.0, 1.0, 0.0],
        [0.0, 1.0, 1.0], [1.0, 0.0, 1.0], [1.0, 1.0, 1.0], [0.5, 0.5, 0.5], [1.5, 1.5, 1.5],
    ],
    "reference_values": [
        -2.012457244274712, -2.012457244274712, -2.012457244274712, -2.012457244274712,
        -2.012457244274712, -2.012457244274712, -2.012457244274712, -2.2994776285300675,
        -2.990326826730122, -0.7998983683589523,
    ],
}

here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "reference_known_answers.json"), "w") as f:
    json.dump(G, f, indent=1, sort_keys=True)
    f.write("\n")
print("wrote", len(G), "golden groups")
