"""Pins the CPU oracle (oracle/zelll_oracle.cpp) against the reference's own known answers
(tests/golden/reference_known_answers.json, transcribed from /root/reference unit tests and
doctests with file:line) and against independent brute-force models.  CPU only."""
import itertools

import numpy as np
import pytest

import oracle
from oracle import CMP_LE, CMP_LT, CMP_NONE, OracleCellGrid, canonical_pairs
from zelll_b200 import workload


# ---------------------------------------------------------------------------------------------
# reference known answers
def test_generate_pointcloud_matches_reference_fixture(golden):
    g = golden["generate_pointcloud_3x3x3_unit_origin0"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    assert pts.tolist() == g["points"]


def test_utils_aabb_shape_strides_cell_index(golden):
    g = golden["test_utils"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    assert len(pts) == g["n_points"]
    cg = OracleCellGrid(pts, g["cutoff"])
    info = cg.info()
    assert info["inf"] == g["aabb_inf"]
    assert info["sup"] == g["aabb_sup"]
    assert info["shape"] == g["grid_shape"]
    assert info["strides"] == g["grid_strides"]
    for case in g["cell_index_cases"]:
        assert cg.try_cell_index(case["p"]) == case["cell"]
        assert cg.flat_cell_index(case["p"]) == case["flat"]
        assert cg.flatten_index(case["cell"]) == case["flat"]


def test_neighbor_indices_2d(golden):
    g = golden["test_neighbor_indices_2d"]
    cg = OracleCellGrid(g["points"], g["cutoff"], ndim=2)
    assert cg.neighbor_indices().tolist() == g["neighbor_indices"]


def test_flatindex_keys_of_chessboard(golden):
    g = golden["test_flatindex"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = OracleCellGrid(pts, g["cutoff"])
    want = []
    for x, y, z in itertools.product(range(3), repeat=3):
        if (x + y + z) % 2 == 0:
            want += [cg.flatten_index([x, y, z])] * 2
    assert cg.keys().tolist() == want


def test_cellgrid_iter_counts(golden):
    g = golden["test_cellgrid_iter"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = OracleCellGrid(pts, g["cutoff"])
    keys, begin, length = cg.cells()
    assert len(keys) == g["nonempty_cells"] == cg.info()["n_cells"]
    assert int(length.sum()) == len(pts)


def test_neighborcell_particle_pairs(golden):
    g = golden["test_neighborcell_particle_pairs"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = OracleCellGrid(pts, g["cutoff"])
    assert cg.pair_count(part=1) == g["intra_half"]
    assert cg.pair_count(part=2) == g["inter_half"]
    assert cg.pair_count() == g["intra_half"] + g["inter_half"]
    assert len(cg.pairs()) == g["intra_half"] + g["inter_half"]


def test_half_full_space(golden):
    g = golden["test_half_full_space_particle_pairs"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = OracleCellGrid(pts, g["cutoff"])
    assert cg.pair_count(part=1, full=True) == g["full_over_half_intra"] * cg.pair_count(part=1)
    assert cg.pair_count(part=2, full=True) == g["full_over_half_inter"] * cg.pair_count(part=2)


def test_doctest_flat_cell_index(golden):
    g = golden["doctest_flat_cell_index"]
    cg = OracleCellGrid(g["points"], g["cutoff"])
    cell = cg.try_cell_index(g["p_ok"])
    assert cell is not None
    assert cg.flat_cell_index(g["p_ok"]) == cg.flatten_index(cell)
    assert cg.try_cell_index(g["p_panics"]) is None  # cell_index() would panic


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_doctest_three_points_all_dtypes_and_dims(golden, dtype):
    g = golden["doctest_three_points"]
    cg = OracleCellGrid(g["points"], g["cutoff"], dtype=dtype)
    assert cg.info()["n"] == 3
    _, _, length = cg.cells()
    assert int(length.sum()) == 3
    assert cg.pair_count(nthreads=1) == cg.pair_count(nthreads=4)  # iter vs par_iter
    assert cg.query_neighbors(g["query_point"]) is not None
    cg2 = OracleCellGrid(g["points_2d"], g["cutoff"], dtype=dtype, ndim=2)
    assert cg2.info()["n"] == 3
    # rebuild with reversed data and unchanged cutoff (cellgrid.rs:178-185, :256-263)
    cg.rebuild(g["points"][::-1], None)
    assert cg.info()["cutoff"] == 1.0
    cg.rebuild_mut(g["points"], None)
    assert cg.info()["cutoff"] == 1.0


# ---------------------------------------------------------------------------------------------
# independent models (numpy brute force; scipy KD-tree) -- the parts no reference test pins
def _brute_pairs(pts, c2, le):
    d = pts[:, None, :] - pts[None, :, :]
    sq = d * d
    dsq = (sq[..., 0] + sq[..., 1]) + sq[..., 2] if pts.shape[1] == 3 else sq[..., 0] + sq[..., 1]
    iu, ju = np.triu_indices(len(pts), 1)
    m = dsq[iu, ju] <= c2 if le else dsq[iu, ju] < c2
    return np.stack([iu[m], ju[m]], axis=1).astype(np.uint32), dsq[iu, ju][m]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("le", [False, True])
def test_filtered_pairs_equal_brute_force(dtype, le):
    n = 1500
    pts = workload.generate_points_random(n, dtype=dtype)
    cutoff = dtype(10.0)
    cg = OracleCellGrid(pts, cutoff, dtype=dtype)
    got = cg.pairs_canonical(CMP_LE if le else CMP_LT, cutoff)
    want, dsq = _brute_pairs(pts, dtype(cutoff * cutoff), le)
    assert np.array_equal(got, canonical_pairs(want))
    # LJ energy against a direct numpy evaluation of the same formula
    r = dtype(1.0) / dsq
    t = (r * r) * r
    e = (dtype(4.0) * t) * (t - dtype(1.0))
    e_t, e_d, cnt = cg.lj_energy(CMP_LE if le else CMP_LT, cutoff)
    assert cnt == len(want)
    want_e = float(np.sum(e.astype(np.float64)))
    assert e_d == pytest.approx(want_e, rel=1e-12)
    assert e_t == pytest.approx(want_e, rel=1e-10 if dtype == np.float64 else 1e-4)


def test_candidates_are_all_pairs_in_adjacent_cells():
    n = 800
    pts = workload.generate_points_random(n)
    cg = OracleCellGrid(pts, 10.0)
    info = cg.info()
    inf = np.array(info["inf"])
    cells = np.floor((pts - inf) / 10.0).astype(np.int64)
    iu, ju = np.triu_indices(n, 1)
    adj = np.all(np.abs(cells[iu] - cells[ju]) <= 1, axis=1)
    want = np.stack([iu[adj], ju[adj]], axis=1).astype(np.uint32)
    assert np.array_equal(cg.pairs_canonical(), canonical_pairs(want))


def test_against_scipy_kdtree_medium():
    from scipy.spatial import cKDTree

    n = 20000
    pts = workload.generate_points_random(n)
    cg = OracleCellGrid(pts, 10.0)
    got = cg.pairs_canonical(CMP_LE, 10.0)
    tree = cKDTree(pts)
    want = tree.query_pairs(10.0, output_type="ndarray")
    # KD-tree uses its own arithmetic; compare away from the boundary shell only
    d = np.linalg.norm(pts[want[:, 0]] - pts[want[:, 1]], axis=1)
    safe = np.abs(d - 10.0) > 1e-9
    dg = np.linalg.norm(pts[got[:, 0]] - pts[got[:, 1]], axis=1)
    safe_g = np.abs(dg - 10.0) > 1e-9
    assert np.array_equal(canonical_pairs(want[safe]), got[safe_g])
    # workload constants quoted in SURVEY/BASELINE: ~81.6 candidates and ~16.0 hits per particle
    assert cg.pair_count() / n == pytest.approx(81.6, rel=0.03)
    assert len(got) / n == pytest.approx(16.0, rel=0.03)


def test_parallel_equals_sequential():
    pts = workload.generate_points_random(30000)
    cg = OracleCellGrid(pts, 10.0)
    assert cg.pair_count(CMP_LE, 10.0, nthreads=1) == cg.pair_count(CMP_LE, 10.0, nthreads=4)
    s = cg.lj_energy(CMP_LT, 10.0, nthreads=1)
    p = cg.lj_energy(CMP_LT, 10.0, nthreads=4)
    assert s[2] == p[2]
    assert p[1] == pytest.approx(s[1], rel=1e-10)  # summation order only


def test_rebuild_mut_semantics():
    pts = workload.generate_points_random(5000)
    cg = OracleCellGrid(pts, 10.0)
    base = cg.pairs_canonical(CMP_LT, 10.0)
    # same data again: no key changed (flatindex.rs:140-152)
    assert cg.rebuild_mut(pts, None) is False
    assert np.array_equal(cg.pairs_canonical(CMP_LT, 10.0), base)
    moved = workload.perturb(pts, 0, 1.0)
    changed = cg.rebuild_mut(moved, None)
    assert changed is True
    fresh = OracleCellGrid(moved, 10.0)
    assert np.array_equal(cg.pairs_canonical(CMP_LT, 10.0), fresh.pairs_canonical(CMP_LT, 10.0))
    assert np.array_equal(cg.keys(), fresh.keys())
    # new cutoff through Some(cutoff)
    cg.rebuild_mut(moved, 7.5)
    assert cg.info()["cutoff"] == 7.5
    assert np.array_equal(cg.pairs_canonical(CMP_LE, 7.5), OracleCellGrid(moved, 7.5).pairs_canonical(CMP_LE, 7.5))


def test_empty_and_single():
    cg = OracleCellGrid(np.zeros((0, 3)), 1.0)
    info = cg.info()
    assert info["n"] == 0 and info["n_cells"] == 0
    assert info["inf"] == [0.0, 0.0, 0.0] and info["shape"] == [1, 1, 1] and info["strides"] == [1, 5, 25]
    assert cg.pair_count() == 0
    cg = OracleCellGrid(np.array([[1.0, 2.0, 3.0]]), 1.0)
    assert cg.info()["n_cells"] == 1 and cg.pair_count() == 0
    assert cg.info()["inf"] == [1.0, 2.0, 3.0] == cg.info()["sup"]


def test_query_neighbors_full_shell():
    pts = workload.generate_points_random(4000)
    cg = OracleCellGrid(pts, 10.0)
    info = cg.info()
    inf = np.array(info["inf"])
    cells = np.floor((pts - inf) / 10.0).astype(np.int64)
    rng = np.random.default_rng(1)
    for q in pts[rng.integers(0, len(pts), 20)] + rng.normal(0, 3.0, (20, 3)):
        got = cg.query_neighbors(q)
        qc = np.floor((q - inf) / 10.0).astype(np.int64)
        inside = np.all((qc >= -1) & (qc <= np.array(info["shape"])))
        if not inside:
            assert got is None
            continue
        want = np.nonzero(np.all(np.abs(cells - qc) <= 1, axis=1))[0]
        assert sorted(got.tolist()) == sorted(want.tolist())
        near = cg.query_neighbors(q, CMP_LE, 10.0)
        d = pts[want] - q
        dsq = (d[:, 0] ** 2 + d[:, 1] ** 2) + d[:, 2] ** 2
        assert sorted(near.tolist()) == sorted(want[dsq <= 100.0].tolist())
    assert cg.query_neighbors(inf - 25.0) is None


def sdf_from_neighbors(x, nbr_xyz, radius, cutoff):
    """SmoothDistanceField::sdf (surface-sampling/src/sdf/numdual.rs:11-58) over the points that
    query_neighbors(x) yielded: filter dist <= cutoff, dist == 0 handled as (1, r, 1)."""
    d = np.sqrt(((nbr_xyz - np.asarray(x)) ** 2).sum(axis=1))
    d = d[d <= cutoff]
    scaled = np.where(d != 0.0, np.exp(-d / radius), 1.0).sum()
    radii = np.where(d != 0.0, np.exp(-d) * radius, radius).sum()
    total = np.where(d != 0.0, np.exp(-d), 1.0).sum()
    return -(radii / total) * np.log(scaled)


def test_reference_sdf_golden(golden):
    """The reference's only floating-point known answer behind query_neighbors (numdual.rs:107-192)."""
    g = golden["test_sdf_autodiff"]
    pts = np.array(g["points"])
    cg = OracleCellGrid(pts, g["cutoff"])
    for x, want in zip(pts, g["reference_values"]):
        lab = cg.query_neighbors(x)
        assert lab is not None
        got = sdf_from_neighbors(x, pts[lab.astype(np.int64)], g["radius"], g["cutoff"])
        assert abs(got - want) <= 4e-16 * abs(want) * len(lab)  # summation order of the fold only
        near = cg.query_neighbors(x, CMP_LE, g["cutoff"])
        d = np.sqrt(((pts - x) ** 2).sum(axis=1))
        assert sorted(near.tolist()) == np.nonzero(d <= g["cutoff"])[0].tolist()


def test_cross_tool_inputs():
    """the LAMMPS deck and CellListMap settings of the reference's comparison
    (more_benches/in.zelllbench.txt:30-36, more_benches/celllistmap.jl:18-43) next to the data writer."""
    deck = workload.lammps_input(10.0, repeat=5, data_file="atoms.txt")
    for needle in ("units lj", "atom_style atomic", "boundary f f f", "read_data ${data} add merge",
                   "pair_style lj/cut ${cutoff}", "pair_coeff 1 1 1.0 1.0", "neighbor 0.0 bin",
                   "neigh_modify delay 0 every 1 check no one ${max_neighbors}", "run ${repeat}",
                   "variable cutoff index 10.0", "variable repeat index 5", "variable data file atoms.txt"):
        assert needle in deck, needle
    cfg = workload.celllistmap_settings(100_000)
    assert cfg["sides"] == [30.0, 30.0, 100_000 / 9.0] and cfg["cutoff"] == 10.0 and cfg["parallel"] is False
    assert workload.celllistmap_settings(100)["sides"][2] == 30.0  # max(c, 3 cutoff), celllistmap.jl:31
    script = workload.celllistmap_script(100_000)
    assert "map_pairwise!" in script and "parallel=false" in script and "/ n" in script
    # the data file those tools read: 10 header lines, then `id type x y z` rows (examples/lammps_data.rs:56-80)
    pts = workload.generate_points_random(5)
    text = workload.lammps_data(pts).splitlines()
    assert text[2] == "5 atoms" and len(text) == 10 + 5 + 1
    assert [float(v) for v in text[10].split()[2:]] == pts[0].tolist()
