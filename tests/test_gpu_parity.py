"""GPU parity tests: the CUDA engine, called through the C ABI (zelll_b200.CellGrid is a ctypes
shim over libzelll_b200.so), against the CPU oracle on the same seeded inputs, against the
reference's known answers, and -- at BASELINE.json's full size -- through size-independent
properties.  Bar: bit-exact for keys / cells / pair sets / counts; 1e-10 (f64) and 1e-5 (f32)
relative for the Lennard-Jones energy."""
import itertools
import os
import pickle

import numpy as np
import pytest

import oracle
from oracle import CMP_LE, CMP_LT, CMP_NONE, OracleCellGrid, canonical_pairs
from zelll_b200 import workload

pytestmark = pytest.mark.gpu

F64_RTOL = 1e-10  # north_star tolerance, f64
F32_RTOL = 1e-5   # north_star tolerance, f32
OCMP = {"none": CMP_NONE, "lt": CMP_LT, "le": CMP_LE}


@pytest.fixture(scope="module")
def zb():
    import zelll_b200

    return zelll_b200


def _cloud(kind: str, n: int, dtype, ndim: int = 3, seed: int = 1):
    rng = np.random.default_rng(seed)
    if kind == "lj":  # the benchmark box (benches/lj.rs:59-66)
        pts = workload.generate_points_random(n, dtype=dtype)
        return pts[:, :ndim].copy(), 10.0
    if kind == "cube":
        side = max(1.0, (n / 4.0) ** (1.0 / ndim))
        return (rng.random((n, ndim)) * side).astype(dtype), 1.0
    if kind == "clusters":  # sparse: a few dense blobs far apart
        centres = rng.random((6, ndim)) * 60.0
        pts = centres[rng.integers(0, 6, n)] + rng.normal(0, 0.7, (n, ndim))
        return pts.astype(dtype), 1.5
    if kind == "plane":  # wide and flat: 2 layers in z
        side = max(2.0, (n / 8.0) ** 0.5)
        pts = rng.random((n, ndim)) * ([side, side, 2.0][:ndim])
        return pts.astype(dtype), 1.0
    if kind == "dense":  # few cells, many particles each (exceeds the shared-memory stage)
        return (rng.random((n, ndim)) * 2.0).astype(dtype), 1.0
    raise ValueError(kind)


def _check_against_oracle(zb, pts, cutoff, dtype, ndim, cmps=("none", "lt", "le"), energy=True):
    cg = zb.CellGrid(pts, cutoff, dtype=dtype, ndim=ndim)
    og = OracleCellGrid(pts, cutoff, dtype=dtype, ndim=ndim)
    info, oinfo = cg.info(), og.info()
    assert info.origin().tolist() == oinfo["inf"]
    assert info.bounding_box()[1].tolist() == oinfo["sup"]
    assert info.shape().tolist() == oinfo["shape"]
    assert info.strides().tolist() == oinfo["strides"]
    assert info.n == oinfo["n"] == len(pts)
    assert info.n_cells == oinfo["n_cells"]
    # FlatIndex.index
    assert np.array_equal(cg.keys(), og.keys())
    assert cg.neighbor_indices().tolist() == og.neighbor_indices().tolist()
    # cells: same key set, same sizes, same label multiset per cell (order inside a cell and of
    # the cells in the buffer is unspecified upstream: hash-map order, iters.rs:262)
    keys, begin, count = cg.cells()
    okeys, obegin, olen = og.cells()
    assert np.all(np.diff(keys) > 0)
    order = np.argsort(okeys)
    assert np.array_equal(keys, okeys[order])
    assert np.array_equal(count.astype(np.uint64), olen[order])
    labels, xyz = cg.cell_storage()
    olabels, oxyz = og.cell_storage()
    assert np.array_equal(xyz, np.asarray(pts, dtype=dtype)[labels])  # records carry their own coordinates
    for k, (b, c) in enumerate(zip(begin, count)):
        ob, oc = int(obegin[order[k]]), int(olen[order[k]])
        assert sorted(labels[b:b + c].tolist()) == sorted(olabels[ob:ob + oc].tolist())
    # pair sets, bit-exact in canonical form
    for cmp in cmps:
        want = og.pairs_canonical(OCMP[cmp], cutoff)
        got = canonical_pairs(cg.particle_pairs(cutoff, cmp))
        assert got.shape == want.shape, (cmp, got.shape, want.shape)
        assert np.array_equal(got, want), cmp
        assert cg.pair_count(cutoff, cmp) == len(want)
    if energy:
        for cmp in ("lt", "le"):
            e_t, e_64, npairs = og.lj_energy(OCMP[cmp], cutoff)
            e, m = cg.lj_energy(cutoff, cmp, return_pairs=True)
            assert m == npairs
            rtol = F64_RTOL if np.dtype(dtype) == np.float64 else F32_RTOL
            if np.isfinite(e_64):
                assert abs(e - e_64) <= rtol * max(abs(e_64), 1e-300), (e, e_64)
    return cg, og


# ---------------------------------------------------------------------------------------------
# the reference's own known answers, on the GPU path
def test_reference_utils_golden(zb, golden):
    g = golden["test_utils"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = zb.CellGrid(pts, g["cutoff"])
    info = cg.info()
    assert info.origin().tolist() == g["aabb_inf"]
    assert info.bounding_box()[1].tolist() == g["aabb_sup"]
    assert cg.aabb() == (g["aabb_inf"], g["aabb_sup"])
    assert info.shape().tolist() == g["grid_shape"]
    assert info.strides().tolist() == g["grid_strides"]
    for case in g["cell_index_cases"]:
        assert info.try_cell_index(case["p"]) == case["cell"]
        assert info.flat_cell_index(case["p"]) == case["flat"]
        assert info.flatten_index(case["cell"]) == case["flat"]
    # the device computes the same keys for these points (incl. the 1.9999999999999998 floor case)
    probe = np.array([c["p"] for c in g["cell_index_cases"]])
    cg2 = zb.CellGrid(np.vstack([pts, probe]), g["cutoff"])
    assert cg2.keys()[len(pts):].tolist() == [c["flat"] for c in g["cell_index_cases"]]


def test_reference_neighbor_indices_2d_golden(zb, golden):
    g = golden["test_neighbor_indices_2d"]
    cg = zb.CellGrid(np.array(g["points"]), g["cutoff"], ndim=2)
    assert cg.neighbor_indices().tolist() == g["neighbor_indices"]


def test_reference_flatindex_golden(zb, golden):
    g = golden["test_flatindex"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = zb.CellGrid(pts, g["cutoff"])
    info = cg.info()
    want = []
    for x, y, z in itertools.product(range(3), repeat=3):
        if (x + y + z) % 2 == 0:
            want += [info.flatten_index([x, y, z])] * 2
    assert cg.keys().tolist() == want


def test_reference_iter_counts_golden(zb, golden):
    g = golden["test_cellgrid_iter"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = zb.CellGrid(pts, g["cutoff"])
    keys, begin, count = cg.cells()
    assert len(keys) == g["nonempty_cells"] == cg.info().n_cells
    assert int(count.sum()) == len(pts)


def test_reference_pair_counts_golden(zb, golden):
    g = golden["test_neighborcell_particle_pairs"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = zb.CellGrid(pts, g["cutoff"])
    assert cg.pair_count() == g["intra_half"] + g["inter_half"]
    assert len(cg.particle_pairs()) == g["intra_half"] + g["inter_half"]
    assert sum(1 for _ in cg) == g["intra_half"] + g["inter_half"]


def test_reference_doctest_three_points(zb, golden):
    g = golden["doctest_three_points"]
    for dtype in (np.float32, np.float64):
        _check_against_oracle(zb, np.array(g["points"], dtype=dtype), g["cutoff"], dtype, 3)
        _check_against_oracle(zb, np.array(g["points_2d"], dtype=dtype), g["cutoff"], dtype, 2)


# ---------------------------------------------------------------------------------------------
# seeded clouds vs the oracle
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,n", [("lj", 1000), ("lj", 20000), ("cube", 5000), ("clusters", 3000),
                                    ("plane", 30000), ("dense", 6000)])
def test_cloud_parity_3d(zb, kind, n, dtype):
    pts, cutoff = _cloud(kind, n, dtype)
    _check_against_oracle(zb, pts, cutoff, dtype, 3, energy=(kind != "dense" or dtype == np.float64))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,n", [("cube", 4000), ("clusters", 2000), ("dense", 3000)])
def test_cloud_parity_2d(zb, kind, n, dtype):
    pts, cutoff = _cloud(kind, n, dtype, ndim=2)
    _check_against_oracle(zb, pts, cutoff, dtype, 2)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 257])
def test_tiny_and_ragged_inputs(zb, n):
    rng = np.random.default_rng(n)
    pts = rng.random((n, 3)) * 3.0
    _check_against_oracle(zb, pts, 1.0, np.float64, 3)


def test_degenerate_clouds(zb):
    same = np.zeros((40, 3)) + 0.25  # all particles coincide: dsq = 0, energy = inf
    cg, og = _check_against_oracle(zb, same, 1.0, np.float64, 3, energy=False)
    assert cg.pair_count(1.0, "le") == 40 * 39 // 2
    assert cg.pair_count(1.0, "lt") == 40 * 39 // 2
    line = np.zeros((500, 3))
    line[:, 2] = np.linspace(0.0, 400.0, 500)  # shape [1, 1, 401]
    _check_against_oracle(zb, line, 1.0, np.float64, 3)
    # cutoff larger than the box: a single cell
    _check_against_oracle(zb, np.random.default_rng(3).random((300, 3)), 5.0, np.float64, 3)
    # negative coordinates / off-origin box
    _check_against_oracle(zb, np.random.default_rng(4).random((2000, 3)) * 9.0 - 100.0, 1.0, np.float32, 3)


def test_filter_radius_differs_from_grid_cutoff(zb):
    pts, cutoff = _cloud("cube", 4000, np.float64)
    cg = zb.CellGrid(pts, cutoff)
    og = OracleCellGrid(pts, cutoff)
    for r in (0.3, 0.77, 1.0):
        assert np.array_equal(canonical_pairs(cg.particle_pairs(r, "le")), og.pairs_canonical(CMP_LE, r))


def test_boundary_distances_lt_vs_le(zb):
    # pairs exactly AT the cutoff: `<` drops them, `<=` keeps them (benches/lj.rs:85 vs cellgrid.rs:86)
    pts = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.6, 0.8, 0.0], [3.0, 4.0, 0.0],
                    [3.0, 4.0, 1.0]])
    for dtype in (np.float32, np.float64):
        cg, og = _check_against_oracle(zb, pts.astype(dtype), 1.0, dtype, 3)
        assert cg.pair_count(1.0, "le") > cg.pair_count(1.0, "lt")


# ---------------------------------------------------------------------------------------------
# rebuild_mut: buffer reuse, cutoff = None, key-change flag
def test_rebuild_mut_sequence(zb):
    pts, cutoff = _cloud("lj", 5000, np.float64)
    cg = zb.CellGrid(pts, cutoff)
    og = OracleCellGrid(pts, cutoff)
    cg.track_key_changes(True)
    cg.rebuild(pts)  # primes the key history
    for step in range(4):
        amp = 0.0 if step == 1 else 0.3
        pts = workload.perturb(pts, step, amp)
        cg.rebuild_mut(pts, None)  # Option::None keeps the cutoff (flatindex.rs:118)
        changed = og.rebuild_mut(pts, None)
        assert cg.cutoff() == cutoff
        assert cg.info().keys_changed == changed
        assert np.array_equal(cg.keys(), og.keys())
        assert np.array_equal(canonical_pairs(cg.particle_pairs(cutoff, "lt")), og.pairs_canonical(CMP_LT, cutoff))
    # a new cutoff and a different particle count through the same handle
    pts2, _ = _cloud("cube", 777, np.float64)
    cg.rebuild(pts2, 0.5)
    og2 = OracleCellGrid(pts2, 0.5)
    assert cg.cutoff() == 0.5
    assert np.array_equal(canonical_pairs(cg.particle_pairs(0.5, "le")), og2.pairs_canonical(CMP_LE, 0.5))


def test_device_resident_input(zb):
    import torch

    pts, cutoff = _cloud("lj", 4000, np.float64)
    t = torch.from_numpy(pts).cuda()
    cg = zb.CellGrid(t, cutoff)
    og = OracleCellGrid(pts, cutoff)
    assert np.array_equal(cg.keys(), og.keys())
    assert np.array_equal(canonical_pairs(cg.particle_pairs(cutoff, "lt")), og.pairs_canonical(CMP_LT, cutoff))


# ---------------------------------------------------------------------------------------------
# Python-binding behaviour (python/src/lib.rs)
def test_python_binding_surface(zb):
    pts = [[0.0, 0.0, 0.0], [1.0, 2.0, 0.0], "bogus", [0.0, 0.1, 0.2], [1.0, 2.0]]
    cg = zb.CellGrid(pts, 1.0)  # unconvertible items are skipped but keep their index (lib.rs:47-57)
    seen = sorted((min(i, j), max(i, j)) for (i, p), (j, q) in cg)
    assert seen == [(0, 3)]
    for (i, p), (j, q) in cg:
        assert p == pts[i] and q == pts[j]
    assert cg.cutoff() == 1.0
    inf, sup = cg.aabb()
    assert inf == [0.0, 0.0, 0.0] and sup == [1.0, 2.0, 0.2]
    near = cg.neighbors([0.5, 1.0, 0.1])
    assert near is not None
    far = cg.query_neighbors([50.0, 50.0, 50.0])
    assert far is None
    empty = zb.CellGrid()
    assert list(empty) == [] and empty.cutoff() == 1.0
    clone = pickle.loads(pickle.dumps(cg))
    assert sorted((min(i, j), max(i, j)) for (i, p), (j, q) in clone) == seen
    it = iter(cg)
    with pytest.raises(RuntimeError):
        cg.rebuild(pts, 1.0)
    del it
    cg.rebuild(pts, 1.0)


def test_query_neighbors_vs_oracle(zb):
    pts, cutoff = _cloud("cube", 3000, np.float64)
    cg = zb.CellGrid(pts, cutoff)
    og = OracleCellGrid(pts, cutoff)
    rng = np.random.default_rng(5)
    lo, hi = pts.min(0) - 2.5 * cutoff, pts.max(0) + 2.5 * cutoff
    queries = lo + rng.random((300, 3)) * (hi - lo)
    queries[:20] = pts[:20]
    for cmp in ("none", "le"):
        offsets, valid, labels = cg.query_neighbors_batch(queries, cutoff, cmp)
        for q in range(len(queries)):
            want = og.query_neighbors(queries[q], OCMP[cmp], cutoff)
            if want is None:
                assert not valid[q] and offsets[q + 1] == offsets[q]
            else:
                assert valid[q]
                got = labels[int(offsets[q]):int(offsets[q + 1])]
                assert sorted(got.tolist()) == sorted(want.tolist())


# ---------------------------------------------------------------------------------------------
# slab-sharded grids, ranks emulated one after another on one GPU
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_union_equals_single_grid(zb, world):
    from zelll_b200 import sharded

    pts, cutoff = _cloud("lj", 20000, np.float64)
    single = zb.CellGrid(pts, cutoff)
    want = canonical_pairs(single.particle_pairs(cutoff, "lt"))
    e_want, m_want = single.lj_energy(cutoff, "lt", return_pairs=True)
    info = single.info()
    inf, sup = info.bounding_box()
    nz = int(info.shape()[2])
    labels = np.arange(len(pts), dtype=np.uint32)
    layer = np.floor((pts[:, 2] - inf[2]) / cutoff).astype(np.int64)
    got, e_sum, m_sum = [], 0.0, 0
    for r in range(world):
        zb_, ze_ = sharded.slab_bounds(nz, world, r)
        sel = (layer >= max(zb_ - 1, 0)) & (layer < ze_)
        g = sharded.ShardedCellGrid(dtype=np.float64)
        g.rebuild_local(pts[sel], labels[sel], cutoff, inf, sup, zb_, ze_)
        got.append(g.particle_pairs(cutoff, "lt"))
        e, m = g.lj_energy(cutoff, "lt", return_pairs=True)
        e_sum += e
        m_sum += m
    got = canonical_pairs(np.concatenate(got))
    assert np.array_equal(got, want)
    assert m_sum == m_want
    assert abs(e_sum - e_want) <= F64_RTOL * abs(e_want)


# ---------------------------------------------------------------------------------------------
# BASELINE.json full size (n = 10^7): size-independent properties
@pytest.mark.parametrize("dtype", [np.float64])
def test_full_size_properties(zb, dtype):
    n = 10_000_000
    pts = workload.generate_points_random(n, dtype=dtype)
    cg = zb.CellGrid(pts, 10.0, dtype=dtype)
    info = cg.info()
    assert info.shape().tolist() == [3, 3, (n + 89) // 90] or info.shape()[2] in ((n // 90), (n // 90) + 1, (n // 90) + 2)
    keys, begin, count = cg.cells()
    assert int(count.sum()) == n and np.all(np.diff(keys) > 0)
    assert np.array_equal(begin, np.concatenate([[0], np.cumsum(count)[:-1]]).astype(np.uint32))
    labels, xyz = cg.cell_storage()
    assert np.array_equal(np.sort(labels), np.arange(n, dtype=np.uint32))  # a permutation
    assert np.array_equal(xyz, pts[labels])
    # every record sits in the cell its key names
    k = cg.keys()
    cell_of_slot = np.repeat(keys, count)
    assert np.array_equal(k[labels], cell_of_slot)
    # count == fused-consumer count == emitted rows; candidates >= filtered
    c_le = cg.pair_count(10.0, "le")
    c_lt = cg.pair_count(10.0, "lt")
    e1, m1 = cg.lj_energy(10.0, "lt", return_pairs=True)
    assert m1 == c_lt <= c_le <= cg.pair_count()
    assert 15.5 * n < c_lt < 16.5 * n  # 16.0 in-cutoff pairs per particle
    # a chunk of the oracle: the first 200k particles in z order form a closed sub-box
    order = np.argsort(pts[:, 2], kind="stable")
    sub = pts[order[:200_000]]
    og = OracleCellGrid(sub, 10.0, dtype=dtype)
    cs = zb.CellGrid(sub, 10.0, dtype=dtype)
    e_o = og.lj_energy(CMP_LT, 10.0)
    e_s, m_s = cs.lj_energy(10.0, "lt", return_pairs=True)
    assert m_s == e_o[2]
    assert abs(e_s - e_o[1]) <= F64_RTOL * abs(e_o[1])
    # permutation invariance: same pair count, energy equal to reduction-order noise
    perm = np.random.default_rng(0).permutation(n)
    cg.rebuild(pts[perm])
    e2, m2 = cg.lj_energy(10.0, "lt", return_pairs=True)
    assert m2 == m1
    assert abs(e2 - e1) <= F64_RTOL * abs(e1)
    # idempotence of rebuild
    cg.rebuild(pts[perm])
    e3, m3 = cg.lj_energy(10.0, "lt", return_pairs=True)
    assert m3 == m2 and abs(e3 - e2) <= 1e-13 * abs(e2)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[2] at FULL size against the oracle itself (all host threads): pair counts
# for `<` and `<=` bit-exact, LJ energy within the north-star tolerance
def test_full_size_counts_and_energy_vs_oracle(zb):
    import os

    n = 10_000_000
    pts = workload.generate_points_random(n)
    threads = max(1, min(os.cpu_count() or 1, oracle.max_threads()))
    og = OracleCellGrid(pts, 10.0)
    cg = zb.CellGrid(pts, 10.0)
    assert cg.info().shape().tolist() == og.info()["shape"]
    assert cg.info().n_cells == og.info()["n_cells"]
    c_lt = og.pair_count(CMP_LT, 10.0, nthreads=threads)
    c_le = og.pair_count(CMP_LE, 10.0, nthreads=threads)
    assert cg.pair_count(10.0, "lt") == c_lt
    assert cg.pair_count(10.0, "le") == c_le
    assert cg.pair_count() == og.pair_count(nthreads=threads)
    _, e64, m = og.lj_energy(CMP_LT, 10.0, nthreads=threads)
    e, m_gpu = cg.lj_energy(10.0, "lt", return_pairs=True)
    assert m_gpu == m == c_lt
    assert abs(e - e64) <= F64_RTOL * abs(e64), (e, e64)


# BASELINE.json configs[0] exactly: benches/lj.rs, n = 10^5 f64, CellGrid::new + particle_pairs energy
def test_config0_lj_1e5_vs_oracle(zb):
    pts = workload.generate_points_random(100_000)
    og = OracleCellGrid(pts, 10.0)
    cg = zb.CellGrid(pts, 10.0)
    for cmp in ("lt", "le"):
        _, e64, m = og.lj_energy(OCMP[cmp], 10.0)
        e, m_gpu = cg.lj_energy(10.0, cmp, return_pairs=True)
        assert m_gpu == m
        assert abs(e - e64) <= F64_RTOL * abs(e64), (cmp, e, e64)
    assert np.array_equal(canonical_pairs(cg.particle_pairs(10.0, "lt")), og.pairs_canonical(CMP_LT, 10.0))


# the reference's one floating-point known answer behind query_neighbors: the smooth distance field
# of surface-sampling/src/sdf/numdual.rs:107-192, fed by the device's batched query
def test_reference_sdf_golden(zb, golden):
    from test_oracle import sdf_from_neighbors

    g = golden["test_sdf_autodiff"]
    pts = np.array(g["points"])
    cg = zb.CellGrid(pts, g["cutoff"])
    offsets, valid, labels = cg.query_neighbors_batch(pts, g["cutoff"], "none")
    off_le, valid_le, labels_le = cg.query_neighbors_batch(pts, g["cutoff"], "le")
    assert valid.all() and valid_le.all()
    for q, want in enumerate(g["reference_values"]):
        lab = labels[int(offsets[q]):int(offsets[q + 1])].astype(np.int64)
        got = sdf_from_neighbors(pts[q], pts[lab], g["radius"], g["cutoff"])
        assert abs(got - want) <= 4e-16 * abs(want) * len(lab), (q, got, want)
        # the device-side `<=` filter keeps exactly the atoms the reference's `dist <= cutoff` keeps
        lab_le = labels_le[int(off_le[q]):int(off_le[q + 1])].astype(np.int64)
        got_le = sdf_from_neighbors(pts[q], pts[lab_le], g["radius"], g["cutoff"])
        assert abs(got_le - want) <= 4e-16 * abs(want) * len(lab), (q, got_le, want)


# ---------------------------------------------------------------------------------------------
# real multi-GPU run (NCCL) when the box has more than one GPU; the host logic is covered on CPU by
# tests/test_sharded_cpu.py (gloo) and the per-rank engine by test_sharded_union_equals_single_grid
@pytest.mark.parametrize("transport", ["peer-memory+speculation", "nccl+wait"])
def test_distributed_nccl_matches_single_gpu(transport):
    """scripts/dist_check.py under torchrun (needs >= 2 GPUs): general, slab-local and native slab paths, then
    repeated and moved frames through the native path -- once with the default transport (exchanges over
    mapped peer memory, speculative box) and once with NCCL collectives and a wait for the box."""
    import os
    import subprocess
    import sys

    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", "scripts/dist_check.py", "200000"]
    env = dict(os.environ)
    if transport == "nccl+wait":
        env.update(ZB_P2P="0", ZB_SLAB_SPEC="0")
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MISMATCH" not in out.stdout and out.stdout.count("-> OK") >= 10


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs as parity cases
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [10_000, 100_000])
def test_config_build_sweep(zb, n, dtype):
    """configs[1]: CellGrid construction (benches/cellgrid.rs) f32/f64 + `<=` pair count."""
    pts = workload.generate_points_random(n, dtype=dtype)
    cg = zb.CellGrid(pts, 10.0, dtype=dtype)
    og = OracleCellGrid(pts, 10.0, dtype=dtype)
    assert np.array_equal(cg.keys(), og.keys())
    assert cg.info().n_cells == og.info()["n_cells"]
    assert cg.pair_count(10.0, "le") == og.pair_count(CMP_LE, 10.0)
    assert cg.pair_count() == og.pair_count()


def test_config_presorted_perturbed_rebuild_loop(zb):
    """configs[3]: z-presorted cloud (examples/cachemisses.rs:57-59), repeated rebuild_mut(None) over
    perturbed steps; every step must match the oracle's rebuild_mut of the same positions."""
    pts = workload.presort_by_z(workload.generate_points_random(30_000))
    cg = zb.CellGrid(pts, 10.0)
    og = OracleCellGrid(pts, 10.0)
    cg.track_key_changes(True)
    cg.rebuild(pts)
    for step in range(6):
        pts = workload.perturb(pts, step, 0.1 * 10.0 if step % 3 else 0.0)
        cg.rebuild_mut(pts, None)
        changed = og.rebuild_mut(pts, None)
        assert cg.info().keys_changed == changed
        assert cg.pair_count(10.0, "lt") == og.pair_count(CMP_LT, 10.0)
        e, m = cg.lj_energy(10.0, "lt", return_pairs=True)
        _, e64, mo = og.lj_energy(CMP_LT, 10.0)
        assert m == mo and abs(e - e64) <= F64_RTOL * abs(e64)


def test_presorted_full_size_step_properties(zb):
    """configs[3] at n = 10^7: the presorted cloud and its shuffled copy give the same counts and
    (to reduction-order noise) the same energy; perturbing by 0 leaves every key unchanged."""
    n = 10_000_000
    pts = workload.generate_points_random(n)
    srt = workload.presort_by_z(pts)
    cg = zb.CellGrid(srt, 10.0)
    e1, m1 = cg.lj_energy(10.0, "lt", return_pairs=True)
    cg.track_key_changes(True)
    cg.rebuild(srt)
    cg.rebuild_mut(srt, None)
    assert cg.info().keys_changed is False
    cg.rebuild_mut(workload.perturb(srt, 0, 1.0), None)
    assert cg.info().keys_changed is True
    cg2 = zb.CellGrid(pts, 10.0)
    e2, m2 = cg2.lj_energy(10.0, "lt", return_pairs=True)
    assert m1 == m2 and abs(e1 - e2) <= F64_RTOL * abs(e2)


def test_error_statuses(zb):
    """The reference panics (cellgrid.rs:227-229, util.rs:229-232); the C ABI returns statuses."""
    from zelll_b200 import _ffi

    cg = zb.CellGrid(np.random.default_rng(0).random((100, 3)), 0.5)
    with pytest.raises(zb.ZelllB200Error) as e:
        cg.rebuild(np.random.default_rng(0).random((100, 3)), -1.0)
    assert e.value.status == _ffi.ERR_BAD_ARG
    with pytest.raises(zb.ZelllB200Error) as e:
        cg.rebuild(np.array([[0.0, 0.0, 0.0], [1e9, 1e9, 1e9]]), 1e-3)   # 10^36 cells
    assert e.value.status == _ffi.ERR_GRID_TOO_LARGE
    with pytest.raises(zb.ZelllB200Error) as e:
        cg.pair_count(1.0, "le")                                          # last rebuild failed
    assert e.value.status == _ffi.ERR_NOT_BUILT
    bad = np.random.default_rng(0).random((10, 3))
    bad[3, 1] = np.inf   # (NaN follows Rust's `as i32` -> 0 and is accepted, like upstream)
    with pytest.raises(zb.ZelllB200Error):
        cg.rebuild(bad, 0.5)
    cg.rebuild(np.random.default_rng(1).random((50, 3)), 0.5)            # the handle recovers
    assert cg.info().n == 50


def test_very_sparse_box(zb):
    """zelll's home turf: few particles in a huge box (10^8 cells here, almost all empty).  The build
    switches to compact sorted cells (O(n) memory, sparse_kernels.cuh) and still gives the reference's
    cells, pair set and energy."""
    rng = np.random.default_rng(7)
    blobs = np.array([[0.0, 0.0, 0.0], [450.0, 20.0, 460.0], [30.0, 440.0, 10.0]])
    pts = blobs[rng.integers(0, 3, 6000)] + rng.normal(0.0, 1.5, (6000, 3))
    cg, og = _check_against_oracle(zb, pts, 1.0, np.float64, 3)
    assert int(np.prod(cg.info().shape().astype(np.int64))) > 5e7


def test_box_beyond_2_31_cells(zb):
    """A box the dense table cannot index (round 1: ZB_ERR_GRID_TOO_LARGE): 4000 particles in clusters
    spread over ~10^11 cells.  The reference handles it trivially (hash map of non-empty cells)."""
    rng = np.random.default_rng(11)
    blobs = rng.random((5, 3)) * 5000.0
    blobs[0] = 0.0
    blobs[1] = 5000.0
    pts = blobs[rng.integers(0, 5, 4000)] + rng.normal(0.0, 1.2, (4000, 3))
    cg = zb.CellGrid(pts, 1.0)
    og = OracleCellGrid(pts, 1.0)
    info, oinfo = cg.info(), og.info()
    shape = info.shape().astype(np.int64)
    assert int(np.prod(shape)) > 2**31
    assert info.shape().tolist() == oinfo["shape"] and info.strides().tolist() == oinfo["strides"]
    assert info.n_cells == oinfo["n_cells"]
    assert np.array_equal(cg.keys(), og.keys())   # the reference's i32 flat keys WRAP in a box this large
    # cells: the same (key, label set) groups (the wrapped keys are neither sorted nor unique)
    keys, begin, count = cg.cells()
    labels, xyz = cg.cell_storage()
    okeys, obegin, olen = og.cells()
    olabels, _ = og.cell_storage()
    got_cells = sorted((int(k), tuple(sorted(labels[b:b + c].tolist()))) for k, b, c in zip(keys, begin, count))
    want_cells = sorted((int(k), tuple(sorted(olabels[int(b):int(b + c)].tolist()))) for k, b, c in zip(okeys, obegin, olen))
    assert got_cells == want_cells
    assert np.array_equal(xyz, pts[labels])
    for cmp in ("none", "lt", "le"):
        want = og.pairs_canonical(OCMP[cmp], 1.0)
        assert np.array_equal(canonical_pairs(cg.particle_pairs(1.0, cmp)), want), cmp
        assert cg.pair_count(1.0, cmp) == len(want)
    _, e64, m = og.lj_energy(CMP_LT, 1.0)
    e, m_gpu = cg.lj_energy(1.0, "lt", return_pairs=True)
    assert m_gpu == m and abs(e - e64) <= F64_RTOL * abs(e64)
    # point queries go through the compact cells too
    lo, hi = pts.min(0), pts.max(0)
    queries = np.vstack([pts[:50], blobs + 0.3, lo + rng.random((50, 3)) * (hi - lo)])
    for cmp in ("none", "le"):
        offsets, valid, labels = cg.query_neighbors_batch(queries, 1.0, cmp)
        for q in range(len(queries)):
            want = og.query_neighbors(queries[q], OCMP[cmp], 1.0)
            if want is None:
                assert not valid[q]
            else:
                assert valid[q]
                assert sorted(labels[int(offsets[q]):int(offsets[q + 1])].tolist()) == sorted(want.tolist())


def test_compact_cell_build_forced():
    """ZB_SPARSE=2 forces the compact-cell build on ordinary clouds (benchmark box, cube, clusters, dense
    cells, 2-D, tiny inputs, f32): cells, pair sets, counts and energies must be the oracle's, and the
    records of a cell come out in input order (the radix sort is stable), as CellStorage::push leaves them."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import numpy as np, sys
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import oracle, zelll_b200
import test_gpu_parity as t
for dtype in (np.float64, np.float32):
    for kind, n, nd in (("lj", 20000, 3), ("cube", 5000, 3), ("clusters", 3000, 3), ("dense", 3000, 3), ("plane", 4000, 3),
                        ("cube", 3000, 2), ("dense", 2000, 2)):
        pts, c = t._cloud(kind, n, dtype, nd)
        cg, og = t._check_against_oracle(zelll_b200, pts, c, dtype, nd)
        labels, _ = cg.cell_storage()
        keys, begin, count = cg.cells()
        for b, m in zip(begin[:300], count[:300]):
            assert np.all(np.diff(labels[b:b + m].astype(np.int64)) > 0)   # input order inside a cell
for n in (0, 1, 2, 3, 31, 32, 33, 257):
    pts = np.random.default_rng(n).random((n, 3)) * 3.0
    t._check_against_oracle(zelll_b200, pts, 1.0, np.float64, 3)
print("sparse ok")
""" % (root, os.path.join(root, "tests"))
    env = dict(os.environ, ZB_SPARSE="2")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "sparse ok" in out.stdout, out.stdout[-1500:] + out.stderr[-2500:]


def test_gridcell_views_reference_counts(zb, golden):
    """GridCell-level API (iters.rs:121-290) on the reference's own fixtures: 14 non-empty cells
    (iters.rs:299-308), sum of cell sizes = n (:311-331), 4 intra + 24 inter half-space pairs on the
    2x2x2 chessboard (:334-356), Full = 2 x Half (:359-387)."""
    g = golden["test_cellgrid_iter"]
    pts = oracle.generate_pointcloud(g["shape"], g["cutoff"], g["origin"])
    cg = zb.CellGrid(pts, g["cutoff"])
    cells = cg.iter_cells()
    assert len(cells) == g["nonempty_cells"]
    assert sum(len(c) for c in cells) == len(pts)
    assert sum(len(ch) for ch in cg.par_iter_cells(4)) == len(cells)
    g2 = golden["test_neighborcell_particle_pairs"]
    pts2 = oracle.generate_pointcloud(g2["shape"], g2["cutoff"], g2["origin"])
    cg2 = zb.CellGrid(pts2, g2["cutoff"])
    intra = sum(len(c.intra_cell_pairs()) for c in cg2.iter_cells())
    inter = sum(len(c.inter_cell_pairs()) for c in cg2.iter_cells())
    assert (intra, inter) == (g2["intra_half"], g2["inter_half"])
    g3 = golden["test_half_full_space_particle_pairs"]
    assert sum(len(c.intra_cell_pairs(True)) for c in cg2.iter_cells()) == g3["full_over_half_intra"] * intra
    assert sum(len(c.inter_cell_pairs(True)) for c in cg2.iter_cells()) == g3["full_over_half_inter"] * inter
    # the per-cell enumeration in the reference's own half space gives the device's pair set
    host = sorted((min(p[0], q[0]), max(p[0], q[0])) for c in cg2.iter_cells() for p, q in c.particle_pairs())
    dev = sorted((min(int(i), int(j)), max(int(i), int(j))) for i, j in cg2.particle_pairs())
    assert host == dev
    # query(): cell of a point (possibly empty), None when too far outside (cellgrid.rs:360-365)
    cell = cg2.query(pts2[0])
    assert cell is not None and any(l == 0 for l, _ in cell.iter())
    assert cg2.query(pts2.max(0) + 5.0 * g2["cutoff"]) is None


@pytest.mark.parametrize("mask,split", [("7", "0"), ("0", "0"), ("0", "1")])
def test_prefilter_modes_are_bit_exact(mask, split):
    """ZB_PREFILTER=<mask> selects which consumers of an f64 grid run through the f32 guard-band prefilter
    kernel (pair_pf_kernels.cuh; default: count only).  With every consumer on it (7) and with none (0) the
    pair sets, counts and energies must be the oracle's.  Run in a subprocess because the switch is read
    from the environment when a handle is created."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import numpy as np, sys
sys.path.insert(0, %r)
import oracle, zelll_b200
from zelll_b200 import workload
for n, kind in ((20000, "lj"), (6000, "cube"), (4000, "dense")):
    rng = np.random.default_rng(n)
    if kind == "lj":
        pts, c = workload.generate_points_random(n), 10.0
    elif kind == "cube":
        pts, c = rng.random((n, 3)) * 11.0, 1.0
    else:
        pts, c = rng.random((n, 3)) * 2.0, 1.0
    cg = zelll_b200.CellGrid(pts, c)
    og = oracle.OracleCellGrid(pts, c)
    for cmp, oc in (("lt", oracle.CMP_LT), ("le", oracle.CMP_LE)):
        want = og.pairs_canonical(oc, c)
        assert np.array_equal(oracle.canonical_pairs(cg.particle_pairs(c, cmp)), want), (kind, cmp)
        assert cg.pair_count(c, cmp) == len(want)
        e, m = cg.lj_energy(c, cmp, return_pairs=True)
        _, e64, mo = og.lj_energy(oc, c)
        assert m == mo and abs(e - e64) <= 1e-10 * abs(e64), (kind, cmp, e, e64)
    # a filter radius below the cell size, and pairs exactly at the threshold
    assert np.array_equal(oracle.canonical_pairs(cg.particle_pairs(0.37 * c, "le")), og.pairs_canonical(oracle.CMP_LE, 0.37 * c))
    cg.rebuild(pts, c)   # single-pass list into a device buffer with head-room (nothing sized on this build)
    dev = cg.particle_pairs_device(c, "lt", capacity=len(og.pairs_canonical(oracle.CMP_LT, c)) + 3000)
    assert np.array_equal(oracle.canonical_pairs(dev.cpu().numpy().view(np.uint32).reshape(-1, 2)), og.pairs_canonical(oracle.CMP_LT, c))
grid = np.stack(np.meshgrid(*[np.arange(6.0)] * 3, indexing="ij"), -1).reshape(-1, 3)   # distances exactly 1
cg = zelll_b200.CellGrid(grid, 1.0); og = oracle.OracleCellGrid(grid, 1.0)
assert cg.pair_count(1.0, "le") == og.pair_count(oracle.CMP_LE, 1.0) > cg.pair_count(1.0, "lt") == og.pair_count(oracle.CMP_LT, 1.0)
print("prefilter ok")
""" % root
    # split = "1": the exact kernel as two launches (staged tiles, then the tiles that did not fit: ZB_SPLIT)
    env = dict(os.environ, ZB_PREFILTER=mask, ZB_SPLIT=split)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "prefilter ok" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_seeded_fuzz_against_oracle(zb):
    """60 seeded random configurations (size, dimension, dtype, anisotropic box, clustering, cutoff,
    filter radius, comparison): pair sets bit-exact, counts equal, energies within tolerance."""
    rng = np.random.default_rng(20260101)
    for case in range(60):
        ndim = int(rng.choice([2, 3], p=[0.3, 0.7]))
        dtype = np.float32 if rng.random() < 0.4 else np.float64
        n = int(rng.choice([1, 2, 5, 40, 300, 1200, 3000]))
        ext = rng.uniform(0.5, 30.0, ndim) * rng.choice([1.0, 0.05], ndim, p=[0.8, 0.2])
        origin = rng.uniform(-50.0, 50.0, ndim)
        pts = origin + rng.random((n, ndim)) * ext
        if rng.random() < 0.3:  # clump a third of the particles
            k = max(1, n // 3)
            pts[:k] = pts[0] + rng.normal(0.0, 0.2, (k, ndim))
        pts = pts.astype(dtype)
        cutoff = float(rng.uniform(0.3, 4.0))
        radius = cutoff * float(rng.choice([1.0, 0.5, 1.3, 0.9]))
        cg = zb.CellGrid(pts, cutoff, dtype=dtype, ndim=ndim)
        og = OracleCellGrid(pts, cutoff, dtype=dtype, ndim=ndim)
        assert np.array_equal(cg.keys(), og.keys()), case
        assert cg.info().shape().tolist() == og.info()["shape"], case
        for cmp in ("none", "lt", "le"):
            want = og.pairs_canonical(OCMP[cmp], radius)
            assert cg.pair_count(radius, cmp) == len(want), (case, cmp)
            assert np.array_equal(canonical_pairs(cg.particle_pairs(radius, cmp)), want), (case, cmp)
        e_t, e64, m = og.lj_energy(CMP_LE, radius)
        e, mm = cg.lj_energy(radius, "le", return_pairs=True)
        assert mm == m, case
        if np.isfinite(e64) and m:
            rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
            assert abs(e - e64) <= rtol * abs(e64), (case, e, e64)


def test_prefetch_double_buffering(zb):
    """zb_grid_prefetch: the staged copy of the next frame is what the following rebuild consumes."""
    import torch

    frames = [torch.from_numpy(workload.generate_points_random(30000, seed=s)).pin_memory() for s in (1, 2, 3)]
    cg = zb.CellGrid(frames[0].numpy(), 10.0)
    want = [OracleCellGrid(f.numpy(), 10.0).lj_energy(CMP_LT, 10.0) for f in frames]
    cg.prefetch(frames[0].numpy())
    for k in range(6):
        cur, nxt = frames[k % 3].numpy(), frames[(k + 1) % 3].numpy()
        cg.rebuild_mut(cur, None)
        cg.prefetch(nxt)
        e, m = cg.lj_energy(10.0, "lt", return_pairs=True)
        assert m == want[k % 3][2] and abs(e - want[k % 3][1]) <= F64_RTOL * abs(want[k % 3][1])
    cg.prefetch(frames[0].numpy())
    cg.rebuild(frames[1].numpy())  # a different array than the prefetched one: copied normally
    e, m = cg.lj_energy(10.0, "lt", return_pairs=True)
    assert m == want[1][2]


def test_full_size_properties_f32(zb):
    """n = 10^7 in f32 (configs[1]): |z| reaches 5.5e5 where an f32 ulp is 0.06, so coordinates are
    coarsely quantised and coincident particles occur -- counts and pair sets must still be exact."""
    n = 10_000_000
    pts = workload.generate_points_random(n, dtype=np.float32)
    cg = zb.CellGrid(pts, 10.0, dtype=np.float32)
    keys, begin, count = cg.cells()
    assert int(count.sum()) == n and np.all(np.diff(keys) > 0)
    c_le, c_lt = cg.pair_count(10.0, "le"), cg.pair_count(10.0, "lt")
    e, m = cg.lj_energy(10.0, "lt", return_pairs=True)
    assert m == c_lt <= c_le <= cg.pair_count()
    assert 15.5 * n < c_lt < 16.5 * n
    # the 150k particles of largest |z| (coarsest quantisation) form a closed sub-box: bit-exact pair set
    order = np.argsort(pts[:, 2], kind="stable")
    sub = np.ascontiguousarray(pts[order[:150_000]])
    og = OracleCellGrid(sub, 10.0, dtype=np.float32)
    cs = zb.CellGrid(sub, 10.0, dtype=np.float32)
    assert np.array_equal(cs.keys(), og.keys())
    for cmp in ("lt", "le"):
        assert np.array_equal(canonical_pairs(cs.particle_pairs(10.0, cmp)), og.pairs_canonical(OCMP[cmp], 10.0))
    # permutation invariance of the counts
    perm = np.random.default_rng(1).permutation(n)
    cg.rebuild(pts[perm])
    assert cg.pair_count(10.0, "lt") == c_lt and cg.pair_count(10.0, "le") == c_le


def test_stable_cell_storage_matches_reference_order(zb):
    """With zb_grid_set_stable the records of every cell are in input order, exactly the slices the
    reference's stable push produces (storage.rs:77-81); the pair set is unchanged."""
    for kind, n in (("lj", 20000), ("dense", 3000), ("cube", 4000)):
        pts, cutoff = _cloud(kind, n, np.float64)
        cg = zb.CellGrid(pts, cutoff)
        cg.set_stable(True)
        cg.rebuild(pts)
        og = OracleCellGrid(pts, cutoff)
        keys, begin, count = cg.cells()
        labels, xyz = cg.cell_storage()
        okeys, obegin, olen = og.cells()
        olabels, _ = og.cell_storage()
        oslot = {int(k): i for i, k in enumerate(okeys)}
        for k, b, c in zip(keys, begin, count):
            ob, oc = int(obegin[oslot[int(k)]]), int(olen[oslot[int(k)]])
            assert labels[b:b + c].tolist() == olabels[ob:ob + oc].tolist()
        assert np.array_equal(canonical_pairs(cg.particle_pairs(cutoff, "le")), og.pairs_canonical(CMP_LE, cutoff))
        e1 = cg.lj_energy(cutoff, "lt")
        cg.rebuild(pts)
        assert np.array_equal(cg.cell_storage()[0], labels)   # reproducible layout
        _, e64, _ = og.lj_energy(CMP_LT, cutoff)
        assert abs(e1 - e64) <= F64_RTOL * abs(e64)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pair_list_single_pass_into_device_buffer(zb, dtype):
    """A DEVICE destination with a caller-chosen capacity is filled in ONE pass (no counting pass): output slots
    of 1024 rows, holes closed afterwards (EmitConsumer / emit_fix_* kernels).  The list must be the oracle's for
    capacities that are exact, barely larger, not a multiple of the slot size and far too large; a capacity that
    is too small reports the rows needed."""
    import torch
    from zelll_b200 import _ffi

    for kind, n in (("lj", 20000), ("cube", 6000), ("dense", 4000), ("clusters", 3000), ("lj", 33), ("lj", 2), ("lj", 0)):
        pts, cutoff = _cloud(kind, n, dtype) if n else (np.zeros((0, 3), dtype=dtype), 10.0)
        og = OracleCellGrid(pts, cutoff, dtype=dtype)
        cg = zb.CellGrid(pts, cutoff, dtype=dtype)
        for cmp in ("none", "lt", "le"):
            want = og.pairs_canonical(OCMP[cmp], cutoff)
            m = len(want)
            for cap in (m + 5000, m, m + 1, 4 * m + 7):
                cg.rebuild(pts, cutoff)  # a new build: nothing is known about the list's length
                got = cg.particle_pairs_device(cutoff, cmp, capacity=cap)
                assert got.shape[0] == m, (kind, n, cmp, cap, got.shape[0], m)
                rows = got.cpu().numpy().view(np.uint32).reshape(-1, 2)
                assert np.array_equal(canonical_pairs(rows), want), (kind, n, cmp, cap)
            if m > 1:
                cg.rebuild(pts, cutoff)
                with pytest.raises(zb.ZelllB200Error) as e:
                    cg.particle_pairs_device(cutoff, cmp, capacity=m - 1)
                assert e.value.status == _ffi.ERR_CAPACITY and str(m) in str(e.value)
                with pytest.raises(zb.ZelllB200Error) as e:  # far too small: most slots are never written
                    cg.particle_pairs_device(cutoff, cmp, capacity=max(1, m // 7))
                assert e.value.status == _ffi.ERR_CAPACITY and str(m) in str(e.value)
                # the handle keeps working, and a sized fetch is exact
                assert np.array_equal(canonical_pairs(cg.particle_pairs(cutoff, cmp)), want)
    torch.cuda.synchronize()


def test_pair_list_never_writes_beyond_capacity(zb):
    """Own bounds check of the one-pass list (compute-sanitizer is not available on the GPU pool): a guard region
    behind the capacity keeps its canary for capacities that are exact, off the 512-row slot size, and too small."""
    import ctypes as C
    import torch
    from zelll_b200 import _ffi

    pts, cutoff = _cloud("lj", 30000, np.float64)
    cg = zb.CellGrid(pts, cutoff)
    m = cg.pair_count(cutoff, "lt")
    want = canonical_pairs(cg.particle_pairs(cutoff, "lt"))
    guard = 8192
    for cap in (m, m + 1, m + 511, m + 513, 2 * m + 77, m - 1, m // 3, 511, 1):
        cg.rebuild(pts, cutoff)
        buf = torch.full((cap + guard, 2), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        cg.use_stream(torch.cuda.current_stream().cuda_stream)
        n_out = C.c_uint64(0)
        rc = cg._lib.zb_grid_pairs(cg._h, _ffi.CMP_LT, float(cutoff), buf.data_ptr(), cap, C.byref(n_out))
        torch.cuda.synchronize()
        assert int(n_out.value) == m
        assert bool((buf[cap:] == 0x5A5A5A5A).all()), cap
        if cap >= m:
            assert rc == _ffi.OK
            assert np.array_equal(canonical_pairs(buf[:m].cpu().numpy().view(np.uint32)), want)
        else:
            assert rc == _ffi.ERR_CAPACITY


def test_pair_list_full_size_single_pass(zb):
    """n = 10^7: 1.6e8 rows written in one pass into a buffer with head-room.  Size-independent properties that
    pin the set: as many rows as the exact count, all rows distinct, every row a pair inside the cutoff."""
    import torch

    n = 10_000_000
    pts = workload.generate_points_random(n)
    cg = zb.CellGrid(pts, 10.0)
    rows = cg.particle_pairs_device(10.0, "lt", capacity=17 * n)
    m = cg.pair_count(10.0, "lt")
    assert rows.shape[0] == m
    i = rows[:, 0].to(torch.int64) & 0xFFFFFFFF
    j = rows[:, 1].to(torch.int64) & 0xFFFFFFFF
    assert int(i.max()) < n and int(j.max()) < n and bool((i != j).all())
    key = (torch.minimum(i, j) << 32) | torch.maximum(i, j)
    key = torch.sort(key).values
    assert bool((key[1:] != key[:-1]).all())
    del key
    x = torch.from_numpy(pts).to(rows.device)
    d = x[i] - x[j]
    dsq = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    assert bool((dsq < 100.0).all())


def test_stage_profile_mask_and_rebuild_mut_shrinking_box(zb):
    """zb_grid_profile's stage mask records only the selected launches; rebuild_mut over clouds whose
    cell count shrinks and grows again (the table is cleared speculatively for the PREVIOUS build's
    size while the bounding box travels to the host) stays exact."""
    pts, cutoff = _cloud("lj", 30000, np.float64)
    cg = zb.CellGrid(pts, cutoff)
    cg.profile(True, stages=["pair_lj"])
    cg.rebuild_mut(pts, None)
    cg.lj_energy(cutoff, "lt")
    st = cg.profile_read()
    assert st["pair_lj"][1] in (1, 2) and st["pair_lj"][0] > 0  # f64: prefilter kernel + exact kernel over its work list
    assert all(v[1] == 0 for k, v in st.items() if k != "pair_lj")
    cg.profile(True)
    cg.rebuild_mut(pts, None)
    st = cg.profile_read()
    assert st["count"][1] == 1 and st["scatter"][1] == 1 and st["bbox"][1] == 1
    cg.profile(False)
    for m in (30000, 2000, 50, 30000, 1, 12000):   # smaller boxes reuse the pre-cleared table, larger ones re-clear
        sub = pts[:m]
        cg.rebuild_mut(sub, None)
        og = OracleCellGrid(sub, cutoff)
        assert cg.info().n_cells == og.info()["n_cells"]
        assert cg.pair_count(cutoff, "le") == og.pair_count(CMP_LE, cutoff)
        assert np.array_equal(canonical_pairs(cg.particle_pairs(cutoff, "lt")), og.pairs_canonical(CMP_LT, cutoff))


def test_binary_matches_tracked_sources():
    """The library under test is the one the tracked sources build (the .so is git-ignored and ships
    prebuilt): zb_build_id() = SHA-256 of csrc/ + header + flags, as zelll_b200/build.py hashes them."""
    from zelll_b200 import _ffi, build

    if os.environ.get("ZB_LIB"):
        pytest.skip("experiment build selected with ZB_LIB")
    got = _ffi.load().zb_build_id().decode()
    assert got == "zb-build-id:" + build.source_hash()


def test_nan_coordinate_is_never_a_pair(zb):
    """A NaN coordinate is accepted like upstream (Rust's `as i32` puts it in cell 0 of its axis) and can never
    pass a distance filter; the prefiltered count kernel must hand such tiles to the exact kernel.  (The
    bounding box of a cloud with NaNs is order-dependent in a sequential min/max fold, so this is a
    self-consistency test: every consumer agrees, and the finite particles' pairs are those of the cloud
    without the NaN rows whenever the box is unchanged.)"""
    pts, cutoff = _cloud("lj", 5000, np.float64)
    pts = pts.copy()
    bad = [17, 4000]
    pts[17, 1] = np.nan
    pts[4000, 2] = np.nan
    cg = zb.CellGrid(pts, cutoff)
    for cmp in ("lt", "le"):
        got = canonical_pairs(cg.particle_pairs(cutoff, cmp))
        assert cg.pair_count(cutoff, cmp) == len(got)       # prefiltered count == exact pair list
        assert not np.any(np.isin(got, bad))
        e, m = cg.lj_energy(cutoff, cmp, return_pairs=True)
        assert m == len(got) and np.isfinite(e)
    # the same cloud with the NaN rows replaced by far-away-in-cell-space duplicates of a corner: same box
    ref_pts = np.delete(pts, bad, axis=0)
    if np.array_equal(ref_pts.min(0), np.nanmin(pts, 0)) and np.array_equal(ref_pts.max(0), np.nanmax(pts, 0)):
        keep = np.delete(np.arange(len(pts)), bad)
        ref = zb.CellGrid(ref_pts, cutoff)
        want = keep[canonical_pairs(ref.particle_pairs(cutoff, "lt")).astype(np.int64)]
        assert np.array_equal(canonical_pairs(want), canonical_pairs(cg.particle_pairs(cutoff, "lt")))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,n", [("cube", 60000), ("plane", 40000)])
def test_wide_grids_row_tiles(zb, kind, n, dtype):
    """Grids many cells across (a plane of cells holds more records than the stage): tiles are segments of one
    x-row and the stage takes the five row segments of the half shell.  25^3 cells / 70 x 70 x 2 cells."""
    pts, cutoff = _cloud(kind, n, dtype)
    cg, og = _check_against_oracle(zb, pts, cutoff, dtype, 3)
    shape = cg.info().shape()
    assert int(shape[0]) * int(shape[1]) + int(shape[0]) + 10 >= 512  # beyond the staged CSR window
