"""CPU-side checks of the drop-in boundary: libzelll_b200.so builds, loads and exports exactly
the symbols include/zelll_b200.h declares; without a GPU the product path fails loudly (no CPU
fallback); host-side GridInfo helpers reproduce the reference's known answers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "zelll_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from zelll_b200 import build as zb_build

    zb_build.build()
    from zelll_b200 import _ffi

    return _ffi.load()


def test_header_symbols_are_exported_and_bound(lib):
    from zelll_b200 import _ffi

    names = _declared()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/zelll_b200.h but not exported"
    assert sorted(_ffi.SIGNATURES) == names, "ctypes binding and header disagree"
    assert lib.zb_abi_version() == _ffi.ABI_VERSION


def test_zb_info_layout_matches_header():
    from zelll_b200 import _ffi

    # 7 doubles, 6 int32, 2 uint64, 4 int32
    assert C.sizeof(_ffi.ZbInfo) == 7 * 8 + 6 * 4 + 2 * 8 + 4 * 4


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.zb_grid_create(1, 3, 0, C.byref(h)) == 2  # ZB_ERR_CUDA
    assert not h.value
    import zelll_b200

    with pytest.raises(zelll_b200.ZelllB200Error):
        zelll_b200.CellGrid(np.zeros((4, 3)), 1.0)


def test_bad_arguments_do_not_crash(lib):
    h = C.c_void_p()
    assert lib.zb_grid_create(7, 3, 0, C.byref(h)) == 1   # bad dtype
    assert lib.zb_grid_create(1, 4, 0, C.byref(h)) == 1   # bad ndim
    assert lib.zb_grid_create(1, 3, 0, None) == 1
    lib.zb_grid_destroy(None)
    assert lib.zb_last_error(None) == b"null handle"
    assert lib.zb_grid_launch_count(None) == 0


def test_gridinfo_host_helpers_match_reference_golden(golden):
    """GridInfo::{flatten_index, try_cell_index, flat_cell_index} (util.rs:171-297) on the host."""
    from zelll_b200 import _ffi
    from zelll_b200.cellgrid import GridInfo

    g = golden["test_utils"]
    raw = _ffi.ZbInfo()
    for d in range(3):
        raw.inf[d], raw.sup[d] = g["aabb_inf"][d], g["aabb_sup"][d]
        raw.shape[d], raw.strides[d] = g["grid_shape"][d], g["grid_strides"][d]
    raw.cutoff, raw.ndim, raw.dtype, raw.keys_changed = g["cutoff"], 3, _ffi.F64, -1
    info = GridInfo(raw)
    for case in g["cell_index_cases"]:
        assert info.try_cell_index(case["p"]) == case["cell"]
        assert info.flat_cell_index(case["p"]) == case["flat"]
        assert info.flatten_index(case["cell"]) == case["flat"]
    d = golden["doctest_flat_cell_index"]
    assert info.keys_changed is None
    with pytest.raises(IndexError):
        far = [info.origin()[k] - 2.5 * info.cutoff() for k in range(3)]
        info.cell_index(far)
    assert info.try_cell_index([info.origin()[k] - 0.5 * info.cutoff() for k in range(3)]) == [-1, -1, -1]


def test_slab_bounds_partition():
    from zelll_b200.sharded import grid_shape, slab_bounds

    for nz in (1, 2, 7, 111112):
        for world in (1, 2, 3, 8):
            b = [slab_bounds(nz, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == nz
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
    assert grid_shape([0.2, 0.25, 0.3], [2.7, 2.75, 2.8], 1.0, np.float64) == [3, 3, 3]


def test_workload_generators_and_lammps_writer():
    """Synthetic workloads of the reference benches (benches/lj.rs:15-34, 59-66) and the LAMMPS data
    layout of examples/lammps_data.rs:56-80."""
    from zelll_b200 import workload

    pts = workload.generate_points_random(1000)
    a, b, c = workload.lj_box(1000)
    assert (a, b) == (30.0, 30.0) and abs(c - 1000 / 9.0) < 1e-9
    assert np.all(np.abs(pts) <= np.array([a, b, c]) / 2)
    assert np.array_equal(pts, workload.generate_points_random(1000))             # seeded, reproducible
    assert np.array_equal(pts[100:200], workload.generate_points_random(100, vol=(a, b, c), first=100))  # chunkable
    srt = workload.presort_by_z(pts)
    assert np.all(np.diff(srt[:, 2]) >= 0)
    assert np.array_equal(workload.perturb(pts, 0, 0.0), pts)
    text = workload.lammps_data(pts[:3]).splitlines()
    assert text[2] == "3 atoms" and text[3] == "1 atom types"
    assert text[4] == "-15 15 xlo xhi" and text[8] == "Atoms # atomic"
    assert text[10].split()[:2] == ["1", "1"] and float(text[10].split()[2]) == pts[0, 0]


def test_header_is_valid_c_and_links(tmp_path, lib):
    """include/zelll_b200.h must compile as plain C (the boundary a cgo / Rust-sys / ctypes binding sees) and a
    C program must link against the library and get a clean error without a GPU (or a grid with one)."""
    import subprocess

    src = tmp_path / "abi_smoke.c"
    src.write_text(
        '#include "zelll_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  zb_grid* g = 0;\n"
        "  int rc = zb_grid_create(ZB_F64, 3, 0, &g);\n"
        '  printf("abi %d rc %d\\n", zb_abi_version(), rc);\n'
        "  if (rc == ZB_OK) {\n"
        "    double xyz[6] = {0, 0, 0, 0.5, 0.5, 0.5}, c = 1.0, e; unsigned long long m; uint64_t mm;\n"
        "    rc = zb_grid_rebuild(g, xyz, 2, &c);\n"
        "    if (rc == ZB_OK) rc = zb_grid_lj_energy(g, ZB_CMP_LT, c, &e, &mm);\n"
        "    m = mm; printf(\"pairs %llu rc %d\\n\", m, rc);\n"
        "    zb_grid_destroy(g);\n"
        "  }\n"
        "  return (rc == ZB_OK || rc == ZB_ERR_CUDA) ? 0 : 1;\n"
        "}\n"
    )
    exe = tmp_path / "abi_smoke"
    libdir = os.path.join(ROOT, "zelll_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lzelll_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi 1" in out.stdout


def test_build_is_hash_based():
    """build() rebuilds when (and only when) the sources' hash differs from the id compiled into the .so."""
    from zelll_b200 import build

    build.build()
    assert build.built_hash() == build.source_hash() and not build.stale()
