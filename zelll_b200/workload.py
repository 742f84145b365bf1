"""Synthetic workloads of the reference benches (benches/lj.rs:15-34, 59-66).

The reference draws its points from rand-0.8 `StdRng::seed_from_u64(3079380797442975911)`;
that stream is not reproducible without the crate (and no golden point is stored upstream),
so we reproduce the DISTRIBUTION -- uniform in the centred box -- with a counter-based
splitmix64 stream that is identical on the host (numpy, here) and needs no state.
"""
from __future__ import annotations

import numpy as np

REFERENCE_SEED = 3079380797442975911  # benches/lj.rs:22 (reused as *our* seed)
CUTOFF = 10.0                         # benches/lj.rs:60


def lj_box(n: int, cutoff: float = CUTOFF):
    """Edge lengths (a, b, c) of the benchmark box: 10 particles per cutoff^3 (benches/lj.rs:60-64)."""
    conc = 10.0 / cutoff**3
    a = 3.0 * cutoff
    b = 3.0 * cutoff
    c = (float(n) / conc) / a / b
    return a, b, c


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(n_values: int, seed: int = REFERENCE_SEED, offset: int = 0) -> np.ndarray:
    """n_values doubles in [0, 1) with 53 random bits each (what rand's `Standard` f64 yields)."""
    with np.errstate(over="ignore"):
        ctr = np.arange(offset, offset + n_values, dtype=np.uint64) + np.uint64(seed % (1 << 64))
    bits = _splitmix64(ctr) >> np.uint64(11)
    return bits.astype(np.float64) * (1.0 / (1 << 53))


def generate_points_random(n: int, vol=None, origin=(0.0, 0.0, 0.0), seed: int = REFERENCE_SEED,
                           dtype=np.float64, first: int = 0) -> np.ndarray:
    """(u - 0.5 + origin) * vol per component, u ~ U[0,1) (benches/lj.rs:15-34).

    `first` offsets the particle counter so slabs/chunks of one global cloud can be generated
    independently.  f32 clouds are the f64 cloud cast, as examples/cachemisses.rs:61-63 does.
    """
    if vol is None:
        vol = lj_box(n)
    u = uniform01(3 * n, seed, 3 * first).reshape(n, 3)
    pts = (u - 0.5 + np.asarray(origin, dtype=np.float64)) * np.asarray(vol, dtype=np.float64)
    return np.ascontiguousarray(pts.astype(dtype, copy=False))


def presort_by_z(points: np.ndarray) -> np.ndarray:
    """examples/cachemisses.rs:57-59: sort_unstable_by z."""
    return np.ascontiguousarray(points[np.argsort(points[:, 2], kind="stable")])


def perturb(points: np.ndarray, step: int, amplitude: float, seed: int = REFERENCE_SEED) -> np.ndarray:
    """x += U(-amplitude, amplitude) per coordinate; our definition (the reference has none)."""
    n = points.shape[0]
    u = uniform01(points.size, seed ^ 0x5DEECE66D, (step + 1) * points.size).reshape(n, -1)
    return np.ascontiguousarray((points.astype(np.float64) + (2.0 * u - 1.0) * amplitude).astype(points.dtype))


def lammps_data(points: np.ndarray, vol=None, seed: int = REFERENCE_SEED) -> str:
    """The benchmark cloud as a LAMMPS `read_data` file, the layout of examples/lammps_data.rs:56-80
    (for cross-checking energies against `lj/cut`, more_benches/in.zelllbench.txt, off-box)."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    a, b, c = lj_box(n) if vol is None else vol

    def fmt(v: float) -> str:  # Rust's `{}` for f64: shortest round-trip repr, no exponent padding
        r = repr(float(v))
        return r[:-2] if r.endswith(".0") else r

    lines = [
        f"# {n} random atom positions taken from zelll benchmarks:",
        f"# generate_points_random({n}, [{a!r}, {b!r}, {c!r}], [0.0, 0.0, 0.0], Some({seed}));",
        f"{n} atoms",
        "1 atom types",
        f"-{fmt(0.5 * a)} {fmt(0.5 * a)} xlo xhi",
        f"-{fmt(0.5 * b)} {fmt(0.5 * b)} ylo yhi",
        f"-{fmt(0.5 * c)} {fmt(0.5 * c)} zlo zhi",
        "",
        "Atoms # atomic",
        "# lammps read_data needs an empty line here: https://docs.lammps.org/Errors_details.html#err0016",
    ]
    lines += [f"{i + 1} 1 {fmt(p[0])} {fmt(p[1])} {fmt(p[2])}" for i, p in enumerate(pts)]
    lines.append("")
    return "\n".join(lines) + "\n"


def lammps_input(cutoff: float = CUTOFF, repeat: int = 100, data_file: str = "atomsinabox.txt",
                 max_neighbors: int = 100) -> str:
    """LAMMPS input deck with the settings of the reference's cross-tool comparison
    (more_benches/in.zelllbench.txt:5-38): reduced LJ units, atomic style, a NON-periodic box that
    `read_data ... add merge` grows to the data file's extent, `pair_style lj/cut <cutoff>` with
    epsilon = sigma = 1 (the dimensionless potential of benches/lj.rs:42-47), zero velocities, and a
    binned neighbour list with zero skin that is rebuilt on every step (`neighbor 0.0 bin`,
    `neigh_modify delay 0 every 1 check no`), i.e. one cell-list build + one energy pass per step --
    the same work as one rebuild_mut + lj_energy here.  The thermo output's `PotEng` is the pair energy
    PER ATOM: compare with lj_energy / n."""
    lines = [
        "# generated by zelll_b200.workload.lammps_input: LJ pair energy of the benchmark cloud",
        f"# lmp -in <this file> -var data {data_file}",
        f"variable cutoff index {float(cutoff)!r}",
        f"variable repeat index {int(repeat)}",
        f"variable data file {data_file}",
        f"variable max_neighbors index {int(max_neighbors)}",
        "units lj",
        "atom_style atomic",
        "boundary f f f",
        "lattice none 1.0",
        "region box block -0.1 0.1 -0.1 0.1 -0.1 0.1",
        "create_box 1 box",
        "read_data ${data} add merge",
        "thermo_style yaml",
        "mass 1 1.0",
        "velocity all zero linear",
        "pair_style lj/cut ${cutoff}",
        "pair_coeff 1 1 1.0 1.0",
        "neighbor 0.0 bin",
        "neigh_modify delay 0 every 1 check no one ${max_neighbors}",
        "run ${repeat}",
    ]
    return "\n".join(lines) + "\n"


def celllistmap_settings(n: int, cutoff: float = CUTOFF) -> dict:
    """The CellListMap.jl run of the reference's comparison (more_benches/celllistmap.jl:18-43) as data:
    box sides (the benchmark box, the long side at least 3 cutoffs), cutoff, serial `map_pairwise!` over
    lj(dsq) = 4 t (t - 1), t = (1 / dsq)^3, result divided by n (energy per atom, as LAMMPS prints it),
    coordinates read from columns 3-5 of the LAMMPS data file after its 10 header lines."""
    a, b, c = lj_box(n, cutoff)
    return {
        "package": "CellListMap.jl", "sides": [a, b, max(c, 3.0 * cutoff)], "cutoff": float(cutoff), "parallel": False,
        "pair_function": "lj(dsq) = (t = (1 / dsq)^3; 4.0 * t * (t - 1.0))", "reduce": "sum / n",
        "data_columns": [3, 4, 5], "data_skip_lines": 10,
    }


def celllistmap_script(n: int, cutoff: float = CUTOFF, data_file: str = "atomsinabox.txt") -> str:
    """A Julia script that evaluates the same energy per atom with CellListMap.jl (settings above)."""
    cfg = celllistmap_settings(n, cutoff)
    sides = ", ".join(repr(float(v)) for v in cfg["sides"])
    return "\n".join([
        "# generated by zelll_b200.workload.celllistmap_script",
        "using CSV, DataFrames, CellListMap",
        f'df = DataFrame(CSV.File("{data_file}"; header=false, skipto={cfg["data_skip_lines"] + 1}, select={cfg["data_columns"]}))',
        "xyz = permutedims(Matrix(df))",
        "n = size(xyz, 2)",
        "lj(dsq) = (t = (1 / dsq)^3; 4.0 * t * (t - 1.0))",
        f"box = Box([{sides}], {float(cutoff)!r})",
        "cl = CellList(xyz, box)",
        f"e = map_pairwise!((x, y, i, j, dsq, acc) -> lj(dsq) + acc, 0.0, box, cl, parallel={str(cfg['parallel']).lower()}) / n",
        'println(n, " ", e)',
    ]) + "\n"
