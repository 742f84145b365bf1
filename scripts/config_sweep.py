"""BASELINE.json configs[0..3] in one run (device-resident input, CUDA-event stage times from the
library): lj-1e5, the construction sweep n = 1e4..1e7 x {f32, f64}, iteration at 1e7, and the
z-presorted cloud; the CPU oracle (C++ restatement of the reference) is timed beside the small sizes.

    python scripts/config_sweep.py > gpurun_out/config_sweep.json
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zelll_b200
from zelll_b200 import workload


def gpu_times(pts, dtype, reps=10):
    t = torch.from_numpy(pts).cuda()
    cg = zelll_b200.CellGrid(t, 10.0, dtype=dtype)
    out = {}
    for name, fn in (("rebuild", lambda: cg.rebuild(t)), ("pair_count_le", lambda: cg.pair_count(10.0, "le")),
                     ("lj_energy_lt", lambda: cg.lj_energy(10.0, "lt"))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        out[name + "_ms"] = (time.perf_counter() - t0) / reps * 1e3
        if name != "rebuild":
            out[name] = r
    return out


def cpu_times(pts, dtype):
    import oracle
    og = oracle.OracleCellGrid(pts, 10.0, dtype=dtype, native=True)
    t0 = time.perf_counter(); og.rebuild(pts, 10.0); tb = time.perf_counter() - t0
    t0 = time.perf_counter(); c = og.pair_count(oracle.CMP_LE, 10.0); tc = time.perf_counter() - t0
    t0 = time.perf_counter(); e = og.lj_energy(oracle.CMP_LT, 10.0); te = time.perf_counter() - t0
    thr = oracle.max_threads()
    t0 = time.perf_counter(); og.lj_energy(oracle.CMP_LT, 10.0, nthreads=thr); tp = time.perf_counter() - t0
    return {"cpu_rebuild_ms": tb * 1e3, "cpu_pair_count_le_ms": tc * 1e3, "cpu_lj_seq_ms": te * 1e3,
            "cpu_lj_par_ms": tp * 1e3, "cpu_threads": thr, "cpu_pair_count_le": c}


def main():
    import oracle
    oracle.build(native=True)
    rows = []
    for n in (10_000, 100_000, 1_000_000, 10_000_000):
        for dtype in (np.float32, np.float64):
            pts = workload.generate_points_random(n, dtype=dtype)
            row = {"config": "build-sweep / iteration", "n": n, "dtype": np.dtype(dtype).name}
            row.update(gpu_times(pts, dtype))
            if n <= 1_000_000:
                row.update(cpu_times(pts, dtype))
                assert row["cpu_pair_count_le"] == row["pair_count_le"]
            rows.append(row)
    pts = workload.presort_by_z(workload.generate_points_random(10_000_000))
    row = {"config": "presorted-1e7", "n": 10_000_000, "dtype": "float64"}
    row.update(gpu_times(pts, np.float64))
    rows.append(row)
    print(json.dumps(rows, indent=1))


main()
