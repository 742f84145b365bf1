#!/usr/bin/env python
"""bench.py -- CellGrid rebuild + Lennard-Jones energy on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = `rebuild_mut(points, cutoff)` + the fused LJ energy pass (`particle_pairs().filter(dsq <
c^2).map(lj).sum()`, benches/lj.rs:100-123) over one synthetic batch: n = 10^7 f64 particles per
GPU, uniform random in the benchmark box (cutoff 10, 10 particles per cutoff^3, 30 x 30 x n/9).
value = in-cutoff neighbor pairs processed per second over the whole job (all ranks), with the
points already resident in HBM; ms_per_step = the rebuild+LJ time; e2e = the same step through the
public API with pinned HOST buffers (H2D of every step's points and D2H of its energy inside the
timed region; the next frame's copy is prefetched while the current one is consumed).  N > 1: weak scaling, the box is slab-decomposed along z, one process per GPU, halo layer
over NCCL send/recv, energy all-reduce.

--impl reference times the reference's CPU algorithm on the host cores: the Rust crate cannot be
compiled in this image (no cargo/rustc), so it is the C++ restatement in oracle/ (`kind: port`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from zelll_b200 import workload  # noqa: E402

CUTOFF = workload.CUTOFF
METRIC = "cellgrid_rebuild_plus_lj_neighbor_pairs_per_s"
UNIT = "pairs/s"
N_PER_GPU = 10_000_000
CPU_SAMPLE_N = 1_000_000


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
class ClockSampler:
    """`nvidia-smi -lms` in the background while the timed region runs (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=10)
            except Exception:
                self.proc.kill()
                out = ""
            for line in out.splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 7:
                    rows.append(parts)

        def num(v):
            try:
                return float(v)
            except ValueError:
                return None

        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if num(r[2]) is not None]
        # "under load" = the samples in the upper half of the power range
        load = sm
        if pw and len(pw) == len(sm) and max(pw) > min(pw):
            thr = 0.5 * (max(pw) + min(pw))
            load = [c for c, w in zip(sm, pw) if w >= thr] or sm
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle = C++ restatement)
def cpu_rebuild_lj(n: int, reps: int, warmup: int = 1):
    """Sequential CellGrid::new (the reference's construction is single-threaded, cellgrid.rs:187-238)
    + par_particle_pairs-shaped LJ over all host threads (cellgrid.rs:447-451).  Returns
    (seconds per step, in-cutoff pairs per step, threads)."""
    import oracle

    try:
        oracle.build(native=True)
        native = True
    except Exception:
        native = False
    threads = max(1, min(os.cpu_count() or 1, oracle.max_threads()))
    pts = workload.generate_points_random(n)
    og = oracle.OracleCellGrid(pts, CUTOFF, native=native)
    times, pairs = [], 0
    for k in range(warmup + reps):
        t0 = time.perf_counter()
        og.rebuild(pts, CUTOFF)
        _, _, pairs = og.lj_energy(oracle.CMP_LT, CUTOFF, nthreads=threads)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
    return sum(times) / len(times), pairs, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 alone
    sec, pairs, threads = cpu_rebuild_lj(CPU_SAMPLE_N, max(1, args.steps), max(1, args.warmup))
    value = pairs / sec
    sample = (f"n={CPU_SAMPLE_N} slice of the n={N_PER_GPU} box (same density and cutoff); sequential rebuild + "
              f"{threads}-thread LJ pass per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args.gpus, N_PER_GPU),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Rust reference not buildable here (no cargo/rustc): C++ restatement of its algorithm (oracle/)",
    }
    print(json.dumps(line))


def _config(n_gpus: int, n_per_gpu: int):
    return {
        "workload": f"rebuild+lj: CellGrid rebuild_mut + LJ energy (dsq < c^2), n={n_per_gpu:.0e} f64 particles per GPU, "
                    f"uniform random box 30x30x(n/9), cutoff 10 (benches/lj.rs; BASELINE.json configs[2]/north_star 10^7)",
        "n_per_gpu": n_per_gpu, "n_total": n_per_gpu * n_gpus, "cutoff": CUTOFF,
        "l2": "inputs (240 MB points + 320 MB cell-sorted records per GPU) exceed the 126 MB L2; no explicit flush",
        "parallelism": "single GPU" if n_gpus == 1 else f"z-slab decomposition over {n_gpus} GPUs, halo send/recv + all-reduce (NCCL)",
    }


# ---------------------------------------------------------------------------------------------
def slab_points(torch, rank: int, world: int, n_per: int, device, spare: int):
    """This rank's slab of the global benchmark box, generated on the device (10^9 x 24 B does not
    fit host RAM).  Global box: 30 x 30 x L, L = world * n_per / 9, centred; layers of height
    `cutoff` are split evenly, and every rank draws uniformly inside its own layers.  Two pinned
    corner particles make the global bounding box (hence the layer count) deterministic."""
    total = world * n_per
    L = total / 9.0
    nz = int(np.floor(L / CUTOFF)) + 1
    from zelll_b200.sharded import slab_bounds

    zb, ze = slab_bounds(nz, world, rank)
    g = torch.Generator(device=device)
    g.manual_seed(workload.REFERENCE_SEED % (2**63) + rank)
    buf = torch.empty((n_per + spare, 3), dtype=torch.float64, device=device)
    u = torch.rand((n_per, 3), dtype=torch.float64, device=device, generator=g)
    buf[:n_per, 0] = (u[:, 0] - 0.5) * 30.0
    buf[:n_per, 1] = (u[:, 1] - 0.5) * 30.0
    zmax_layers = min(ze, L / CUTOFF)  # the last slab ends where the box ends
    lo, hi = zb + 1e-7, zmax_layers - 1e-7
    buf[:n_per, 2] = -L / 2.0 + CUTOFF * (lo + u[:, 2] * (hi - lo))
    if rank == 0:
        buf[0] = torch.tensor([-15.0, -15.0, -L / 2.0], dtype=torch.float64)
    if rank == world - 1:
        # just inside the open upper faces: uniform draws never reach +15, the grid stays 3 x 3 x nz cells
        # like the single-GPU box (a corner AT +15 would open a fourth, empty cell column in x and y)
        buf[n_per - 1] = torch.tensor([15.0 - 1e-9, 15.0 - 1e-9, -L / 2.0 + CUTOFF * (nz - 1) + 0.5 * (L - CUTOFF * (nz - 1))],
                                      dtype=torch.float64)
    del u
    return buf


def _bind_to_gpu_numa_node(index: int) -> None:
    """Pin this rank to the CPUs next to its GPU (NVML's ideal affinity) so that the pinned host
    buffers it allocates -- and the H2D traffic of the e2e leg -- stay on the GPU's NUMA node."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass  # affinity is an optimisation only


def run_ours(args):
    import torch
    import torch.distributed as dist

    import zelll_b200
    from zelll_b200.sharded import NativeSlabGrid

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: zelll_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    _bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: keeps the host frames NUMA-local
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=device)

    n_per = args.n_per_gpu
    steps, warmup = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src = _peaks()
    stream = torch.cuda.current_stream(device)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- set-up: points resident in HBM, pinned host copy for the e2e leg ------------------
    if not distributed:
        host_pts = workload.generate_points_random(n_per)
        pinned = torch.from_numpy(host_pts).pin_memory()
        dev_pts = pinned.to(device, non_blocking=True)
        grid = zelll_b200.CellGrid(dev_pts, CUTOFF, device=local_rank)

        def step_resident():
            grid.rebuild_mut(dev_pts, None)
            return grid.lj_energy(CUTOFF, "lt", return_pairs=True)

        # e2e: frames stream from pinned host memory (trajectory analysis).  Double buffering through
        # the public API: while frame k is being consumed, prefetch() copies frame k+1; every step
        # still moves its own 240 MB over PCIe inside the timed region.
        pinned_b = torch.from_numpy(host_pts.copy()).pin_memory()
        frames = [pinned.numpy(), pinned_b.numpy()]
        state = {"k": 0}
        grid.prefetch(frames[0])

        def step_e2e():
            k = state["k"]
            state["k"] = k + 1
            grid.prefetch(frames[(k + 1) % 2])               # H2D of the next frame, overlapping this step's kernels
            grid.rebuild_mut(frames[k % 2], None)            # finds its own frame staged (else copies itself)
            return grid.lj_energy(CUTOFF, "lt", return_pairs=True)  # host doubles: D2H inside the call

        engine = grid
    else:
        spare = 8192  # the halo is one 3x3-cell layer (~90 particles)
        buf = slab_points(torch, rank, world, n_per, device, spare)
        dg = NativeSlabGrid(dtype=np.float64, device=local_rank)   # NCCL driven from the C ABI
        pinned = None
        if not args.no_e2e:
            pinned = torch.empty((n_per, 3), dtype=torch.float64).pin_memory()
            pinned.copy_(buf[:n_per])
        engine = dg

        def step_resident():
            dg.rebuild_slab_local(buf, n_per, CUTOFF, label_offset=rank * n_per)
            return dg.lj_energy_allreduce(CUTOFF, "lt", return_pairs=True)

        # e2e: this rank's slab streams from pinned host memory, double-buffered on a side stream
        bufs = [buf, None]
        copy_stream = torch.cuda.Stream(device) if not args.no_e2e else None
        state = {"k": 0}
        if not args.no_e2e:
            bufs[1] = torch.empty_like(buf)
            with torch.cuda.stream(copy_stream):
                bufs[0][:n_per].copy_(pinned, non_blocking=True)

        def step_e2e():
            k = state["k"]
            state["k"] = k + 1
            cur, nxt = bufs[k % 2], bufs[(k + 1) % 2]
            stream.wait_stream(copy_stream)                  # this step's H2D has landed
            with torch.cuda.stream(copy_stream):             # next frame's H2D overlaps this step's kernels
                nxt[:n_per].copy_(pinned, non_blocking=True) # (nxt was read by the build before last: done)
            dg.rebuild_slab_local(cur, n_per, CUTOFF, label_offset=rank * n_per)
            return dg.lj_energy_allreduce(CUTOFF, "lt", return_pairs=True)

    engine.use_stream(stream.cuda_stream)

    def timed(fn, k, profile=False, drain=None):
        barrier()
        if profile:
            engine.profile(True, stages=None if profile is True else profile)
        launches0 = engine.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for _ in range(k):
            out = fn()
        if drain is not None:
            drain()  # e2e: the copy prefetched by the last step completes inside the timed region (k copies in all)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stages = engine.profile_read() if profile else None
        if profile:
            engine.profile(False)
        return float(t.item()), out, engine.launch_count - launches0, stages

    # ---- warm-up, then the timed region (device-resident inputs) -------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.2 s to come up: start it ahead of the warm-up
        time.sleep(0.5)
    for _ in range(warmup):
        energy, pairs = step_resident()
    # Inside the timed region only the dominant kernel is bracketed by CUDA events (its live launch
    # duration feeds `roofline`); bracketing all five stages costs ~30 us per step (12 event records).
    ms_total, (energy, pairs), launches, stages = timed(step_resident, steps, profile=("pair_lj",))
    # per-stage breakdown of the other kernels: a short extra pass, outside the timed region
    _, _, _, stages_all = timed(step_resident, min(steps, 5), profile=True)
    for name, v in stages_all.items():
        if name != "pair_lj":
            stages[name] = v
    # ---- the same step end to end from pinned host memory --------------------------------------
    if args.no_e2e:
        ms_e2e, pairs_e = float("nan"), 0
    else:
        for _ in range(2):
            step_e2e()
        drain = (lambda: stream.wait_stream(copy_stream)) if distributed else grid.prefetch_wait
        ms_e2e, (energy_e, pairs_e), _, _ = timed(step_e2e, steps, drain=drain)
    clocks = sampler.stop() if rank == 0 else None

    ms_step = ms_total / steps
    ms_step_e2e = ms_e2e / steps
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch device time from the library's event pairs) --
    alg_bytes = {  # ALGORITHMIC bytes per particle, f64 (SURVEY.md 8d / DESIGN.md section 4)
        "bbox": 24.0, "count": 24.0, "scan": 0.8, "scatter": 24.0 + 24.0 + 4.0, "pair_lj": 24.8,
    }
    n_local = n_per + (int(dg.slab.n_halo) if distributed else 0)
    kernels = {}
    for name, (ms, cnt) in stages.items():
        if cnt and name in alg_bytes:
            per = ms / cnt
            gbs = alg_bytes[name] * n_local / (per * 1e-3) / 1e9
            kernels[name] = {"ms_per_launch": per, "launches": cnt, "algorithmic_GB": alg_bytes[name] * n_local / 1e9,
                             "GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches"])
    traffic, ncu_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get(dom)
        ncu_note = tj.get("_ncu_" + dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                "ms_per_launch": kernels[dom]["ms_per_launch"],
                "candidate_tests_per_s": 81.6 * n_local / (kernels[dom]["ms_per_launch"] * 1e-3) if dom == "pair_lj" else None,
                "ncu": ncu_note,
                "step_algorithmic_GB": 101.6 * n_local / 1e9,
                "step_frac_of_hbm_peak": 101.6 * n_local / 1e9 / (ms_step * 1e-3) / hbm_peak}

    # ---- CPU baseline on this box's host cores (bounded sample) ---------------------------------
    cpu = None
    if not distributed and not args.no_cpu:
        sec, cpairs, threads = cpu_rebuild_lj(CPU_SAMPLE_N, reps=5, warmup=1)
        cpu = {"value": cpairs / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"n={CPU_SAMPLE_N} slice of the same box, 5 steps of sequential rebuild + {threads}-thread LJ "
                         f"pass (C++ restatement of the reference; the Rust crate cannot be built here)",
               "ms_per_step_sample": sec * 1e3}

    total_pairs = int(pairs)
    line = {
        "metric": METRIC, "value": total_pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": _config(world, n_per),
        "e2e": None if args.no_e2e else {"value": int(pairs_e) / (ms_step_e2e * 1e-3), "unit": UNIT,
                                         "ms_per_step": ms_step_e2e, "h2d_bytes_per_step": n_per * 24 * world,
                                         "d2h_bytes_per_step": 16 * world},
        "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
        "clocks": clocks, "pairs_per_step": total_pairs, "energy": energy,
        "particles_per_s": n_per * world / (ms_step * 1e-3),
    }
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (very large --n-per-gpu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
