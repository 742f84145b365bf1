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
def host_threads() -> int:
    """Host threads the CPU arm may use: the cores this process is allowed on, NOT OMP_NUM_THREADS
    (torch.distributed.run exports OMP_NUM_THREADS=1 to every rank)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_rebuild_lj(n: int, reps: int, warmup: int, threads: int, layout: str = "uniform"):
    """Sequential CellGrid::new (the reference's construction is single-threaded, cellgrid.rs:187-238)
    + the LJ pass: threads == 1 is `particle_pairs` (benches/lj.rs:100-123), threads > 1 is
    `par_particle_pairs` over that many rayon-style workers (cellgrid.rs:447-451, benches/iters.rs:81-85).
    Returns (seconds per step, seconds of the rebuild alone, in-cutoff pairs per step)."""
    import oracle

    try:
        oracle.build(native=True)
        native = True
    except Exception:
        native = False
    pts = workload.generate_points_random(n)
    if layout == "presorted":
        pts = workload.presort_by_z(pts)
    og = oracle.OracleCellGrid(pts, CUTOFF, native=native)
    times, builds, pairs = [], [], 0
    for k in range(warmup + reps):
        t0 = time.perf_counter()
        og.rebuild(pts, CUTOFF)
        t1 = time.perf_counter()
        _, _, pairs = og.lj_energy(oracle.CMP_LT, CUTOFF, nthreads=threads)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
            builds.append(t1 - t0)
    return sum(times) / len(times), sum(builds) / len(builds), pairs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 alone
    threads = host_threads()
    n = N_PER_GPU if args.n_per_gpu is None else args.n_per_gpu
    # one step = the stated per-GPU batch (n = 10^7: the whole workload at N = 1); at N > 1 the job is
    # N such batches, the CPU rate is measured on one of them
    n_cpu = min(n, 10_000_000)
    sec, sec_build, pairs = cpu_rebuild_lj(n_cpu, max(1, args.steps), max(1, args.warmup), threads, args.layout)
    value = pairs / sec
    sample = (f"n={n_cpu} particles per step ({'the full single-GPU batch' if n_cpu == n else 'slice of the per-GPU batch'}"
              f"{'' if args.gpus == 1 else f', 1 of the {args.gpus} slabs of the job'}), same density, cutoff and layout; "
              f"sequential rebuild ({sec_build * 1e3:.0f} ms) + {threads}-thread par_particle_pairs LJ pass per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args.gpus, n, args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Rust reference not buildable here (no cargo/rustc): C++ restatement of its algorithm (oracle/)",
    }
    print(json.dumps(line))


def _config(n_gpus: int, n_per_gpu: int, args=None):
    layout = getattr(args, "layout", "uniform")
    scaling = getattr(args, "scaling", "weak")
    box_xy = getattr(args, "box_xy", 3)
    return {
        "workload": f"rebuild+lj: CellGrid rebuild_mut + LJ energy (dsq < c^2), n={n_per_gpu:.3g} f64 particles per GPU, "
                    f"{layout} box {10 * box_xy}x{10 * box_xy}xL at 10 particles per cutoff^3, cutoff 10 (benches/lj.rs; "
                    f"BASELINE.json configs[2]/north_star 10^7" + ("; configs[3] presorted" if layout == "presorted" else "")
                    + ("; configs[4] scaling" if n_gpus > 1 or n_per_gpu != N_PER_GPU else "") + ")",
        "n_per_gpu": n_per_gpu, "n_total": n_per_gpu * n_gpus, "cutoff": CUTOFF, "layout": layout, "scaling": scaling,
        "box_cells_xy": box_xy,
        "l2": f"inputs ({24 * n_per_gpu / 1e6:.0f} MB points + {32 * n_per_gpu / 1e6:.0f} MB cell-sorted records per GPU) "
              f"exceed the 126 MB L2; no explicit flush" if n_per_gpu >= 4_000_000 else
              "inputs smaller than L2: every step rewrites the cell-sorted records, no explicit flush",
        "parallelism": "single GPU" if n_gpus == 1 else f"z-slab decomposition over {n_gpus} GPUs, halo send/recv + all-reduce (NCCL)",
    }


# ---------------------------------------------------------------------------------------------
def slab_points(torch, rank: int, world: int, n_per: int, device, spare: int, box_xy: int = 3, presorted: bool = False):
    """This rank's slab of the global benchmark box, generated on the device (10^9 x 24 B does not
    fit host RAM).  Global box: a x a x L with a = 10 * box_xy and L = world * n_per / (density * a^2),
    centred; layers of height `cutoff` are split evenly, and every rank draws uniformly inside its own
    layers.  Two pinned corner particles make the global bounding box (hence the layer count)
    deterministic.  presorted: the slab's rows are sorted by z (examples/cachemisses.rs:57-59)."""
    total = world * n_per
    a = CUTOFF * box_xy
    L = total * CUTOFF**3 / 10.0 / (a * a)
    nz = int(np.floor(L / CUTOFF)) + 1
    from zelll_b200.sharded import slab_bounds

    zb, ze = slab_bounds(nz, world, rank)
    g = torch.Generator(device=device)
    g.manual_seed(workload.REFERENCE_SEED % (2**63) + rank)
    buf = torch.empty((n_per + spare, 3), dtype=torch.float64, device=device)
    u = torch.rand((n_per, 3), dtype=torch.float64, device=device, generator=g)
    if presorted:
        u[:, 2] = torch.sort(u[:, 2]).values
    buf[:n_per, 0] = (u[:, 0] - 0.5) * a
    buf[:n_per, 1] = (u[:, 1] - 0.5) * a
    zmax_layers = min(ze, L / CUTOFF)  # the last slab ends where the box ends
    lo, hi = zb + 1e-7, zmax_layers - 1e-7
    buf[:n_per, 2] = -L / 2.0 + CUTOFF * (lo + u[:, 2] * (hi - lo))
    if rank == 0:
        buf[0] = torch.tensor([-a / 2.0, -a / 2.0, -L / 2.0], dtype=torch.float64)
    if rank == world - 1:
        # just inside the open upper faces: uniform draws never reach +a/2, the grid stays box_xy^2 x nz cells
        # like the single-GPU box (a corner AT +a/2 would open one more, empty cell column in x and y)
        buf[n_per - 1] = torch.tensor([a / 2.0 - 1e-9, a / 2.0 - 1e-9, -L / 2.0 + CUTOFF * (nz - 1) + 0.5 * (L - CUTOFF * (nz - 1))],
                                      dtype=torch.float64)
    del u
    return buf


def parity_check(torch, dist, rank: int, world: int, local_rank: int, device, n: int = 200_000):
    """configs[4] parity under the driver: the native NCCL slab path (the one the timed region runs)
    against the single-GPU grid of the same cloud -- all-reduced pair counts (`<` and `<=`), LJ energy
    to 1e-10, and the union of the sharded pair lists bit-exact in canonical form."""
    import zelll_b200
    from zelll_b200.sharded import NativeSlabGrid, slab_bounds

    def canonical(p):
        p = np.asarray(p).reshape(-1, 2).astype(np.uint64)
        key = (np.minimum(p[:, 0], p[:, 1]) << np.uint64(32)) | np.maximum(p[:, 0], p[:, 1])
        key.sort()
        return key

    pts = workload.generate_points_random(n)
    single = zelll_b200.CellGrid(pts, CUTOFF, device=local_rank)
    e_ref, m_ref = single.lj_energy(CUTOFF, "lt", return_pairs=True)
    c_ref = single.pair_count(CUTOFF, "le")
    want = canonical(single.particle_pairs(CUTOFF, "lt"))
    order = np.argsort(pts[:, 2], kind="stable")
    spts = pts[order]
    inf_z = spts[0, 2]
    nz = int(np.floor((spts[-1, 2] - inf_z) / CUTOFF)) + 1
    layer = np.floor((spts[:, 2] - inf_z) / CUTOFF).astype(np.int64)
    zb, ze = slab_bounds(nz, world, rank)
    sel = np.nonzero((layer >= zb) & (layer < ze))[0]
    buf = torch.zeros((len(sel) + 4096, 3), dtype=torch.float64, device=device)
    buf[: len(sel)] = torch.from_numpy(spts[sel]).to(device)
    ng = NativeSlabGrid(dtype=np.float64, device=local_rank)
    ng.rebuild_slab_local(buf, len(sel), CUTOFF, label_offset=int(sel[0]) if len(sel) else 0)
    e, m = ng.lj_energy_allreduce(CUTOFF, "lt", return_pairs=True)
    ct = torch.tensor([ng.pair_count(CUTOFF, "le")], dtype=torch.int64, device=device)
    dist.all_reduce(ct)
    local_pairs = order.astype(np.uint64)[ng.particle_pairs(CUTOFF, "lt").astype(np.int64)]
    gathered = [None] * world
    dist.all_gather_object(gathered, np.asarray(local_pairs, dtype=np.uint64))
    got = canonical(np.concatenate(gathered))
    rel = abs(e - e_ref) / abs(e_ref)
    res = {"world": world, "n": n, "path": "native NCCL slab step vs single-GPU grid",
           "pairs": bool(m == m_ref), "pairs_le": bool(int(ct.item()) == c_ref),
           "pair_set": bool(np.array_equal(got, want)), "energy_rel": rel}
    res["ok"] = bool(res["pairs"] and res["pairs_le"] and res["pair_set"] and rel <= 1e-10)
    flag = torch.tensor([0 if res["ok"] else 1], dtype=torch.int64, device=device)
    dist.all_reduce(flag)
    res["ok"] = bool(flag.item() == 0)
    del ng, single
    return res


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

    if args.scaling == "strong":
        n_total = args.n_total if args.n_total else N_PER_GPU
        n_per = n_total // world
    else:
        n_per = N_PER_GPU if args.n_per_gpu is None else args.n_per_gpu
    steps, warmup = args.steps, max(args.warmup, 3)
    box_xy = args.box_xy
    # N > 1: the native NCCL path is checked against the single-GPU grid BEFORE anything is timed
    parity = None
    if distributed and not args.no_parity:
        parity = parity_check(torch, dist, rank, world, local_rank, device)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "multi-GPU parity check failed", "parity_check": parity}))
            dist.destroy_process_group()
            raise SystemExit(3)
    hbm_peak, peak_src = _peaks()
    stream = torch.cuda.current_stream(device)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- set-up: points resident in HBM, pinned host copy for the e2e leg ------------------
    if not distributed:
        a = CUTOFF * box_xy
        host_pts = workload.generate_points_random(n_per, vol=(a, a, n_per * CUTOFF**3 / 10.0 / (a * a)))
        if args.layout == "presorted":
            host_pts = workload.presort_by_z(host_pts)
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
        # the halo is one layer of box_xy^2 cells, ~10 particles each (3x3: ~90 rows)
        halo_cap = max(8192, int(box_xy * box_xy * 10 * 1.5) + 4096)
        spare = halo_cap
        buf = slab_points(torch, rank, world, n_per, device, spare, box_xy, args.layout == "presorted")
        dg = NativeSlabGrid(dtype=np.float64, device=local_rank)   # NCCL driven from the C ABI
        pinned = None
        if not args.no_e2e:
            pinned = torch.empty((n_per, 3), dtype=torch.float64).pin_memory()
            pinned.copy_(buf[:n_per])
        engine = dg

        def step_resident():
            dg.rebuild_slab_local(buf, n_per, CUTOFF, label_offset=rank * n_per, halo_cap=halo_cap)
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
            dg.rebuild_slab_local(cur, n_per, CUTOFF, label_offset=rank * n_per, halo_cap=halo_cap)
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

    # ---- the same step under SUSTAINED load (seconds, not milliseconds): does the burst number hold? ----
    sustained = None
    if args.sustained_s > 0:
        k_sus = max(steps, int(args.sustained_s * 1e3 / max(ms_total / steps, 1e-3)))
        s2 = ClockSampler(local_rank)
        if rank == 0:
            s2.start()
            time.sleep(0.3)
        ms_sus, _, _, _ = timed(step_resident, k_sus)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"steps": k_sus, "ms_per_step": ms_sus / k_sus, "clocks": c2}

    # ---- f32 grids (benches/lj.rs:11, examples/cachemisses.rs:40-47 make f32 a first-class config): a short
    # extra leg on rank 0's GPU, outside the headline: rebuild + pair count at the same n, LJ at n = 10^6
    # (in this box f32 coordinates collide beyond ~10^6 particles: |z| reaches n / 18)
    f32_leg = None
    if not distributed and not args.no_f32:
        def _time(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                out = fn()
            e1.record(stream)
            torch.cuda.synchronize(device)
            return e0.elapsed_time(e1) / reps, out

        p32 = torch.from_numpy(host_pts.astype(np.float32)).to(device)
        g32 = zelll_b200.CellGrid(p32, CUTOFF, dtype=np.float32, device=local_rank)
        g32.use_stream(stream.cuda_stream)
        ms_b, _ = _time(lambda: g32.rebuild_mut(p32, None))
        ms_c, c32 = _time(lambda: g32.pair_count(CUTOFF, "le"))
        n_small = min(n_per, 1_000_000)
        q32 = torch.from_numpy(workload.generate_points_random(n_small, dtype=np.float32)).to(device)
        g32.rebuild(q32, CUTOFF)
        ms_l, e32 = _time(lambda: (g32.rebuild_mut(q32, None), g32.lj_energy(CUTOFF, "lt"))[1])
        f32_leg = {"n": n_per, "rebuild_ms": ms_b, "pair_count_le_ms": ms_c, "pairs_le": int(c32),
                   "n_lj": n_small, "rebuild_plus_lj_ms": ms_l, "energy_f32": float(e32),
                   "algorithmic_bytes_per_particle": {"rebuild": 40.8, "lj": 12.8}}
        del g32, p32, q32

    # ---- the materialised neighbour list (zb_grid_pairs; python/src/lib.rs:283-315 collects the same list on the
    # host): rebuild + ONE pass over the pairs into a device buffer with head-room, outside the headline.  Rows
    # = the pairs the LJ step keeps, so the same pairs/s metric applies (8 B written per pair).
    list_leg = None
    if not distributed and not args.no_pair_list and n_per * 17 * 8 < 40e9:
        cap_rows = int(pairs * 1.05) + 4096
        def list_step():
            with torch.cuda.stream(stream):  # particle_pairs_device runs on torch's current stream
                grid.rebuild_mut(dev_pts, None)
                return grid.particle_pairs_device(CUTOFF, "lt", capacity=cap_rows)
        for _ in range(3):
            rows = list_step()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(steps, 10))
        e0.record(stream)
        for _ in range(reps):
            rows = list_step()
        e1.record(stream)
        torch.cuda.synchronize(device)
        ms_list = e0.elapsed_time(e1) / reps
        list_leg = {"rebuild_plus_pair_list_ms": ms_list, "rows": int(rows.shape[0]), "rows_match_lj_pairs": bool(rows.shape[0] == pairs),
                    "pairs_per_s": rows.shape[0] / (ms_list * 1e-3), "bytes_written": int(rows.shape[0]) * 8,
                    "passes_over_the_pairs": 1}
        del rows

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
    n_local = int(engine.info().n) if distributed else n_per  # own rows + the halo rows counted on the device
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
                "candidate_tests_per_s": 81.6 * n_local / (kernels[dom]["ms_per_launch"] * 1e-3) if dom == "pair_lj" and box_xy == 3 else None,
                "ncu": ncu_note,
                "step_algorithmic_GB": 101.6 * n_local / 1e9,
                "step_frac_of_hbm_peak": 101.6 * n_local / 1e9 / (ms_step * 1e-3) / hbm_peak}

    # FP64 issue ceiling of the exact-arithmetic part (profiles/fp64_peak.json: DADD/DMUL microbenchmark on
    # this GPU model): the reference's test is 9 separately rounded f64 operations per candidate pair
    fpath = os.path.join(ROOT, "profiles", "fp64_peak.json")
    if dom.startswith("pair") and os.path.exists(fpath):
        with open(fpath) as f:
            fj = json.load(f)
        peak_ops = float(fj["dadd"]["lane_ops_per_s"])
        tests_s = roofline["candidate_tests_per_s"] or 0.0
        roofline["fp64"] = {"ops_per_test": 9, "achieved_ops_s": 9 * tests_s, "peak_ops_s": peak_ops,
                            "frac": 9 * tests_s / peak_ops, "peak_source": "profiles/fp64_peak.json (scripts/issue_peaks.cu, DADD)",
                            "note": "equivalent f64 operations of the reference's exact test per second; above 1.0 means the "
                                    "f32 prefilter keeps most tests off the FP64 pipe"}

    # ---- CPU baseline on this box's host cores (bounded sample) ---------------------------------
    cpu = None
    if not distributed and not args.no_cpu:
        nthr = host_threads()
        par = min(nthr, 16)  # benches/iters.rs:81-85 sweeps rayon pools up to 16 threads
        sec_p, build_p, cpairs = cpu_rebuild_lj(CPU_SAMPLE_N, reps=5, warmup=1, threads=par, layout=args.layout)
        sec_s, build_s, _ = cpu_rebuild_lj(CPU_SAMPLE_N, reps=2, warmup=1, threads=1, layout=args.layout)
        sec_a = sec_p
        if nthr != par:
            sec_a, _, _ = cpu_rebuild_lj(CPU_SAMPLE_N, reps=5, warmup=1, threads=nthr, layout=args.layout)
        cpu = {"value": cpairs / sec_a, "unit": UNIT, "cores": nthr, "kind": "port",
               "sample": f"n={CPU_SAMPLE_N} slice of the same box, sequential rebuild + {nthr}-thread LJ pass per step "
                         f"(C++ restatement of the reference; the Rust crate cannot be built here)",
               "ms_per_step_sample": sec_a * 1e3, "host_cpus": os.cpu_count(),
               "sequential": {"value": cpairs / sec_s, "cores": 1, "ms_per_step_sample": sec_s * 1e3,
                              "ms_rebuild_sample": build_s * 1e3, "shape": "benches/lj.rs:100-123 (CellGrid::new + particle_pairs)"},
               "parallel": {"value": cpairs / sec_p, "cores": par, "ms_per_step_sample": sec_p * 1e3,
                            "ms_rebuild_sample": build_p * 1e3, "shape": "benches/iters.rs:81-85 (par_particle_pairs, min(nproc,16) threads)"}}

    total_pairs = int(pairs)
    line = {
        "metric": METRIC, "value": total_pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": _config(world, n_per, args),
        "e2e": None if args.no_e2e else {"value": int(pairs_e) / (ms_step_e2e * 1e-3), "unit": UNIT,
                                         "ms_per_step": ms_step_e2e, "h2d_bytes_per_step": n_per * 24 * world,
                                         "d2h_bytes_per_step": 16 * world},
        "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
        "clocks": clocks, "pairs_per_step": total_pairs, "energy": energy,
        "particles_per_s": n_per * world / (ms_step * 1e-3),
    }
    if sustained is not None:
        line["sustained"] = sustained
    if f32_leg is not None:
        line["f32"] = f32_leg
    if list_leg is not None:
        line["pair_list"] = list_leg
    if parity is not None:
        line["parity_check"] = parity
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-gpu", type=lambda v: int(float(v)), default=None, help="weak scaling: particles per GPU (default 10^7)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--n-total", type=lambda v: int(float(v)), default=None, help="strong scaling: total particles (default 10^7)")
    ap.add_argument("--layout", default="uniform", choices=["uniform", "presorted"],
                    help="presorted = rows sorted by z (examples/cachemisses.rs:57-59, BASELINE.json configs[3])")
    ap.add_argument("--box-xy", type=int, default=3, help="box width in cells along x and y (3 = benches/lj.rs; e.g. 300 for a wide halo)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the pre-timing parity check")
    ap.add_argument("--sustained-s", type=float, default=2.0,
                    help="seconds of back-to-back steps timed after the K-step region (reported as `sustained`; 0 = skip)")
    ap.add_argument("--no-f32", action="store_true", help="skip the f32 leg")
    ap.add_argument("--no-pair-list", action="store_true", help="skip the materialised pair-list leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (very large --n-per-gpu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
