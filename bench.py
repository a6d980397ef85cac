#!/usr/bin/env python3
"""bench.py -- grid-point updates per second of SW4's explicit elastic time step (the Cartesian
hot path: fused rhs4sg+predictor, rhs4sg+corrector, supergrid damping, boundary conditions, halo
exchange) on a synthetic half-space, z-slab decomposed over the GPUs of one node.

  python bench.py [--gpus N --steps K --warmup W]                 our CUDA path (one rank per GPU)
  python bench.py --impl reference [--steps K --warmup W]         the reference CPU (C/OpenMP) path, all host threads
  python bench.py --impl reference-cuda                           the reference's OWN CUDA path on this GPU (timing only)
  python bench.py --config strong|testil256|host ...              the other BASELINE.json configurations (see parse())

One step = one full time step of the whole grid (EW::timesteploop body, reference EW.C:2527-2842).
One grid-point update = one interior grid point advanced by one time step.  Workload per GPU:
nx x ny x nzl interior points (default 2048 x 2048 x 128 = BASELINE.json's weak-scaling sweep), fp64,
free surface on top of slab 0, supergrid layers (gp=30) on the other sides, point sources, surface
receivers.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_POINT_STEP = 216.0      # algorithmic: pass A 15 doubles/pt (R u,um,mu,la,rho; W up,uacc), pass B 12 (R uacc,up,mu,la,rho; W up)
BYTES_PASS_A = 15 * 8.0
BYTES_PASS_B = 12 * 8.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--config", default="sweep", choices=["sweep", "strong", "testil256", "host", "loh1-h100", "loh1-h50"],
                    help="sweep: weak-scaling sweep, nx x ny x nzl per GPU (BASELINE.json config 5, the default line); strong: nx x ny x "
                         "nz-total split over the GPUs (config 5, strong scaling); testil256: standalone rhs4sg on the reference "
                         "harness's 256^3 fields (config 2); host: the reference's own program (main, parser, set-up) on this "
                         "repository's kernels, host/_build/sw4lite_b200, on a generated .in file; loh1-h100 / loh1-h50: the LOH.1 "
                         "layer-over-half-space run of the reference (config 3), whole run, station checked against the golden file")
    ap.add_argument("--nz-total", type=int, default=256, help="--config strong: interior planes of the whole grid")
    ap.add_argument("--host-grid", default="640x640x320", help="grid of --config host and --impl reference-cuda")
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--nzl", type=int, default=128, help="interior planes per GPU (weak scaling)")
    ap.add_argument("--cpu-grid", default="320x320x160", help="grid of the CPU sample (cpu_baseline / --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def ref_input(path, nx, ny, nz, h, steps):
    """a reference .in file of the bench workload's pattern (tests/cartesian/basic.in) at a CPU-sized grid"""
    txt = "\n".join([
        "fileio path=%s verbose=0" % os.path.join(path, "out"),
        "grid nx=%d ny=%d nz=%d h=%g" % (nx, ny, nz, h),
        "time steps=%d" % steps,
        "developer checkfornan=0 reporttiming=0 corder=1 cfl=1.3",
        "supergrid gp=30",
        "block vp=4000 vs=2000 r=2600",
        "source x=%g y=%g z=%g mxy=1e18 t0=0 freq=10 type=C6SmoothBump" % (0.5 * nx * h, 0.5 * ny * h, 0.3 * nz * h),
        "rec x=%g y=%g depth=0 file=sta01 usgsformat=1 sacformat=0" % (0.4 * nx * h, 0.3 * ny * h), ""])
    f = os.path.join(path, "bench.in")
    open(f, "w").write(txt)
    return f


class quiet_stdout:
    """the reference prints its set-up log with printf/cout: keep it off stdout (one JSON line only)"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        try:                                    # C stdio buffers too (NCCL's version banner is a printf)
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def time_reference(grid, steps, warmup):
    """the reference's own CPU kernels (oracle/_ref, C/OpenMP, all host threads) stepping a bounded
    sample of the workload; returns (Gpts/s, ms per step, threads, sample description)"""
    from oracle import refshim
    if not refshim.available():
        raise RuntimeError("oracle/_ref/libsw4ref.so is not built")
    nx, ny, nz = [int(x) for x in grid.split("x")]
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count() or 1
    refshim.set_num_threads(int(os.environ.get("SW4B200_REF_THREADS", ncores)))
    with tempfile.TemporaryDirectory() as tmp, quiet_stdout():
        ew = refshim.RefEW(ref_input(tmp, nx, ny, nz, 10.0, steps + warmup), tmp)
        for _ in range(warmup):
            ew.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            ew.step()
        dt = time.perf_counter() - t0
    pts = nx * ny * nz
    return pts * steps / dt / 1e9, 1e3 * dt / steps, refshim.num_threads(), \
        "%dx%dx%d half-space, supergrid gp=30, 1 point source, corder=1, %d steps after %d warm-up" % (nx, ny, nz, steps, warmup)


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = a.steps, a.warmup
    g, ms, threads, sample = time_reference(a.cpu_grid, steps, warmup)
    cfg = workload_config(a, a.gpus)
    cfg["workload"] = ("CPU sample %s of: " % a.cpu_grid) + cfg["workload"]
    cfg["grid_timed"] = [int(x) for x in a.cpu_grid.split("x")]
    cfg["stepping"] = ("the reference's kernels and order of operations (EW.C:2527-2763) driven step by step through oracle/ref_shim.C "
                       "(the reference's timesteploop symbol is weakened in libsw4ref.so); sequencing pinned by the golden pointsource line")
    line = {"impl": "reference", "metric": "grid-point updates/sec per timestep", "value": g, "unit": "Gpts/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": g, "unit": "Gpts/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": g, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def host_input(path, nx, ny, nz, h, steps):
    """the bench workload's pattern (tests/cartesian/basic.in) as a .in file for the reference's own main(): timing on"""
    txt = "\n".join([
        "fileio path=%s verbose=0" % os.path.join(path, "out"),
        "grid nx=%d ny=%d nz=%d h=%g" % (nx, ny, nz, h),
        "time steps=%d" % steps,
        "developer checkfornan=0 reporttiming=1 corder=1 cfl=1.3",
        "supergrid gp=30",
        "block vp=4000 vs=2000 r=2600",
        "block vp=6000 vs=3464 r=2700 z1=%g" % (0.6 * nz * h),
        "source x=%g y=%g z=%g mxy=1e18 t0=0 freq=10 type=C6SmoothBump" % (0.5 * nx * h, 0.5 * ny * h, 0.3 * nz * h),
        "rec x=%g y=%g depth=0 file=sta01 usgsformat=1 sacformat=0" % (0.4 * nx * h, 0.3 * ny * h), ""])
    f = os.path.join(path, "bench.in")
    open(f, "w").write(txt)
    return f


def run_program(exe, grid, steps):
    """run a build of the reference's program (its own CUDA build, or its host linked to libsw4b200.so) on the generated input
    and read its own timers (`developer reporttiming=1`, EW.C:2863-2882: the summary skips the first step).
    Returns dict(gpts, ms_per_step, total_s, solver_s, columns)"""
    nx, ny, nz = [int(x) for x in grid.split("x")]
    with tempfile.TemporaryDirectory() as tmp:
        inp = host_input(tmp, nx, ny, nz, 10.0, steps)
        t0 = time.perf_counter()
        r = subprocess.run([exe, inp], cwd=tmp, capture_output=True, text=True, timeout=3000)
        wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("%s failed: %s" % (exe, (r.stdout + r.stderr)[-1500:]))
    lines = r.stdout.splitlines()
    cols = None
    for n, l in enumerate(lines):
        if l.strip().startswith("Total") and "Scheme" in l and n + 1 < len(lines):
            cols = []                   # (the numbers are followed by "Clock tick is ..." on the same line)
            for tok in lines[n + 1].split():
                try:
                    cols.append(float(tok))
                except ValueError:
                    break
    solver = [l for l in lines if "Execution time, solver phase" in l]
    if not cols:
        raise RuntimeError("no timing summary in the output of %s: %s" % (exe, r.stdout[-1500:]))
    total = cols[0]
    names = ["total", "bc_comm", "bc_phys", "scheme", "supergrid", "forcing"]
    return {"gpts": nx * ny * nz * (steps - 1) / total / 1e9, "ms_per_step": 1e3 * total / (steps - 1), "total_s": total,
            "solver_phase": solver[-1].strip() if solver else None, "columns": dict(zip(names, cols)), "wall_s": wall,
            "devices_line": next((l.strip() for l in lines if "CUDA device" in l or "Cuda devices" in l), None)}


REF_CUDA_EXE = os.path.join(ROOT, "oracle", "_ref", "cuda", "sw4lite_ref_cuda")
HOST_EXE = os.path.join(ROOT, "host", "_build", "sw4lite_b200")


def program_line(a, impl, exe, what):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(a.steps + 1, 3)
    nx, ny, nz = [int(x) for x in a.host_grid.split("x")]
    base = {"impl": impl, "metric": "grid-point updates/sec per timestep", "unit": "Gpts/s", "n_gpus": 1, "steps": steps - 1,
            "warmup": 1, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic Cartesian half-space %s, free surface + supergrid gp=30, two material blocks, one "
                                   "moment source, one receiver, corder=1: %s; timed by the program's own `developer reporttiming=1` "
                                   "summary (all steps but the first)" % (a.host_grid, what), "grid": [nx, ny, nz]}}
    if not os.path.exists(exe):
        base.update({"unavailable": "%s is not built (needs /root/reference at build time)" % os.path.relpath(exe, ROOT)})
        print(json.dumps(base))
        return
    try:
        r = run_program(exe, a.host_grid, steps)
    except Exception as e:
        base.update({"unavailable": "run failed: %s" % str(e)[-400:]})
        print(json.dumps(base))
        return
    base.update({"value": r["gpts"], "ms_per_step": r["ms_per_step"], "program_timers_s": r["columns"], "solver_phase": r["solver_phase"],
                 "wall_s": r["wall_s"], "devices": r["devices_line"],
                 "e2e": {"value": r["gpts"], "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 24,
                         "note": "the program's own time loop: wavefield device resident, one receiver sample to the host per step"},
                 "step_frac_of_hbm": BYTES_PER_POINT_STEP * r["gpts"] / peaks()[0]})
    print(json.dumps(base))


def main_reference_cuda(a):
    """the reference's own CUDA program (EW_cuda.C / device-routines.C, built unmodified for sm_100 by oracle/build_ref_cuda.py)"""
    program_line(a, "reference-cuda", REF_CUDA_EXE,
                 "the UNMODIFIED reference CUDA build (Makefile.cuda recipe, -arch=sm_100; rhs4_v2 & co., device-routines.C:9992)")


def main_host(a):
    """the reference's own main()/parser/set-up/time loop linked against libsw4b200.so (host/EW_cuda_b200.C)"""
    program_line(a, "ours-host", HOST_EXE,
                 "the reference's own host program (main, .in parser, set-up, EW::timesteploop) on libsw4b200.so through host/EW_cuda_b200.C")


def workload_config(a, n):
    return {"workload": "synthetic Cartesian half-space %dx%dx%d (z-slabs of %d planes per GPU%s), free surface + supergrid gp=30, "
                        "216-point source, 64 surface receivers; full time step (fused rhs4sg+predictor, rhs4sg+corrector, addsgd4, "
                        "bcfortsg, halo exchange)" % (a.nx, a.ny, a.nzl * n, a.nzl, ", strong scaling: total grid fixed" if a.config == "strong" else ""),
            "grid": [a.nx, a.ny, a.nzl * n], "per_gpu": [a.nx, a.ny, a.nzl], "parallelism": "z-slab x%d" % n,
            "l2": "inputs larger than L2 (%.1f GB of fields per GPU)" % (15 * 8 * (a.nx + 4) * (a.ny + 4) * (a.nzl + 4) / 1e9)}


def main_ours(a):
    import torch
    import torch.distributed as dist
    import ctypes as C
    import sw4lite_b200 as S
    from sw4lite_b200.setup import CartesianProblem
    from sw4lite_b200.slabs import SlabStepper

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        with quiet_stdout():          # (NCCL prints its version banner on stdout when NCCL_DEBUG asks for it: keep stdout to the one line)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    lib = S.init(local)
    N = world
    strong = a.config == "strong"
    if strong:
        if a.nz_total % N:
            raise SystemExit("--config strong: --nz-total must be a multiple of the number of GPUs")
        a.nzl = a.nz_total // N          # fixed total grid, planes per GPU shrink with N
    nz = a.nzl * N
    prob = CartesianProblem(a.nx, a.ny, nz, h=10.0, vp=4000.0, vs=2000.0, rho=2600.0, gp=30, beta=0.02, corder=1,
                            layers=[(0.6 * nz * 10.0, 6000.0, 3464.0, 2700.0)])
    # one moment-tensor-like source = 6x6x6 weighted grid-point forces (GridPointSource.C), centred in slab 0
    rng = np.random.default_rng(1)
    ci, cj, ck = a.nx // 2, a.ny // 2, max(8, min(a.nzl // 2, nz - 8))
    for di in range(-3, 3):
        for dj in range(-3, 3):
            for dk in range(-3, 3):
                prob.add_point_force(ci + di, cj + dj, ck + dk, rng.uniform(-1e12, 1e12, 3), freq=2.0, t0=0.0)
    with quiet_stdout():              # (NCCL prints its version banner on stdout when a communicator is created)
        S.lib.comm_init(rank, N)      # the library's own NCCL communicator for the halo exchange (csrc/exchange.cu)
    blk = prob.make_block(device=local, rank=rank, nranks=N, comm=True)
    nrec = 0
    if rank == 0:
        ri = np.linspace(40, a.nx - 40, 8).astype(np.int32); rj = np.linspace(40, a.ny - 40, 8).astype(np.int32)
        rec = np.array([[i, j, 1] for i in ri for j in rj], dtype=np.int32)
        blk.set_receiver_points(rec); nrec = len(rec)
    # smooth non-trivial initial wavefield, defined by global indices so that all slabs agree on the halo planes
    kb = blk.bounds[4]
    for name, ph in (("U", 0.0), ("Um", 0.013)):
        dst = blk.device_ptr(name)
        for c in range(3):
            fk = torch.sin(0.05 * (torch.arange(blk.nk, device="cuda", dtype=torch.float64) + kb) + c + ph)
            fj = torch.cos(0.07 * torch.arange(blk.nj, device="cuda", dtype=torch.float64) + 0.3 * c)
            fi = torch.sin(0.11 * torch.arange(blk.ni, device="cuda", dtype=torch.float64) + 0.7 * c + ph)
            t = (1e-3 * fk[:, None, None] * fj[None, :, None] * fi[None, None, :]).contiguous()
            torch.cuda.synchronize()
            S.lib.check(lib.sw4b200_memcpy_d2d(C.c_void_p(dst + c * 8 * blk.npts), C.c_void_p(t.data_ptr()), 8 * blk.npts, None))
            S.lib.check(lib.sw4b200_sync_device())
            del t
    torch.cuda.empty_cache()

    nsel = len(blk.src_sel)
    total = a.warmup + a.steps
    times = [s * prob.dt for s in range(2 * total + 2)]
    f_all = np.array([prob.forces(t)[blk.src_sel] for t in times]).reshape(len(times), nsel, 3) if nsel else np.zeros((len(times), 0, 3))
    ftt_all = np.array([prob.forces(t, tt=True)[blk.src_sel] for t in times]).reshape(len(times), nsel, 3) if nsel else np.zeros((len(times), 0, 3))
    main = torch.cuda.ExternalStream(lib.sw4b200_stream(0))
    interior = a.nx * a.ny * a.nzl          # points owned by this rank

    def barrier():
        S.lib.check(lib.sw4b200_sync_device())
        if world > 1:
            dist.barrier()
        S.lib.check(lib.sw4b200_sync_device())

    stepper = SlabStepper(blk, None) if N > 1 else None

    def run_steps(first, n, e2e):
        """n steps starting at global step `first`.  e2e: through the per-step public API with host
        source amplitudes in and receiver samples out every step"""
        if N > 1:
            for s in range(first, first + n):
                stepper.step(f_all[s], ftt_all[s])
                if e2e and nrec:
                    blk.record()
        elif e2e:
            for s in range(first, first + n):
                blk.step(f_all[s], ftt_all[s], record=nrec > 0)
        else:
            blk.run(first, n)

    if N == 1:
        blk.set_source_series(f_all, ftt_all)

    def timed(first, n, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        run_steps(first, n, e2e)
        e1.record(main)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # warm-up, then the device-resident number
    run_steps(0, a.warmup, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.sw4b200_kernel_launch_count()
    S.lib.check(lib.sw4b200_profile_reset()); S.lib.check(lib.sw4b200_profile_enable(1))
    ms = timed(a.warmup, a.steps, False)
    S.lib.check(lib.sw4b200_profile_enable(0))
    launches = lib.sw4b200_kernel_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    # end-to-end number: per-step public API, host buffers in and out
    ms_e2e = timed(total, a.steps, True)

    # one-off cost outside every timed number: moving the fields to the GPU once (and the wavefield back for norms / check
    # points).  Measured here on a 1 GB pinned buffer through the library's own copy calls, scaled to this block's arrays.
    one_off = None
    if rank == 0:
        try:
            nb = 1 << 30
            hp = lib.sw4b200_malloc_host(nb); dp_ = lib.sw4b200_malloc(nb)
            rates = []
            for direction in (0, 1):
                best = 1e30
                for _ in range(3):
                    S.lib.check(lib.sw4b200_sync_device())
                    t0 = time.perf_counter()
                    if direction == 0:
                        S.lib.check(lib.sw4b200_memcpy_h2d(C.c_void_p(dp_), C.c_void_p(hp), nb, None))
                    else:
                        S.lib.check(lib.sw4b200_memcpy_d2h(C.c_void_p(hp), C.c_void_p(dp_), nb, None))
                    S.lib.check(lib.sw4b200_sync_device())
                    best = min(best, time.perf_counter() - t0)
                rates.append(nb / best / 1e9)
            lib.sw4b200_free(C.c_void_p(dp_)); lib.sw4b200_free_host(C.c_void_p(hp))
            one_off = {"h2d_gbs_pinned": rates[0], "d2h_gbs_pinned": rates[1], "upload_gb": 9 * 8 * blk.npts / 1e9,
                       "upload_s": 9 * 8 * blk.npts / 1e9 / rates[0], "download_u_gb": 3 * 8 * blk.npts / 1e9,
                       "download_u_s": 3 * 8 * blk.npts / 1e9 / rates[1],
                       "note": "once per run, outside the timed loops: U, Um, mu, lambda, rho of this GPU's block up (9 doubles per point), "
                               "the solution down for norms or a check point (3 doubles per point); %d steps of this workload cost as much "
                               "as the upload" % max(1, int(round(9 * 8 * blk.npts / 1e9 / rates[0] / (ms * 1e-3 / a.steps))))}
        except Exception as e:
            one_off = {"error": str(e)}
    prof = {}
    for name in ("rhs_fast_pred", "rhs_fast_corr", "closure", "rhs_v1", "addsgd", "shell", "bc", "exchange_pred", "exchange_corr"):
        tot = C.c_double(0); cnt = C.c_longlong(0)
        lib.sw4b200_profile_read(name.encode(), C.byref(tot), C.byref(cnt))
        if cnt.value:
            prof[name] = {"ms_per_step": tot.value / a.steps, "launches_per_step": cnt.value / a.steps}
    tiles = {}
    for name in ("tiles_plain", "tiles_general"):   # thread blocks of the interior kernel without / with stretching factors
        tot = C.c_double(0); cnt = C.c_longlong(0)
        lib.sw4b200_profile_read(name.encode(), C.byref(tot), C.byref(cnt))
        tiles[name] = cnt.value
    if tiles["tiles_plain"] + tiles["tiles_general"]:
        prof["interior_tiles"] = {"plain_frac": tiles["tiles_plain"] / (tiles["tiles_plain"] + tiles["tiles_general"])}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    gpts = N * interior * a.steps / (ms * 1e-3) / 1e9
    gpts_e2e = N * interior * a.steps / (ms_e2e * 1e-3) / 1e9
    peak, peak_src = peaks()
    # roofline of the dominant kernel (the fused rhs4sg+predictor pass): algorithmic bytes of the rows it
    # computes / its CUDA-event duration inside the timed region
    roof = None
    if "rhs_fast_pred" in prof:
        onesided_rows = 6 if rank == 0 else 0
        rows = a.nzl - onesided_rows
        per_launch_bytes = BYTES_PASS_A * a.nx * a.ny * rows / prof["rhs_fast_pred"]["launches_per_step"]
        dur = prof["rhs_fast_pred"]["ms_per_step"] / prof["rhs_fast_pred"]["launches_per_step"]
        ach = per_launch_bytes / (dur * 1e-3) / 1e9
        # DRAM traffic of that kernel from the committed ncu --set full capture (bytes per point, scaled to this launch)
        traffic, traffic_src, kname = None, None, "k_rhs_fast4<16,EPI_PRED> (fused rhs4sg + predictor + acceleration)"
        try:
            tfile = next(f for f in ("r02_traffic.json", "r01e_traffic.json") if os.path.exists(os.path.join(ROOT, "profiles", f)))
            tr = json.load(open(os.path.join(ROOT, "profiles", tfile)))
            ent = tr["pred"]
            kname = ent.get("kernel", kname)
            traffic = ent["dram_bytes_per_point"] * a.nx * a.ny * rows / prof["rhs_fast_pred"]["launches_per_step"]
            traffic_src = "profiles/%s (ncu --set full dram__bytes_read+write of the same kernel at %s, per point, scaled to this launch)" % (
                tfile, ent.get("shape", "768x768x96"))
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": kname,
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_point": BYTES_PASS_A, "points_per_launch": a.nx * a.ny * rows,
                "ms_per_launch": dur, "step_frac_of_hbm": BYTES_PER_POINT_STEP * gpts / N / peak}
        # the co-bound: fp64 pipe.  Measured FMA rate of this GPU (register-only chain kernel) against the kernel's fp64
        # instruction count per point (ncu source page, thread instructions DADD+DMUL+DFMA / points: 244 with the derived
        # coefficient arrays and the march without stretching factors on the plain tiles, profiles/r02y; round 1: 261)
        tf, fr = C.c_double(0), C.c_double(0)
        if lib.sw4b200_measure_fp64_peak(C.byref(tf), C.byref(fr)) == 0 and fr.value > 0:
            FP64_INSTR_PER_POINT = 244.0
            roof["fp64_cobound"] = {"measured_fma_tflops": tf.value, "kernel_fp64_instr_per_point": FP64_INSTR_PER_POINT,
                                    "pipe_frac": FP64_INSTR_PER_POINT * a.nx * a.ny * rows / prof["rhs_fast_pred"]["launches_per_step"] / (dur * 1e-3) / fr.value,
                                    "pass_floor_ms": 1e3 * FP64_INSTR_PER_POINT * a.nx * a.ny * rows / prof["rhs_fast_pred"]["launches_per_step"] / fr.value,
                                    "note": "every fp64 instruction (add, mul or fma) occupies the pipe like an FMA"}
    line = {"metric": "grid-point updates/sec per timestep", "value": gpts, "unit": "Gpts/s", "n_gpus": N, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, N),
            "e2e": {"value": gpts_e2e, "unit": "Gpts/s", "h2d_bytes_per_step": int(2 * 3 * 8 * nsel),
                    "d2h_bytes_per_step": int(3 * 8 * nrec), "ms_per_step": ms_e2e / a.steps,
                    "note": "per-step C-ABI (sw4b200_grid_step / _part): host source amplitudes in, host receiver samples out each "
                            "step; the wavefield stays device resident as in the reference's own time loop (EW.C:2455-2477)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": prof}
    line["e2e"]["one_off"] = one_off
    if N > 1:
        line["config"]["exchange"] = {1: "peer-to-peer pushes by the copy engines (CUDA IPC over NVLink), flags for ordering",
                                      0: "NCCL send/receive straight from the field arrays"}.get(lib.sw4b200_grid_exchange_transport(blk.h), "none")
    if N == 1 and not a.no_cpu_baseline:
        try:
            g, cms, threads, sample = time_reference(a.cpu_grid, 4, 1)
            line["cpu_baseline"] = {"value": g, "unit": "Gpts/s", "cores": threads, "kind": "reference", "sample": sample,
                                    "ms_per_step": cms}
        except Exception as e:       # the checker is absent: say so instead of inventing a number
            line["cpu_baseline"] = {"value": None, "unit": "Gpts/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main_loh1(a):
    """BASELINE.json config 3: LOH.1 (tests/loh1/LOH.1-h100.in / -h50.in) on N GPUs as z-slabs: the whole run (t = 0..9 s) is
    timed, and the station trace must reproduce the reference's golden sta10.txt.  Set-up: CartesianProblem.loh1 (equal to the
    reference's own set-up, tests/test_setup.py) + the discretised source of tests/golden/loh1-*-setup.npz."""
    import torch
    import torch.distributed as dist
    import ctypes as C
    import sw4lite_b200 as S
    from sw4lite_b200.setup import CartesianProblem
    from sw4lite_b200.slabs import SlabStepper
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        with quiet_stdout():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    lib = S.init(local)
    with quiet_stdout():
        S.lib.comm_init(rank, world)
    tag = a.config.split("-")[1]
    prob = CartesianProblem.loh1(100.0 if tag == "h100" else 50.0)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "loh1-%s-setup.npz" % tag))
    n = int(fx["nsteps"])
    assert n == prob.nsteps and abs(float(fx["dt"]) - prob.dt) <= 1e-16
    blk = prob.make_block(device=local, rank=rank, nranks=world, comm=True)
    k0, k1 = blk.bounds[4] + 2, blk.bounds[5] - 2
    sel = [m for m, p in enumerate(fx["ijk"]) if k0 <= p[2] <= k1]
    if sel:
        blk.set_source_points(fx["ijk"][sel])
    f_all = fx["F0"][sel][None] * fx["g"][:, None, None]
    ftt_all = fx["F0"][sel][None] * fx["gtt"][:, None, None]
    rec = fx["rec"]
    owner = bool(k0 <= rec[0][2] <= k1)
    if owner:
        blk.set_receiver_points(rec)
    main = torch.cuda.ExternalStream(lib.sw4b200_stream(0))
    stepper = SlabStepper(blk, None) if world > 1 else None

    def zero_state():
        pitch = lib.sw4b200_grid_row_pitch(blk.h)
        nb = 3 * 8 * pitch * blk.nj * blk.nk
        for name in ("U", "Um", "Up"):
            S.lib.check(lib.sw4b200_memset_zero(C.c_void_p(blk.device_ptr(name)), nb, None))
        S.lib.check(lib.sw4b200_sync_device())

    # the slab stepper with the receiver sampled before the arrays rotate
    def slab_step(s):
        b = blk
        f, ftt = (f_all[s], ftt_all[s]) if sel else (None, None)
        b.predictor_part(1, f); b.begin_exchange(None, with_acc=True); b.predictor_part(2, f); b.end_exchange(None, with_acc=True)
        b.enforce_bc()
        b.corrector_part(1, ftt); b.begin_exchange(None); b.corrector_part(2, ftt); b.end_exchange(None)
        b.enforce_bc()
        if owner:
            b.record_resident(s)
        b.cycle()

    def barrier():
        S.lib.check(lib.sw4b200_sync_device())
        if world > 1:
            dist.barrier()
        S.lib.check(lib.sw4b200_sync_device())

    if world == 1:
        blk.set_source_series(f_all, ftt_all)
    # warm-up on the real problem, then back to the initial state (U = Um = 0)
    for s in range(min(a.warmup, n)):
        if world == 1:
            blk.run(s, 1)
        else:
            slab_step(s)
    zero_state()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.sw4b200_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    if world == 1:
        blk.run(0, n)
    else:
        for s in range(n):
            slab_step(s)
    e1.record(main)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.sw4b200_kernel_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    err = -1.0
    if owner:
        trace = blk.fetch_records(0, n)[:, 0, :]
        gold = np.array([l.split() for l in open(os.path.join(ROOT, "tests", "golden", "loh1-%s-sta10" % tag, "sta10.txt")) if not l.startswith("#")],
                        dtype=np.float64)
        scale = np.abs(gold[:, 1:4]).max()
        err = float(np.abs(trace - gold[1:, 1:4]).max() / scale) if gold.shape[0] == n + 1 else float("inf")
    if world > 1:
        t = torch.tensor([ms, err], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, err = float(t[0].item()), float(t[1].item())
    if rank == 0:
        pts = prob.nx * prob.ny * prob.nz
        gpts = pts * n / (ms * 1e-3) / 1e9
        peak, _ = peaks()
        line = {"metric": "grid-point updates/sec per timestep", "value": gpts, "unit": "Gpts/s", "n_gpus": world, "steps": n, "warmup": a.warmup,
                "ms_per_step": ms / n, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "the reference's LOH.1 input (tests/loh1/LOH.1-%s.in)" % tag,
                "config": {"workload": "LOH.1 layer over half-space, %dx%dx%d, h=%g, free surface + supergrid gp=30 on five sides, Gaussian moment "
                                       "source, %d steps (t=9 s), z-slabs over %d GPU(s); whole run timed" % (prob.nx, prob.ny, prob.nz, prob.h, n, world),
                           "grid": [prob.nx, prob.ny, prob.nz], "parallelism": "z-slab x%d" % world},
                "solver_seconds": ms * 1e-3, "station_rel_diff_to_golden": err, "station_ok": bool(0 <= err < 1e-9),
                "e2e": {"value": gpts, "unit": "Gpts/s", "h2d_bytes_per_step": 0 if world == 1 else int(2 * 3 * 8 * len(sel)), "d2h_bytes_per_step": 0,
                        "note": "the run as a user makes it: source table in, station trace out after the last step (24 bytes per step, fetched once)"},
                "gpu_launches": int(launches), "clocks": clocks, "step_frac_of_hbm": BYTES_PER_POINT_STEP * gpts / world / peak}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main_testil(a):
    """BASELINE.json config 2: the standalone rhs4sg kernel on the reference harness's analytic 256^3 fields
    (tests/testil/testil.C:97,209-227,411-420; rate on (n-4)^3 points and 666 flop per point as testil.C:380-390 prints)"""
    import torch
    import ctypes as C
    import sw4lite_b200 as S
    from tests.fields import Box, harness_fields
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    lib = S.init(0)
    n = 256
    box = Box(n, n, n)
    h = 1.0 / (n - 1)
    f = harness_fields(box, h)
    ones = np.ones(n)
    onesided = (C.c_int * 6)(0, 0, 0, 0, 1, 1)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    u, mu, la, sx = dev(f["u"]), dev(f["mu"]), dev(f["la"]), dev(ones)
    lu = torch.zeros(3 * box.npts, dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    main = torch.cuda.ExternalStream(lib.sw4b200_stream(0))
    pts = (n - 4) ** 3

    def apply():
        S.lib.check(lib.sw4b200_rhs4sg(1, *box.bounds, n - 4, onesided, p(lu), p(u), p(mu), p(la), h, p(sx), p(sx), p(sx), None))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > the 126 MB L2: written between timed applications
    for _ in range(a.warmup):
        apply()
    sampler = ClockSampler(0); sampler.start()
    n0 = lib.sw4b200_kernel_launch_count()
    S.lib.check(lib.sw4b200_profile_reset()); S.lib.check(lib.sw4b200_profile_enable(1))
    ms = 0.0
    for _ in range(a.steps):
        with torch.cuda.stream(main):
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main); apply(); e1.record(main)
        S.lib.check(lib.sw4b200_sync_device())
        ms += e0.elapsed_time(e1)
    S.lib.check(lib.sw4b200_profile_enable(0))
    launches = lib.sw4b200_kernel_launch_count() - n0
    clocks = sampler.stop()
    prof = {}
    for name in ("rhs_fast_lu", "rhs_fast2_lu", "closure", "rhs_v1"):
        tot = C.c_double(0); cnt = C.c_longlong(0)
        lib.sw4b200_profile_read(name.encode(), C.byref(tot), C.byref(cnt))
        if cnt.value:
            prof[name] = {"ms_per_step": tot.value / a.steps, "launches_per_step": cnt.value / a.steps}
    # end to end: host arrays in, host array out (sw4b200_rhs4sg_host stages through its own device buffers)
    hlu = np.zeros(3 * box.npts)
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    hu, hmu, hla = f["u"], f["mu"], f["la"]
    t0 = time.perf_counter()
    for _ in range(a.steps):
        S.lib.check(lib.sw4b200_rhs4sg_host(1, *box.bounds, n - 4, onesided, dp(hlu), dp(hu), dp(hmu), dp(hla), h, dp(ones), dp(ones), dp(ones)))
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    gpts = pts * a.steps / (ms * 1e-3) / 1e9
    peak, peak_src = peaks()
    roof = None
    if "rhs_fast_lu" in prof:
        dur = prof["rhs_fast_lu"]["ms_per_step"] / prof["rhs_fast_lu"]["launches_per_step"]
        rows = n - 4 - 12
        ach = 64.0 * (n - 4) ** 2 * rows / (dur * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_rhs_fast4<16,EPI_LU> (interior rows 7..%d)" % (n - 4 - 6), "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_point": 64.0,
                "points_per_launch": (n - 4) ** 2 * rows, "ms_per_launch": dur, "whole_operator_frac_of_hbm": 64.0 * gpts / peak}
    line = {"metric": "grid-point updates/sec of one rhs4sg application", "value": gpts, "unit": "Gpts/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "gflops_at_666_flop_per_point": 666.0 * gpts,
            "config": {"workload": "standalone rhs4sg_rev on the reference kernel harness's analytic fields, %d^3 points (tests/testil), both "
                                   "SBP closures, corder=1; rate on (n-4)^3 points as testil.C:390" % n, "grid": [n, n, n],
                       "l2": "L2 flushed (256 MB written) between timed applications"},
            "e2e": {"value": pts * a.steps / (ms_e2e * 1e-3) / 1e9, "unit": "Gpts/s", "h2d_bytes_per_step": int(8 * 8 * box.npts),
                    "d2h_bytes_per_step": int(3 * 8 * box.npts), "ms_per_step": ms_e2e / a.steps,
                    "note": "sw4b200_rhs4sg_host: pageable host arrays in (u, mu, lambda, lu), lu out, synchronous"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": prof}
    if not a.no_cpu_baseline:
        try:
            from oracle import refshim
            acof, ghcof, bope, _ = refshim.get_stencil_coefficients()
            try:
                ncores = len(os.sched_getaffinity(0))
            except Exception:
                ncores = os.cpu_count() or 1
            refshim.set_num_threads(ncores)
            reps = 3
            refshim.rhs4sg(1, box.bounds, n - 4, (0, 0, 0, 0, 1, 1), acof, bope, ghcof, hlu, hu, hmu, hla, h, ones, ones, ones)
            t0 = time.perf_counter()
            for _ in range(reps):
                refshim.rhs4sg(1, box.bounds, n - 4, (0, 0, 0, 0, 1, 1), acof, bope, ghcof, hlu, hu, hmu, hla, h, ones, ones, ones)
            dt = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": pts / dt / 1e9, "unit": "Gpts/s", "cores": refshim.num_threads(), "kind": "reference",
                                    "sample": "rhs4sg_rev of the reference on the same 256^3 fields, mean of %d applications" % reps,
                                    "gflops_at_666_flop_per_point": 666.0 * pts / dt / 1e9}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "Gpts/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    print(json.dumps(line))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    elif args.impl == "reference-cuda":
        main_reference_cuda(args)
    elif args.config == "host":
        main_host(args)
    elif args.config == "testil256":
        main_testil(args)
    elif args.config.startswith("loh1"):
        main_loh1(args)
    else:
        main_ours(args)
