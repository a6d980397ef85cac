#!/usr/bin/env python3
"""bench.py -- grid-point updates per second of SW4's explicit elastic time step (the Cartesian
hot path: fused rhs4sg+predictor, rhs4sg+corrector, supergrid damping, boundary conditions, halo
exchange) on a synthetic half-space, z-slab decomposed over the GPUs of one node.

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one rank per GPU)
  python bench.py --impl reference [--steps K --warmup W]    the reference CPU (C/OpenMP) path

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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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
    steps = min(a.steps, 10)
    warmup = min(a.warmup, 2)
    g, ms, threads, sample = time_reference(a.cpu_grid, steps, warmup)
    line = {"impl": "reference", "metric": "grid-point updates/sec per timestep", "value": g, "unit": "Gpts/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": g, "unit": "Gpts/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": g, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(a, n):
    return {"workload": "synthetic Cartesian half-space %dx%dx%d (z-slabs of %d planes per GPU), free surface + supergrid gp=30, "
                        "216-point source, 64 surface receivers; full time step (fused rhs4sg+predictor, rhs4sg+corrector, addsgd4, "
                        "bcfortsg, halo exchange)" % (a.nx, a.ny, a.nzl * n, a.nzl),
            "grid": [a.nx, a.ny, a.nzl * n], "per_gpu": [a.nx, a.ny, a.nzl], "parallelism": "z-slab x%d" % n,
            "l2": "inputs larger than L2 (%.1f GB of fields per GPU)" % (15 * 8 * (a.nx + 4) * (a.ny + 4) * (a.nzl + 4) / 1e9)}


def main_ours(a):
    import torch
    import torch.distributed as dist
    import ctypes as C
    import sw4lite_b200 as S
    from sw4lite_b200.setup import CartesianProblem
    from sw4lite_b200.slabs import HaloExchange, SlabStepper

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = S.init(local)
    N = world
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
    blk = prob.make_block(device=local, rank=rank, nranks=N)
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

    ex = HaloExchange(blk, rank, N, device="cuda") if N > 1 else None
    stepper = SlabStepper(blk, ex) if N > 1 else None

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

    prof = {}
    for name in ("rhs_fast_pred", "rhs_fast_corr", "closure", "rhs_v1", "addsgd", "shell", "bc"):
        tot = C.c_double(0); cnt = C.c_longlong(0)
        lib.sw4b200_profile_read(name.encode(), C.byref(tot), C.byref(cnt))
        if cnt.value:
            prof[name] = {"ms_per_step": tot.value / a.steps, "launches_per_step": cnt.value / a.steps}
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
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01e_traffic.json")))
            ent = tr["pred"]
            kname = ent.get("kernel", kname)
            traffic = ent["dram_bytes_per_point"] * a.nx * a.ny * rows / prof["rhs_fast_pred"]["launches_per_step"]
            traffic_src = "profiles/r01e_traffic.json (ncu dram__bytes_read+write per point of the same kernel, scaled to this launch)"
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": kname,
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_point": BYTES_PASS_A, "points_per_launch": a.nx * a.ny * rows,
                "ms_per_launch": dur, "step_frac_of_hbm": BYTES_PER_POINT_STEP * gpts / N / peak}
    line = {"metric": "grid-point updates/sec per timestep", "value": gpts, "unit": "Gpts/s", "n_gpus": N, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, N),
            "e2e": {"value": gpts_e2e, "unit": "Gpts/s", "h2d_bytes_per_step": int(2 * 3 * 8 * nsel),
                    "d2h_bytes_per_step": int(3 * 8 * nrec), "ms_per_step": ms_e2e / a.steps,
                    "note": "per-step C-ABI (sw4b200_grid_step / _part): host source amplitudes in, host receiver samples out each "
                            "step; the wavefield stays device resident as in the reference's own time loop (EW.C:2455-2477)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": prof}
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


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
