"""-m gpu: the drop-in boundary exercised by the reference's OWN host program.  host/_build/sw4lite_b200 is the
reference's main(), .in parser, set-up, sources, receivers and error norms (compiled unmodified from
/root/reference/src by host/build_host.py in the build container) with its GPU operator layer replaced by
host/EW_cuda_b200.C + libsw4b200.so.  The program is run on the reference's regression inputs and must
reproduce the reference's golden outputs."""
import os
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "..", "host", "_build", "sw4lite_b200")
INPUTS = os.path.join(HERE, "golden", "inputs")
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="host/_build/sw4lite_b200 not built (needs /root/reference at build time)")


def run(infile, cwd):
    r = subprocess.run([EXE, os.path.join(INPUTS, infile)], cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def station(path):
    return np.array([l.split() for l in open(path) if not l.startswith("#")], dtype=np.float64)


@needs_exe
def test_reference_program_pointsource_error_line(tmp_path):
    """config 1: tests/pointsource/pointsource.in through the reference's own main() on the GPU kernels"""
    out = run("pointsource.in", str(tmp_path))
    line = [l for l in out.splitlines() if l.startswith("Errors at time")]
    assert line and line[-1].strip() == "Errors at time 0.6 Linf = 0.569416 L2 = 0.0245361 norm of solution = 3.7439", out[-2000:]
    f = [p for p in tmp_path.rglob("PointSourceErr.txt")]
    assert f
    vals = [float(x) for x in open(f[0]).read().split()]
    gold = (0.6, 0.569416364119, 0.0245360919934, 3.74390307494)
    for a, b in zip(vals, gold):
        assert abs(a - b) <= 1e-10 * abs(b)
    assert "CUDA device" in out and "sw4b200" in out          # the GPU branch of the time loop ran


@needs_exe
def test_reference_program_topography_stations(tmp_path):
    """config 4 pattern: pytest/reference/topo/curvilinear.in (curvilinear grid + Cartesian grid); the reference's
    TimeSeries objects write the station files, which must match the reference's golden ones"""
    run("curvilinear.in", str(tmp_path))
    for name in ("sta01.txt", "sta02.txt", "sta03.txt"):
        f = [p for p in tmp_path.rglob(name)]
        assert f, name
        mine = station(str(f[0]))
        gold = station(os.path.join(HERE, "golden", "curvilinear-output", name))
        assert mine.shape == gold.shape
        assert np.abs(mine[:, 1:] - gold[:, 1:]).max() <= 1e-10 * np.abs(gold[:, 1:]).max(), name


def timing_summary(out):
    """(total seconds of steps 2..n, columns) from the program's `developer reporttiming=1` summary (EW.C:5317-5358)"""
    lines = out.splitlines()
    for n, l in enumerate(lines):
        if l.strip().startswith("Total") and "Scheme" in l:
            cols = []
            for tok in lines[n + 1].split():
                try:
                    cols.append(float(tok))
                except ValueError:
                    break
            return cols
    return None


@needs_exe
@pytest.mark.parametrize("name,npts,nsteps", [("LOH.1-h100", 301 * 301 * 171, 536), ("LOH.1-h50", 601 * 601 * 341, 1073)])
def test_reference_program_loh1_station(tmp_path, name, npts, nsteps):
    """config 3: tests/loh1/LOH.1-h{100,50}.in through the reference's own main() on the grid-block kernels (odd ni: rows
    padded on the device, TMA kernels); the station file the reference's TimeSeries writes must match the reference's golden
    sta10.txt; the program's own timers give the solver rate"""
    out = run(name + ".in", str(tmp_path))
    f = [p for p in tmp_path.rglob("sta10.txt")]
    assert f, out[-2000:]
    mine = station(str(f[0]))
    gold = station(os.path.join(HERE, "golden", "loh1-%s-sta10" % name.split("-")[1], "sta10.txt"))
    assert mine.shape == gold.shape
    scale = np.abs(gold[:, 1:4]).max()
    err = np.abs(mine[:, 1:4] - gold[:, 1:4]).max() / scale
    cols = timing_summary(out)
    rate = npts * (nsteps - 1) / cols[0] / 1e9 if cols else float("nan")
    print("%s through the C++ host: station rel. diff %.3g; solver %.3f s for steps 2..%d = %.2f Gpts/s (scheme %.3f, supergrid %.3f, bc %.3f)"
          % (name, err, cols[0], nsteps, rate, cols[3], cols[4], cols[2]))
    assert err < 1e-9


CHK_BASE = """fileio verbose=1 path=%s
grid x=2.4 y=2.0 z=1.6 h=0.04
time t=0.6
testpointsource rho=1 cp=1.6 cs=0.8 halfspace=1
supergrid gp=10
source x=1.2 y=1.0 z=0.6 Mxx=1 Myy=1 Mzz=1 Mxy=0 Mxz=0 Myz=0 t0=0 freq=1 type=C6SmoothBump
developer checkfornan=0 cfl=1.3 reporttiming=0 corder=0
"""


@needs_exe
def test_reference_program_checkpoint_restart(tmp_path):
    """f4: check point / restart of a device-resident run in the reference's file format, through the reference's own CheckPoint
    class (CheckPoint.C:251-370, EW.C:2778-2791, 2403-2415; corder=0: the reference's extract_subarray addresses (c,i,j,k)).
    The run restarted from the file written at cycle 10 must end with the error norms of the uninterrupted run."""
    def go(tag, extra):
        d = tmp_path / tag
        d.mkdir()
        inp = d / "run.in"
        inp.write_text(CHK_BASE % (str(d / "out")) + extra)
        r = subprocess.run([EXE, str(inp)], cwd=str(d), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        err = [p for p in d.rglob("PointSourceErr.txt")]
        assert err, r.stdout[-2000:]
        return [float(x) for x in open(err[0]).read().split()], d, r.stdout
    full, _, _ = go("full", "")
    _, d1, out1 = go("first", "checkpoint cycle=10 file=chk\n")
    files = [p for p in d1.rglob("*.sw4checkpoint")]
    assert len(files) == 1, out1[-2000:]
    again, _, out2 = go("second", "restart file=%s\n" % str(files[0]))
    assert "reading check point" in out2
    assert full[1] > 0
    for a, b in zip(again, full):
        assert abs(a - b) <= 1e-12 * abs(b), (again, full)


SLAB_EXE = os.path.join(HERE, "..", "host", "_build", "slab_driver")


@pytest.mark.skipif(not os.path.exists(SLAB_EXE), reason="host/_build/slab_driver not built")
def test_cxx_slab_driver_matches_the_python_driver():
    """the C++ z-slab driver (host/slab_driver.C: its own set-up in C++, the C-ABI phases, exchange entry points) on one GPU
    against the Python driver on the same problem: same dt, same final wavefield (set-up formulas agree to rounding)"""
    import json
    from sw4lite_b200.setup import CartesianProblem
    nx, ny, nz, steps, warm = 68, 36, 40, 4, 2
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([SLAB_EXE, "--nx", str(nx), "--ny", str(ny), "--nzl", str(nz), "--steps", str(steps), "--warmup", str(warm),
                        "--gp", "8", "--h", "100"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    prob = CartesianProblem(nx, ny, nz, h=100.0, gp=8, beta=0.02, corder=1, layers=[(0.6 * nz * 100.0, 6000.0, 3464.0, 2700.0)])
    assert abs(out["dt"] - prob.dt) <= 1e-15 * prob.dt
    lcg, M = 12345, (1 << 64) - 1
    ci, cj, ck = nx // 2, ny // 2, max(8, min(32, nz - 8))
    for di in range(-3, 3):
        for dj in range(-3, 3):
            for dk in range(-3, 3):
                a = []
                for c in range(3):
                    lcg = (lcg * 6364136223846793005 + 1442695040888963407) & M
                    a.append(((lcg >> 11) / 9007199254740992.0 * 2 - 1) * 1e12)
                prob.add_point_force(ci + di, cj + dj, ck + dk, a, freq=2.0)
    blk = prob.make_block()
    ni, nj, nk = prob.ni, prob.nj, prob.nk
    u = np.zeros((2, 3, nk, nj, ni))
    for c in range(3):
        for ph, phs in ((0, 0.0), (1, 0.013)):
            fi = np.sin(0.11 * np.arange(ni) + 0.7 * c + phs); fj = np.cos(0.07 * np.arange(nj) + 0.3 * c)
            fk = 1e-3 * np.sin(0.05 * (np.arange(nk) + prob.bounds[4]) + c + phs)
            u[ph, c] = fk[:, None, None] * fj[None, :, None] * fi[None, None, :]
    blk.upload("U", u[0].ravel()); blk.upload("Um", u[1].ravel())
    for n in range(warm + steps):
        blk.step(prob.forces(n * prob.dt), prob.forces(n * prob.dt, tt=True))
    own = blk.download("U").reshape(3, nk, nj, ni)[:, 2:-2, 2:-2, 2:-2]
    ss, mx = float((own ** 2).sum()), float(np.abs(own).max())
    print("C++ slab driver vs Python: sum_sq %.15g / %.15g, max %.15g / %.15g" % (out["checksum"]["sum_sq"], ss, out["checksum"]["max_abs"], mx))
    assert mx > 0 and abs(out["checksum"]["max_abs"] - mx) <= 1e-9 * mx and abs(out["checksum"]["sum_sq"] - ss) <= 1e-9 * ss


@needs_exe
def test_reference_program_gaussian_hill_station(tmp_path):
    """config 4: tests/topo/gaussianHill.in (Gaussian hill topography: 101 x 101 x 51 Cartesian + 101 x 101 x 110 curvilinear points,
    corder=no, 789 steps to t=3) through the reference's own main() on this repository's kernels; station sta04 against the
    reference's golden file (tests/topo/gaussianHill-sta-04/sta04.txt)"""
    out = run("gaussianHill.in", str(tmp_path))
    f = [p for p in tmp_path.rglob("sta04.txt")]
    assert f, out[-2000:]
    mine = station(str(f[0]))
    gold = station(os.path.join(HERE, "golden", "gaussianHill-sta-04", "sta04.txt"))
    assert mine.shape == gold.shape and len(gold) == 790
    scale = np.abs(gold[:, 1:4]).max()
    err = np.abs(mine[:, 1:4] - gold[:, 1:4]).max() / scale
    print("gaussianHill through the C++ host: sta04 rel. diff %.3g over %d samples (amplitude %.3g)" % (err, len(gold), scale))
    assert scale > 0 and err < 1e-9


REF_EXE = os.path.join(HERE, "..", "oracle", "_ref", "sw4lite_ref")


@needs_exe
@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/sw4lite_ref not built")
def test_reference_program_gaussian_hill_rev_against_the_cpu_reference(tmp_path):
    """config 4 at its multi-rank size: tests/topo/gaussianHill-rev.in (128 x 128 x 1900 Cartesian + 128 x 128 x 106 curvilinear
    points, corder=yes, 100 steps).  No golden file exists for it, so the unmodified CPU reference program (oracle/_ref) runs the
    same input beside the GPU build; the four station files must agree"""
    a, b = tmp_path / "gpu", tmp_path / "cpu"
    a.mkdir(); b.mkdir()
    out = run("gaussianHill-rev.in", str(a))
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 8))
    r = subprocess.run([REF_EXE, os.path.join(INPUTS, "gaussianHill-rev.in")], cwd=str(b), capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    worst, amp = 0.0, 0.0
    for name in ("sta01.txt", "sta02.txt", "sta03.txt", "sta04.txt"):
        fa, fb = list(a.rglob(name)), list(b.rglob(name))
        assert fa and fb, name
        ma, mb = station(str(fa[0])), station(str(fb[0]))
        assert ma.shape == mb.shape and len(mb) == 101
        scale = np.abs(mb[:, 1:4]).max()
        amp = max(amp, scale)
        if scale > 0:
            worst = max(worst, np.abs(ma[:, 1:4] - mb[:, 1:4]).max() / scale)
    cols, cols_cpu = timing_summary(out), timing_summary(r.stdout)
    pts = 128 * 128 * 1900 + 128 * 128 * 106
    print("gaussianHill-rev: stations GPU vs CPU reference rel. diff %.3g (amplitude %.3g); solver GPU %.3f s = %.2f Gpts/s, CPU reference %.1f s = %.3f Gpts/s (%d threads)"
          % (worst, amp, cols[0], pts * 99 / cols[0] / 1e9, cols_cpu[0], pts * 99 / cols_cpu[0] / 1e9, os.cpu_count() or 8))
    assert amp > 0 and worst < 1e-9
