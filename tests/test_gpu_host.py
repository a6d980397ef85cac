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
