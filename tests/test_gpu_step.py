"""-m gpu: whole time steps of the Cartesian path on the GPU against the reference CPU kernels
stepped by the reference's own EW object (oracle/_ref), on the reference's pointsource test
(config 1: tests/pointsource/pointsource.in, 201x201x101, free surface + supergrid, 23 steps).
Gates: per-step max|gpu-cpu|/max|cpu| <= 1e-12 on the new solution, and the final
`Errors at time ... Linf ... L2 ...` line equal to the golden one to printed precision."""
import os
import numpy as np
import pytest

from oracle import refshim
from tests.fields import relerr

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref/libsw4ref.so not present")
INPUTS = os.path.join(os.path.dirname(__file__), "golden", "inputs")

GOLDEN_LINE = "Errors at time 0.6 Linf = 0.569416 L2 = 0.0245361 norm of solution = 3.7439"   # reference README.md:66
GOLDEN_FILE = (0.6, 0.569416364119, 0.0245360919934, 3.74390307494)  # pytest/reference/pointsource/pointsource-h0p04/PointSourceErr.txt


def block_from_reference(ew, g=0, **kw):
    """build our device block from the arrays the reference's setup produced"""
    from sw4lite_b200.solver import GridBlock
    G = ew.grids[g]
    blk = GridBlock(ew.corder, G.bounds, (G.nx, G.ny, G.nz), G.h, ew.dt, G.onesided, G.bctype, G.wind,
                    sg_order=ew.sgorder if ew.usesg else 0, beta=ew.beta if ew.usesg else 0.0, **kw)
    for name in ("mu", "lambda", "rho", "strx", "stry", "strz", "dcx", "dcy", "dcz", "cox", "coy", "coz"):
        blk.upload(name, ew.array(name, g))
    return blk


class SourceMap:
    """unique source points of grid g and the reduction of per-source forces onto them
    (EW::Force sums the sources sharing a grid point, EW.C:3092-3121)"""

    def __init__(self, ew, g=0):
        idx, _, ident = ew.point_sources()
        self.sel = [r for r in range(len(ident) - 1) if idx[ident[r], 0] == g]
        self.ranges = [(ident[r], ident[r + 1]) for r in self.sel]
        self.points = np.array([idx[a, 1:4] for a, _ in self.ranges], dtype=np.int32).reshape(-1, 3)

    def reduce(self, f):
        out = np.zeros((len(self.ranges), 3))
        for n, (a, b) in enumerate(self.ranges):
            for s in range(a, b):
                out[n] += f[s]
        return out


@needs_ref
def test_pointsource_per_step_parity_and_error_line(tmp_path):
    ew = refshim.RefEW(os.path.join(INPUTS, "pointsource.in"), str(tmp_path))
    assert ew.ngrids == 1 and ew.nsteps == 23 and ew.corder == 1
    blk = block_from_reference(ew)
    src = SourceMap(ew)
    blk.set_source_points(src.points)
    worst = 0.0
    t = ew.tstart
    for step in range(ew.nsteps):
        f = src.reduce(ew.eval_forces(t, False))
        ftt = src.reduce(ew.eval_forces(t, True))
        ew.step()                       # reference CPU step (rotates: new solution is now "U")
        blk.step(f, ftt)
        t += ew.dt
        ours = blk.download("U")
        ref = ew.array("U", 0)
        e = relerr(ours, ref)
        worst = max(worst, e)
        assert e < 1e-12, "step %d: %g" % (step + 1, e)
    errs = ew.pointsource_error(ew.t, [blk.download("U")])
    line = "Errors at time %g Linf = %g L2 = %g norm of solution = %g" % (ew.t, errs[0], errs[1], errs[2])
    print(line, " worst per-step rel. diff %.3g" % worst)
    assert line == GOLDEN_LINE
    for a, b in zip((ew.t,) + tuple(errs), GOLDEN_FILE):
        assert abs(a - b) <= 1e-10 * abs(b)
