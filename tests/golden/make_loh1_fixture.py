#!/usr/bin/env python3
"""Generates tests/golden/loh1-h{100,50}-setup.npz from the reference's own set-up (oracle/_ref; run in the build container):
what the reference's parser + source discretisation produce for tests/loh1/LOH.1-h*.in that the synthetic set-up of
sw4lite_b200/setup.py does not restate -- the unique grid points of the discretised moment source with their force vectors
(Source::set_grid_point_sources4 / GridPointSource), the source time function sampled at every step (Gaussian and its
second derivative, time_functions.C:284), the receiver's grid point, dt and the number of steps.
The forces of step s are F0 * g[s] (and F0 * gtt[s]): F0 is the force at the step where |g| peaks, g the amplitude relative
to it.   python tests/golden/make_loh1_fixture.py"""
import os
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests.test_gpu_step import SourceMap  # noqa: E402

for name in ("h100", "h50"):
    inp = os.path.join(ROOT, "tests", "golden", "inputs", "LOH.1-%s.in" % name)
    with tempfile.TemporaryDirectory() as tmp:
        ew = refshim.RefEW(inp, tmp)
        G = ew.grids[0]
        src = SourceMap(ew)
        n = ew.nsteps
        times = ew.tstart + ew.dt * np.arange(n)
        f = np.array([src.reduce(ew.eval_forces(t, False)) for t in times])        # (n, nu, 3)
        ftt = np.array([src.reduce(ew.eval_forces(t, True)) for t in times])
        s0 = int(np.argmax(np.abs(f).max(axis=(1, 2))))
        F0 = f[s0].copy()
        q = np.unravel_index(np.argmax(np.abs(F0)), F0.shape)
        g = f[:, q[0], q[1]] / F0[q]
        gtt = ftt[:, q[0], q[1]] / F0[q]
        assert np.abs(f - F0[None] * g[:, None, None]).max() <= 1e-14 * np.abs(f).max()
        assert np.abs(ftt - F0[None] * gtt[:, None, None]).max() <= 1e-14 * np.abs(ftt).max()
        recs, modes = ew.receivers()
        out = os.path.join(ROOT, "tests", "golden", "loh1-%s-setup.npz" % name)
        np.savez_compressed(out, ijk=src.points.astype(np.int32), F0=F0, g=g, gtt=gtt, rec=np.array(recs)[:, 1:4].astype(np.int32),
                            dt=ew.dt, nsteps=n, tstart=ew.tstart, nxyz=np.array([G.nx, G.ny, G.nz]), h=G.h, beta=ew.beta)
        print(name, "sources", len(F0), "steps", n, "dt", ew.dt, "->", out, os.path.getsize(out), "bytes")
