#!/usr/bin/env python3
"""quick kernel timings on the GPU box (development aid; bench.py is the contract)"""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sw4lite_b200 as S
from tests.fields import Box, random_fields
from tests.gpuutil import ints

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = S.init(0)
box = Box(n, n, n)
r = np.random.default_rng(0)
npts = box.npts
t = lambda m: torch.rand(m, dtype=torch.float64, device="cuda") + 1.0
u, um, up, out = t(3 * npts), t(3 * npts), t(3 * npts), t(3 * npts)
mu, la, rho = t(npts), t(npts), t(npts)
sx, sy, sz = t(n), t(n), t(n)
dc = [t(n) * 0.01 for _ in range(3)]; co = [t(n) * 0.3 for _ in range(3)]
p = lambda x: C.c_void_p(x.data_ptr())
st = torch.cuda.Stream()
sp = C.c_void_p(st.cuda_stream)
os_ = ints((0, 0, 0, 0, 1, 0))
interior = (n - 4) ** 3

def timeit(name, fn, bytes_per_pt, reps=5):
    with torch.cuda.stream(st):
        for _ in range(2): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-28s %8.3f ms  %7.2f Gpts/s  %7.1f GB/s (algorithmic)" % (name, ms, interior / ms / 1e6, interior * bytes_per_pt / ms / 1e6))

chk = lambda rc: S.lib.check(rc)
timeit("rhs4sg (lu)", lambda: chk(lib.sw4b200_rhs4sg(1, *box.bounds, n - 4, os_, p(out), p(u), p(mu), p(la), 0.1, p(sx), p(sy), p(sz), sp)), 64)
timeit("rhs4_pred (fused)", lambda: chk(lib.sw4b200_rhs4_pred(1, *box.bounds, n - 4, os_, p(out), p(u), p(um), p(mu), p(la), p(rho), None, p(sx), p(sy), p(sz), 0.1, 0.01, sp)), 96)
timeit("rhs4_corr (fused, sg4)", lambda: chk(lib.sw4b200_rhs4_corr(1, *box.bounds, n - 4, os_, p(out), p(up), p(u), p(um), p(mu), p(la), p(rho), None, p(sx), p(sy), p(sz), p(dc[0]), p(dc[1]), p(dc[2]), p(co[0]), p(co[1]), p(co[2]), 0.02, 4, 0.1, 0.01, sp)), 120)
timeit("rhs4_corr (fused, no sg)", lambda: chk(lib.sw4b200_rhs4_corr(1, *box.bounds, n - 4, os_, p(out), p(up), p(u), p(um), p(mu), p(la), p(rho), None, p(sx), p(sy), p(sz), p(dc[0]), p(dc[1]), p(dc[2]), p(co[0]), p(co[1]), p(co[2]), 0.0, 0, 0.1, 0.01, sp)), 120)
print("launches", lib.sw4b200_kernel_launch_count())

# curvilinear operator (general kernel)
if len(sys.argv) > 2:
    nc = int(sys.argv[2])
    cb = Box(nc, nc, nc // 2)
    cn = cb.npts
    cu, clu = t(3 * cn), t(3 * cn)
    cmu, cla, cjac = t(cn), t(cn), t(cn)
    cmet = t(4 * cn)
    csx, csy = t(nc), t(nc)
    cint = (nc - 4) ** 2 * (nc // 2 - 4)
    interior = cint
    for top in (0, 1):
        os2 = ints((0, 0, 0, 0, top, 0))
        timeit("rhs4sgcurv (top=%d)" % top, lambda: chk(lib.sw4b200_rhs4sgcurv(1, *cb.bounds, p(cu), p(cmu), p(cla), p(cmet), p(cjac), p(clu), os2, p(csx), p(csy), sp)), 104)
