import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_step import _problem, _initial
nslabs = 2
prob = _problem(nz=41)
u0, um0 = _initial(prob)
whole = prob.make_block()
whole.upload("U", u0); whole.upload("Um", um0)
slabs = [prob.make_block(rank=r, nranks=nslabs) for r in range(nslabs)]
nij = whole.ni * whole.nj
full = lambda a: a.reshape(3, whole.nk, nij)
for s in slabs:
    k0 = s.bounds[4] - whole.bounds[4]
    s.upload("U", np.ascontiguousarray(full(u0)[:, k0:k0 + s.nk]).ravel())
    s.upload("Um", np.ascontiguousarray(full(um0)[:, k0:k0 + s.nk]).ravel())
    print("slab", s.bounds, s.src_sel, s.desc.halo_lo, s.desc.halo_hi, list(s.desc.onesided), list(s.desc.bctype))
buf = [[torch.zeros(6 * nij, dtype=torch.float64, device="cuda") for _ in range(2)] for _ in slabs]
def exchange():
    for r, s in enumerate(slabs):
        for side in (0, 1):
            if (side == 0 and r > 0) or (side == 1 and r < nslabs - 1):
                s.pack(side, buf[r][side])
    for r, s in enumerate(slabs):
        if r > 0: s.unpack(0, buf[r - 1][1])
        if r < nslabs - 1: s.unpack(1, buf[r + 1][0])
def cmp(tag, name="Up"):
    ref = full(whole.download(name))
    for r, s in enumerate(slabs):
        k0 = s.bounds[4] - whole.bounds[4]
        a = s.download(name).reshape(3, s.nk, nij)
        b = ref[:, k0:k0 + s.nk]
        d = np.abs(a - b).max(axis=(0, 2))
        bad = [(int(k + s.bounds[4]), float(x)) for k, x in enumerate(d) if x > 0]
        print(tag, name, "slab", r, "max|diff| by plane k:", bad[:12], "scale", np.abs(b).max())
t = 0.0
for step in range(2):
    f, ftt = prob.forces(t), prob.forces(t, tt=True)
    whole.predictor(f)
    for s in slabs: s.predictor_part(1, f[s.src_sel])
    exchange()
    for s in slabs: s.predictor_part(2, f[s.src_sel])
    cmp("step%d after predictor" % step)
    whole.enforce_bc()
    for s in slabs: s.enforce_bc()
    cmp("step%d after bc" % step)
    whole.corrector(ftt)
    for s in slabs: s.corrector_part(1, ftt[s.src_sel])
    exchange()
    for s in slabs: s.corrector_part(2, ftt[s.src_sel])
    cmp("step%d after corrector" % step)
    whole.enforce_bc(); whole.cycle()
    for s in slabs: s.enforce_bc(); s.cycle()
    cmp("step%d end" % step, "U")
    t += prob.dt
