"""helpers for the -m gpu parity tests: device buffers via torch (plumbing only), calls through
the C-ABI of libsw4b200.so"""
import ctypes as C
import numpy as np


def oracle():
    """the checker: the reference itself when oracle/_ref was built, else the pinned restatement"""
    from oracle import refshim, port
    return refshim if refshim.available() else port


class Dev:
    def __init__(self):
        import torch
        import sw4lite_b200 as S
        self.torch = torch
        self.lib = S.init(0)
        self.S = S

    def put(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a)).cuda()
        self.torch.cuda.synchronize()
        return t

    def zeros(self, n):
        t = self.torch.zeros(n, dtype=self.torch.float64, device="cuda")
        self.torch.cuda.synchronize()
        return t

    def get(self, t):
        self.check(self.lib.sw4b200_sync_device())
        return t.cpu().numpy()

    @staticmethod
    def p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def check(self, rc):
        self.S.lib.check(rc)


def ints(a):
    arr = (C.c_int * len(a))(*[int(x) for x in a])
    return arr
