"""sw4lite_b200: B200-native (sm_100a) implementation of SW4's explicit elastic time-stepping hot
path behind a C-ABI (include/sw4b200.h).  This package is the thin Python host side: ctypes
binding (lib.py), the grid-block solver mirror of EW::timesteploop (solver.py), z-slab
decomposition over torch.distributed (slabs.py) and synthetic problem setup (setup.py)."""
from .lib import load, init, Sw4b200Error, GridDesc  # noqa: F401

__version__ = "0.1"
