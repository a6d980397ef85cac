"""ctypes binding of libsw4b200.so (include/sw4b200.h).  No CPU fallback: importing works
anywhere (so the symbol table can be checked), but every compute entry point needs
sw4b200_init() to have succeeded on a CUDA device."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (SW4B200_LIB: a variant build of the same library, for A/B measurements of kernel changes)
LIBPATH = os.environ.get("SW4B200_LIB") or os.path.join(HERE, "libsw4b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_llp = C.POINTER(C.c_longlong)
VP = C.c_void_p
I = C.c_int
D = C.c_double


class GridDesc(C.Structure):
    _fields_ = [("corder", I), ("ifirst", I), ("ilast", I), ("jfirst", I), ("jlast", I), ("kfirst", I),
                ("klast", I), ("nx", I), ("ny", I), ("nz", I), ("h", D), ("dt", D), ("onesided", I * 6),
                ("bctype", I * 6), ("wind", I * 36), ("sg_order", I), ("beta", D), ("curvilinear", I),
                ("halo_lo", I), ("halo_hi", I)]


# name -> (restype, argtypes); every symbol declared in include/sw4b200.h
B6 = [I] * 6
SIGNATURES = {
    "sw4b200_init": (I, [I]),
    "sw4b200_finalize": (I, []),
    "sw4b200_device_count": (I, []),
    "sw4b200_last_error": (C.c_char_p, []),
    "sw4b200_version": (C.c_char_p, []),
    "sw4b200_stream": (VP, [I]),
    "sw4b200_sync_stream": (I, [I]),
    "sw4b200_sync_device": (I, []),
    "sw4b200_kernel_launch_count": (I, []),
    "sw4b200_profile_enable": (I, [I]),
    "sw4b200_set_option": (I, [C.c_char_p, I]),
    "sw4b200_profile_reset": (I, []),
    "sw4b200_profile_read": (I, [C.c_char_p, c_dp, c_llp]),
    "sw4b200_malloc": (VP, [C.c_size_t]),
    "sw4b200_free": (I, [VP]),
    "sw4b200_malloc_host": (VP, [C.c_size_t]),
    "sw4b200_free_host": (I, [VP]),
    "sw4b200_memcpy_h2d": (I, [VP, VP, C.c_size_t, VP]),
    "sw4b200_memcpy_d2h": (I, [VP, VP, C.c_size_t, VP]),
    "sw4b200_memcpy_d2d": (I, [VP, VP, C.c_size_t, VP]),
    "sw4b200_memset_zero": (I, [VP, C.c_size_t, VP]),
    "sw4b200_get_stencil_coefficients": (I, [c_dp] * 4),
    "sw4b200_copy_stencilcoefficients": (I, [c_dp] * 4),
    "sw4b200_rhs4sg": (I, [I] + B6 + [I, c_ip] + [VP] * 4 + [D] + [VP] * 3 + [VP]),
    "sw4b200_predfort": (I, [I] + B6 + [VP] * 6 + [D, VP]),
    "sw4b200_corrfort": (I, [I] + B6 + [VP] * 4 + [D, VP]),
    "sw4b200_dpdmtfort": (I, B6 + [VP] * 4 + [D, VP]),
    "sw4b200_addsgd": (I, [I, I] + B6 + [VP] * 13 + [D, VP]),
    "sw4b200_bcfortsg": (I, [I] + B6 + [c_ip, I, I, I, VP, D, c_ip, VP, VP, C.POINTER(VP), VP, VP, VP]),
    "sw4b200_rhs4sgcurv": (I, [I] + B6 + [VP] * 6 + [c_ip, VP, VP, VP]),
    "sw4b200_addsgdc": (I, [I, I] + B6 + [VP] * 11 + [D, VP]),
    "sw4b200_freesurfcurvisg": (I, [I] + B6 + [I, I] + [VP] * 7 + [VP]),
    "sw4b200_enforce_cart_topo": (I, [I, VP] + B6 + [VP, I, I, VP]),
    "sw4b200_rhs4_pred": (I, [I] + B6 + [I, c_ip] + [VP] * 10 + [D, D, VP]),
    "sw4b200_rhs4_corr": (I, [I] + B6 + [I, c_ip] + [VP] * 17 + [D, I, D, D, VP]),
    "sw4b200_rhs4_corr_acc": (I, [I] + B6 + [I, c_ip] + [VP] * 9 + [D, D, VP]),
    "sw4b200_add_point_forces": (I, [I, C.c_size_t, VP, VP, I, VP, VP, D, VP]),
    "sw4b200_gather_points": (I, [I, C.c_size_t, VP, I, VP, VP, VP]),
    "sw4b200_rhs4sg_host": (I, [I] + B6 + [I, c_ip] + [c_dp] * 4 + [D] + [c_dp] * 3),
    "sw4b200_grid_create": (VP, [C.POINTER(GridDesc)]),
    "sw4b200_grid_destroy": (I, [VP]),
    "sw4b200_grid_upload": (I, [VP, C.c_char_p, c_dp]),
    "sw4b200_grid_download": (I, [VP, C.c_char_p, c_dp]),
    "sw4b200_grid_device_ptr": (VP, [VP, C.c_char_p]),
    "sw4b200_grid_array_size": (C.c_size_t, [VP, C.c_char_p]),
    "sw4b200_grid_row_pitch": (I, [VP]),
    "sw4b200_grid_set_source_points": (I, [VP, I, c_ip]),
    "sw4b200_grid_set_receiver_points": (I, [VP, I, c_ip]),
    "sw4b200_grid_predictor": (I, [VP, c_dp]),
    "sw4b200_grid_enforce_bc": (I, [VP]),
    "sw4b200_grid_enforce_cart_topo": (I, [VP, VP]),
    "sw4b200_grid_corrector": (I, [VP, c_dp]),
    "sw4b200_grid_cycle": (I, [VP]),
    "sw4b200_grid_record": (I, [VP, c_dp]),
    "sw4b200_grid_step": (I, [VP, c_dp, c_dp, c_dp]),
    "sw4b200_grid_predictor_part": (I, [VP, I, c_dp]),
    "sw4b200_grid_corrector_part": (I, [VP, I, c_dp]),
    "sw4b200_grid_set_source_series": (I, [VP, I, c_dp, c_dp]),
    "sw4b200_grid_run": (I, [VP, I, I]),
    "sw4b200_grid_fetch_records": (I, [VP, I, I, c_dp]),
    "sw4b200_grid_record_resident": (I, [VP, I]),
    "sw4b200_grid_fill_profile": (I, [VP, C.c_char_p, c_dp]),
    "sw4b200_grid_set_stream": (I, [VP, I]),
    "sw4b200_grid_pack_halo": (I, [VP, I, I, VP, VP]),
    "sw4b200_grid_unpack_halo": (I, [VP, I, I, VP, VP]),
    "sw4b200_grid_halo_doubles": (I, [VP, I]),
    "sw4b200_grid_sync": (I, [VP]),
    "sw4b200_grid_predictor_dev": (I, [VP, I, VP]),
    "sw4b200_grid_corrector_dev": (I, [VP, I, VP]),
    "sw4b200_comm_unique_id": (I, [VP]),
    "sw4b200_comm_init": (I, [I, I, VP]),
    "sw4b200_comm_finalize": (I, []),
    "sw4b200_comm_allreduce": (I, [c_dp, I, I]),
    "sw4b200_timer_start": (I, []),
    "sw4b200_timer_stop_ms": (I, [c_dp]),
    "sw4b200_grid_set_neighbours": (I, [VP, I, I]),
    "sw4b200_grid_exchange_transport": (I, [VP]),
    "sw4b200_grid_exchange_begin": (I, [VP, I]),
    "sw4b200_grid_exchange_end": (I, [VP]),
    "sw4b200_measure_fp64_peak": (I, [c_dp, c_dp]),
}

_lib = None


class Sw4b200Error(RuntimeError):
    pass


def load():
    """dlopen libsw4b200.so and type every entry point.  Raises if the library is missing:
    there is no other implementation to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise Sw4b200Error("libsw4b200.so is not built (run `python -m sw4lite_b200.build` or "
                               "__graft_entry__.build()); the product has no CPU fallback")
        lib = C.CDLL(LIBPATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise Sw4b200Error(load().sw4b200_last_error().decode())


_inited = {}


def init(device=0):
    lib = load()
    check(lib.sw4b200_init(device))
    _inited[device] = True
    return lib


def comm_init(rank, nranks):
    """communicator of the library's halo exchange (csrc/exchange.cu): rank 0 creates the NCCL unique id, torch.distributed
    (already initialised by the caller; plumbing only) hands it to the other ranks"""
    import torch
    import torch.distributed as dist
    lib = load()
    buf = (C.c_ubyte * 128)()
    if nranks > 1:
        if rank == 0:
            check(lib.sw4b200_comm_unique_id(buf))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        for n, v in enumerate(t.cpu().tolist()):
            buf[n] = v
    check(lib.sw4b200_comm_init(int(rank), int(nranks), buf))


def comm_finalize():
    check(load().sw4b200_comm_finalize())
