"""ctypes binding of libdff_b200.so (include/dff_b200.h).  No CPU fallback: if the library is
missing or no B200 is visible, the calls raise -- nothing is computed on the host."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFF_LIB_PATH") or os.path.join(_HERE, "libdff_b200.so")

DFF_MD_BAOAB = 0
DFF_MD_BROWNIAN = 1
FLAG_CLAMPED, FLAG_CENTER, FLAG_NONFINITE = 1, 2, 4
NUM_GLOBAL_WEIGHTS, NUM_LAYER_WEIGHTS = 6, 18


class DffError(RuntimeError):
    pass


class MdParams(C.Structure):
    _fields_ = [("integrator", C.c_int), ("t_norm", C.c_float), ("force_scale", C.c_float), ("dt", C.c_float),
                ("vscale", C.c_float), ("noisescale", C.c_float), ("beta", C.c_float), ("dtau", C.c_float)]


class ModelOpts(C.Structure):
    _fields_ = [("conservative", C.c_int), ("use_intrinsic_coords", C.c_int), ("use_distances", C.c_int), ("use_abs_coords", C.c_int)]


_lib = None

# every symbol include/dff_b200.h declares: name -> (restype, argtypes)
_fp, _u32p, _vp = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_void_p
SYMBOLS = {
    "dff_last_error": (C.c_char_p, []),
    "dff_version": (C.c_int, []),
    "dff_device_count": (C.c_int, []),
    "dff_model_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.c_int, C.c_int]),
    "dff_model_create_ex": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "dff_model_create_v2": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.c_int, C.c_int, C.POINTER(ModelOpts)]),
    "dff_model_destroy": (None, [_vp]),
    "dff_model_num_beads": (C.c_int, [_vp]),
    "dff_model_hidden": (C.c_int, [_vp]),
    "dff_model_layers": (C.c_int, [_vp]),
    "dff_model_device": (C.c_int, [_vp]),
    "dff_model_launch_count": (C.c_int64, [_vp]),
    "dff_model_last_config": (C.c_char_p, [_vp]),
    "dff_model_flops_per_sample": (C.c_double, [_vp]),
    "dff_score_dev": (C.c_int, [_vp, _vp, C.c_float, C.c_int, _vp, _vp, _vp]),
    "dff_score_dev_t": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "dff_score_host": (C.c_int, [_vp, _vp, C.c_float, C.c_int, _vp, _vp]),
    "dff_ddpm_steps_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), _vp, C.c_uint64,
                                     C.c_uint64, _vp, _vp]),
    "dff_langevin_steps_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(MdParams), _vp, _vp, C.c_uint64,
                                         C.c_uint64, C.c_int, _vp, _vp, _vp, _vp]),
    "dff_ddpm_sample_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp), C.c_uint64, _vp]),
    "dff_langevin_run_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(MdParams), _vp, C.c_uint64, C.c_int,
                                        _vp, _vp, _vp]),
    "dff_pwd_num_pairs": (C.c_int, [C.c_int, C.c_int]),
    "dff_pwd_max_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "dff_pwd_hist_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_float, _vp, C.c_int, _vp, _vp]),
    "dff_contacts_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_float, _vp, C.c_int, _vp, _vp, _vp]),
    "dff_dihedrals_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "dff_rmsd_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "dff_debug_read_stash": (C.c_int64, [_vp, _vp, C.c_int64]),
    "dff_debug_tc_gemm": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "dff_debug_stash_layout": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


def lib():
    """Load (once) and return the shared library; raises DffError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DffError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                           "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise DffError(f"libdff_b200 error {rc}: {lib().dff_last_error().decode()}")
