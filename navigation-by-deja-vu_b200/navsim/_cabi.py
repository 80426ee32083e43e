"""ctypes binding of include/navsim_b200.h (the C-ABI shared library).

The library is built in-tree by __graft_entry__.build() (nvcc, sm_100a) at
navigation-by-deja-vu_b200/lib/libnavsim_b200.so.  There is no CPU fallback:
if the library is missing, or no B200 is visible, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libnavsim_b200.so")

OK, REACHED_END, TOO_FAR, OUT_OF_BOUNDS, INDEX_ERROR = 0, 1, -1, -2, -3
E_INVALID, E_CUDA, E_NO_DEVICE = -100, -101, -102
PTR_KEYS, PTR_TIE, PTR_POSES = 0, 1, 2

_vp, _i, _d, _i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
_sz = C.c_ssize_t

# name -> (restype, argtypes); mirrors include/navsim_b200.h one to one
SIGNATURES = {
    "nvb_engine_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "nvb_engine_destroy": (None, [_vp]),
    "nvb_last_error": (C.c_char_p, []),
    "nvb_version": (C.c_char_p, []),
    "nvb_sync": (_i, [_vp]),
    "nvb_stream_handle": (_vp, [_vp]),
    "nvb_set_landscape": (_i, [_vp, _vp, _i, _i, _sz, _sz, _sz]),
    "nvb_landscape_label_grains": (_i, [_vp, _i, _i, C.POINTER(_i64)]),
    "nvb_landscape_grains_get": (_i, [_vp, _vp, _vp]),
    "nvb_landscape_paint": (_i, [_vp, _vp, _vp, _i64]),
    "nvb_landscape_flip": (_i, [_vp, _i, _i]),
    "nvb_diffuse": (_i, [_vp, _vp, _i64, _i64, _d, _vp]),
    "nvb_landscape_download": (_i, [_vp, _vp]),
    "nvb_set_sensor": (_i, [_vp, _i, _i, _i, _i, _vp, _i]),
    "nvb_set_saccade": (_i, [_vp, _i, _vp]),
    "nvb_set_nav_params": (_i, [_vp, _d, _d, _d, _d, _d]),
    "nvb_fill_sensor": (_i, [_vp, _vp, _i, _i, _d, _d, _d, _d]),
    "nvb_downscale_chem": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "nvb_glimpse_batch": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "nvb_library_build": (_i, [_vp, _vp, _vp, _vp, _i, C.POINTER(_i)]),
    "nvb_library_upload": (_i, [_vp, _vp, _vp, _i]),
    "nvb_set_training_path": (_i, [_vp, _vp, _i]),
    "nvb_library_download": (_i, [_vp, _vp]),
    "nvb_library_set_shard": (_i, [_vp, _i64, _i64]),
    "nvb_familiarity": (_i, [_vp, _vp, _i, _vp]),
    "nvb_familiarity_min": (_i, [_vp, _vp, _i, _vp, _vp]),
    "nvb_agents_set": (_i, [_vp, _vp, _vp, _i]),
    "nvb_agents_step": (_i, [_vp, _i, _i, _i]),
    "nvb_agents_rewind": (_i, [_vp]),
    "nvb_agents_step_io": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "nvb_set_options": (_i, [_vp, _i, _i]),
    "nvb_kernel_time_ms": (_d, [_vp, C.POINTER(_i64)]),
    "nvb_agents_get": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nvb_agents_log": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "nvb_agents_steps_done": (_i, [_vp]),
    "nvb_agents_phase": (_i, [_vp, _i, _i, _i]),
    "nvb_p2p_export": (_i, [_vp, _vp]),
    "nvb_p2p_attach": (_i, [_vp, _i, _i, _vp]),
    "nvb_p2p_error": (_i, [_vp]),
    "nvb_device_ptr": (_vp, [_vp, _i]),
    "nvb_debug_sincos": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "nvb_debug_step_clocks": (_i, [_vp, _vp]),
    "nvb_debug_timeline": (_i, [_vp, _i, _vp]),
    "nvb_set_distance_kernel": (_i, [_vp, _i]),
    "nvb_distance_kernel": (_i, [_vp]),
    "nvb_launch_count": (_i64, [_vp]),
    "nvb_probe_sad_peak": (_d, [_vp, _i]),
    "nvb_probe_mma_peak": (_d, [_vp, _i]),
    "nvb_tc_planes": (_i, [_vp]),
    "nvb_time_distance_kernel": (_d, [_vp, _i]),
}

_lib = None


class NavsimB200Error(RuntimeError):
    pass


def lib():
    """The loaded C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NavsimB200Error(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().nvb_last_error().decode()


def check(rc):
    """Raises for API-level failures; per-agent / per-pose status codes pass through."""
    if rc <= E_INVALID:
        msg = last_error()
        if rc == E_INVALID:
            raise ValueError(msg)
        raise NavsimB200Error(msg)
    return rc


def ptr(a):
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)


def default_device():
    for key in ("NAVSIM_B200_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


def quant_lut(nlevels):
    """256-entry table of the float32 level rounding in get_sensor_mat
    (navsim/NavBySceneFamiliarity.py:178-186), built with the same NumPy ops."""
    buf = np.arange(256, dtype=np.uint8).astype(np.float32)
    buf /= 255
    buf *= (nlevels - 1)
    np.rint(buf, out=buf)
    buf /= (nlevels - 1)
    buf *= 255
    out = np.empty(256, np.uint8)
    out[:] = buf
    return out
