"""Drop-in for the reference's Cython module `navsim.util` (navsim/util.pyx).

Same names, argument meaning and error behaviour; the arithmetic runs in the
sm_100a kernels behind the C ABI (include/navsim_b200.h).  This module is the
seam `from navsim.util import sads_familiarity, downscale_chem,
fill_sensor_from` (navsim/NavBySceneFamiliarity.py:20) binds to.

Hot-path functions (device):
    sads_familiarity   util.pyx:10-25  (closure over the library; kernel :28-73)
    downscale_chem     util.pyx:91-134
    fill_sensor_from   util.pyx:137-168
Off the hot path:
    diffuse            util.pyx:186-235 (offline landscape generation; stencil kernel)
    set_HS_where_equal util.pyx:76-88, ssds :171-184 (plain host NumPy; the batched driver paints
                       chemistry on the device instead, nvb_landscape_paint)
"""
import ctypes as C
import math
import zlib

import numpy as np

from . import _cabi
from ._cabi import check, ptr

_default_engine = None
_landscape_key = None
_landscape_ref = None     # strong reference: the cached array's address cannot be handed to another array


def _engine():
    """Process-wide engine handle for the stateless single-call functions."""
    global _default_engine
    if _default_engine is None:
        h = C.c_void_p()
        check(_cabi.lib().nvb_engine_create(_cabi.default_device(), None, C.byref(h)))
        _default_engine = h
    return _default_engine


def _as_u8(name, a, ndim):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise ValueError("Buffer dtype mismatch, expected 'uint8_t' but got %r for %s" % (a.dtype.name, name))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    return a


def sads_familiarity(chem_weight=0.0):
    """util.pyx:10-25.  sads_familiarity(cw)(scenes) -> func(scene, fambuf);
    func.max_familiarity = H * W.  The library is uploaded when the closure is
    created (that is where the reference binds it, util.pyx:11-20); mutate
    `scenes` afterwards and call func.rebind() to refresh the device copy."""

    def sads_familiarity_internal(scenes):
        assert 0 <= chem_weight <= 1
        scenes_arr = _as_u8("scenes", scenes, 4)
        N, H, W, ch = scenes_arr.shape
        if ch != 3:
            raise ValueError("scenes must have 3 channels")
        maxfam = H * W
        lib = _cabi.lib()
        h = C.c_void_p()
        check(lib.nvb_engine_create(_cabi.default_device(), None, C.byref(h)))
        lut = np.tile(np.arange(256, dtype=np.uint8), (3, 1))
        check(lib.nvb_set_sensor(h, W, H, 2, 2, ptr(lut), 0))
        check(lib.nvb_set_nav_params(h, 1.0, float("inf"), 2.0, 0.8, float(chem_weight)))

        def rebind():
            check(lib.nvb_library_upload(h, ptr(np.ascontiguousarray(scenes_arr)), None, N))

        rebind()

        def func(scene, fambuf):
            q = np.ascontiguousarray(_as_u8("scene", scene, 3))
            if q.shape != (H, W, 3):
                raise ValueError("scene must have shape %r" % ((H, W, 3),))
            fb = np.asarray(fambuf)
            if fb.dtype != np.float64 or fb.ndim != 1:
                raise ValueError("Buffer dtype mismatch, expected 'float_t' for fambuf")
            out = fb if fb.flags.c_contiguous else np.empty(N, np.float64)
            check(lib.nvb_familiarity(h, ptr(q), 1, ptr(out)))
            if out is not fb:
                fb[:N] = out

        func.max_familiarity = maxfam
        func.chem_weight = float(chem_weight)
        func.rebind = rebind
        func._engine = _EngineOwner(h)
        return func

    sads_familiarity_internal.chem_weight = float(chem_weight)
    return sads_familiarity_internal


class _EngineOwner(object):
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            _cabi.lib().nvb_engine_destroy(self.h)
        except Exception:
            pass


def downscale_chem(image, factor_rows, factor_cols):
    """util.pyx:94-134: (R, C, 3) uint8 -> new (R // fr, C // fc, 3) uint8."""
    img = np.ascontiguousarray(_as_u8("image", image, 3))
    if img.shape[2] != 3:
        raise ValueError("image must have 3 channels")
    fr, fc = int(factor_rows), int(factor_cols)
    out = np.empty((img.shape[0] // fr, img.shape[1] // fc, 3), np.uint8)
    check(_cabi.lib().nvb_downscale_chem(_engine(), ptr(img), img.shape[0], img.shape[1], fr, fc,
                                         ptr(out)))
    return out


def fill_sensor_from(sensor, xpos, ypos, angle, landscape):
    """util.pyx:137-168: fills `sensor` (Hpx, Wpx, 3) in place with
    nearest-neighbour samples of `landscape` rotated about (xpos, ypos).
    Raises IndexError where the reference's bounds-checked indexing does."""
    global _landscape_key, _landscape_ref
    sensor = _as_u8("sensor", sensor, 3)
    landscape = _as_u8("landscape", landscape, 3)
    if not sensor.flags.writeable:
        raise ValueError("buffer source array is read-only")
    lib = _cabi.lib()
    h = _engine()
    # The device copy of the landscape is reused while the caller keeps passing the same
    # array: same address / shape / strides, the array object kept alive here (the reference's
    # driver makes a fresh landscape.copy() per trial, scripts/run_experiment.py:186-193, and
    # a freed copy's address is readily reused), and the same checksum of a 64 x 64 sample
    # grid.  In-place edits between the sampled pixels still need invalidate_landscape_cache().
    key = (landscape.ctypes.data, landscape.shape, landscape.strides, _sample_checksum(landscape))
    if key != _landscape_key:
        s = landscape.strides
        check(lib.nvb_set_landscape(h, ptr(landscape), landscape.shape[0], landscape.shape[1],
                                    s[0], s[1], s[2]))
        _landscape_key = key
        _landscape_ref = landscape
    rot = -(0.5 * math.pi - float(angle))       # util.pyx:143
    out = sensor if sensor.flags.c_contiguous else np.empty(sensor.shape, np.uint8)
    rc = check(lib.nvb_fill_sensor(h, ptr(out), sensor.shape[0], sensor.shape[1], float(xpos),
                                   float(ypos), math.cos(rot), math.sin(rot)))
    if rc == _cabi.INDEX_ERROR:
        raise IndexError("Index out of bounds (axis 0)")
    if out is not sensor:
        sensor[...] = out


def _sample_checksum(a):
    r, c = max(1, a.shape[0] // 64), max(1, a.shape[1] // 64)
    return zlib.crc32(a[::r, ::c].tobytes())


def invalidate_landscape_cache():
    """Forget the device copy made by fill_sensor_from (call after mutating the
    landscape array in place)."""
    global _landscape_key, _landscape_ref
    _landscape_key = None
    _landscape_ref = None


# ---- off the hot path: host NumPy -------------------------------------------
def set_HS_where_equal(labels, image, H, S):
    """util.pyx:76-88: paint hue/saturation per labelled grain (label 0 = none)."""
    labels = np.asarray(labels)
    m = labels > 0
    idx = labels[m] - 1
    image[:, :, 0][m] = np.asarray(H)[idx]
    image[:, :, 1][m] = np.asarray(S)[idx]


def ssds(a_np, b_np):
    """util.pyx:171-184: sum of squared differences of two float64 2-D arrays."""
    d = np.asarray(a_np, dtype=np.float64) - np.asarray(b_np, dtype=np.float64)
    return float(np.sum(d * d))


def diffuse(initial_condition, nstep, c=1.0, delta_t_factor=0.5):
    """util.pyx:186-235: explicit 2-D heat equation, periodic boundaries -- `nstep` stencil
    passes on the device (nvb_diffuse), bit-identical to the Cython loop; same short circuit,
    same sanity assertions."""
    if nstep == 0:                                                       # :190-191
        return initial_condition
    initial_condition = np.asarray(initial_condition)
    mat = np.ascontiguousarray(initial_condition, dtype=np.float64)     # :193
    assert initial_condition.shape[0] == initial_condition.shape[1]     # :199
    side = mat.shape[0]
    delta_s = 1.0 / (side + 1)                                           # :202
    delta_t = delta_t_factor * ((delta_s) ** 2 / (2 * c))                # :203
    multiplier = c * (delta_t / (delta_s * delta_s))                     # :205
    out = np.empty_like(mat)
    check(_cabi.lib().nvb_diffuse(_engine(), ptr(mat), side, int(nstep), float(multiplier), ptr(out)))
    assert np.sum(out) - np.sum(initial_condition) < 0.0000001           # :229-232
    assert np.max(out) <= np.max(initial_condition)
    assert np.min(out) >= np.min(initial_condition)
    assert not np.any(np.isnan(out))
    return out


def diffuse_host(initial_condition, nstep, c=1.0, delta_t_factor=0.5):
    """The same recurrence in NumPy (tests)."""
    if nstep == 0:
        return initial_condition
    mat = np.array(initial_condition, dtype=np.float64, copy=True)
    side = mat.shape[0]
    delta_s = 1.0 / (side + 1)
    delta_t = delta_t_factor * (delta_s ** 2 / (2 * c))
    mult = c * (delta_t / (delta_s * delta_s))
    for _ in range(int(nstep)):
        mat = mat + mult * (np.roll(mat, -1, 0) + np.roll(mat, 1, 0) - 4 * mat
                            + np.roll(mat, -1, 1) + np.roll(mat, 1, 1))
    return mat
