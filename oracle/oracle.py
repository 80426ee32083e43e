"""ctypes front end of the oracle's C restatement (oracle/navsim_oracle.c).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

OK, REACHED_END, TOO_FAR, OUT_OF_BOUNDS, INDEX_ERROR = 0, 1, -1, -2, -3

_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)


class _World(C.Structure):
    _fields_ = [
        ("land", C.c_void_p), ("rows", C.c_long), ("cols", C.c_long),
        ("s_row", C.c_long), ("s_col", C.c_long), ("s_chan", C.c_long),
        ("W", C.c_long), ("H", C.c_long), ("pw", C.c_long), ("ph", C.c_long),
        ("lut", (C.c_uint8 * 256) * 3), ("mask_middle_n", C.c_long),
        ("A", C.c_long), ("offsets", C.c_void_p),
        ("step_size", C.c_double), ("max_dist", C.c_double),
        ("threshold_factor", C.c_double), ("coverage_factor", C.c_double),
        ("chem_weight", C.c_double),
        ("N", C.c_long), ("scenes", C.c_void_p), ("path", C.c_void_p),
    ]


class _Agent(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("angle", C.c_double),
                ("navigated_for_frames", C.c_long), ("nav_err", C.c_double),
                ("n_nav_err", C.c_long), ("coverage", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.nvo_fill_sensor.restype = C.c_int
        _lib.nvo_get_sensor_mat.restype = C.c_int
        _lib.nvo_train_from_path.restype = C.c_int
        _lib.nvo_step_forward.restype = C.c_int
        _lib.nvo_run.restype = C.c_int
    return _lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def quant_lut(nlevels):
    out = np.empty(256, np.uint8)
    lib().nvo_quant_lut(C.c_int(int(nlevels)), _ptr(out))
    return out


def fill_sensor(sensor, x, y, angle, landscape):
    """util.pyx:137-168; sensor (Hpx, Wpx, 3) uint8 C-contiguous, in place."""
    assert sensor.flags.c_contiguous and sensor.dtype == np.uint8
    Hpx, Wpx, _ = sensor.shape
    s = landscape.strides
    return lib().nvo_fill_sensor(_ptr(sensor), C.c_long(Hpx), C.c_long(Wpx), C.c_double(x),
                                 C.c_double(y), C.c_double(angle), _ptr(landscape),
                                 C.c_long(landscape.shape[0]), C.c_long(landscape.shape[1]),
                                 C.c_long(s[0]), C.c_long(s[1]), C.c_long(s[2]))


def downscale_chem(image, fr, fc):
    image = np.ascontiguousarray(image)
    R, Cc, _ = image.shape
    out = np.empty((R // fr, Cc // fc, 3), np.uint8)
    lib().nvo_downscale_chem(_ptr(image), C.c_long(R), C.c_long(Cc), C.c_long(fr), C.c_long(fc),
                             _ptr(out))
    return out


def sads_hsv(scenes, scene, cw=0.0):
    scenes = np.ascontiguousarray(scenes)
    scene = np.ascontiguousarray(scene)
    N, H, W, _ = scenes.shape
    fam = np.empty(N, np.float64)
    lib().nvo_sads_hsv(_ptr(scenes), C.c_long(N), C.c_long(H), C.c_long(W), _ptr(scene),
                       _ptr(fam), C.c_double(cw))
    return fam


def sad_int(scenes, scene):
    scenes = np.ascontiguousarray(scenes)
    scene = np.ascontiguousarray(scene)
    N = scenes.shape[0]
    P = scenes.shape[1] * scenes.shape[2]
    xt = np.empty(N, np.uint32)
    vt = np.empty(N, np.uint32)
    lib().nvo_sad_int(_ptr(scenes), C.c_long(N), C.c_long(P), _ptr(scene), _ptr(xt), _ptr(vt))
    return xt, vt


class World:
    """One landscape + sensor + saccade + library: the state the reference keeps
    in a NavBySceneFamiliarity instance (NavBySceneFamiliarity.py:59-116)."""

    def __init__(self, landscape, sensor_dimensions, step_size, n_test_angles=60,
                 sensor_pixel_dimensions=(1, 1), max_distance_to_training_path=np.inf,
                 n_sensor_levels=5, mask_middle_n=0, threshold_factor=2.,
                 coverage_threshold_factor=0.8, saccade_degrees=180., chem_weight=0.0):
        assert landscape.dtype == np.uint8 and landscape.ndim == 3 and landscape.shape[2] == 3
        self.landscape = landscape
        self.W, self.H = int(sensor_dimensions[0]), int(sensor_dimensions[1])
        self.pw, self.ph = int(sensor_pixel_dimensions[0]), int(sensor_pixel_dimensions[1])
        assert (self.W * self.pw) % 2 == 0 and (self.H * self.ph) % 2 == 0
        if not isinstance(n_sensor_levels, tuple):
            n_sensor_levels = (256, 256, n_sensor_levels)
        self.n_sensor_levels = n_sensor_levels
        sd2 = saccade_degrees / 2
        self.offsets = np.linspace(-(np.pi * sd2 / 180.), np.pi * sd2 / 180., n_test_angles)
        self.A = int(n_test_angles)
        self.scenes = None
        self.path = None
        w = self._w = _World()
        w.land = landscape.ctypes.data
        w.rows, w.cols = landscape.shape[0], landscape.shape[1]
        w.s_row, w.s_col, w.s_chan = landscape.strides
        w.W, w.H, w.pw, w.ph = self.W, self.H, self.pw, self.ph
        for ch in range(3):
            lut = quant_lut(n_sensor_levels[ch])
            for k in range(256):
                w.lut[ch][k] = int(lut[k])
        w.mask_middle_n = int(mask_middle_n)
        w.A = self.A
        w.offsets = self.offsets.ctypes.data
        w.step_size = float(step_size)
        w.max_dist = float(max_distance_to_training_path)
        w.threshold_factor = float(threshold_factor)
        w.coverage_factor = float(coverage_threshold_factor)
        w.chem_weight = float(chem_weight)
        w.N = 0
        self.step_size = float(step_size)

    def _scratch(self):
        return np.empty(self.H * self.ph * self.W * self.pw * 3 + self.H * self.W * 3, np.uint8)

    def get_sensor_mat(self, position, angle):
        out = np.empty((self.H, self.W, 3), np.uint8)
        sc = self._scratch()
        rc = lib().nvo_get_sensor_mat(C.byref(self._w), C.c_double(position[0]),
                                      C.c_double(position[1]), C.c_double(angle), _ptr(out),
                                      _ptr(sc))
        return rc, out

    def set_library(self, scenes, path):
        self.scenes = np.ascontiguousarray(scenes, dtype=np.uint8)
        self.path = np.ascontiguousarray(path, dtype=np.float64)
        self._w.N = self.scenes.shape[0]
        self._w.scenes = self.scenes.ctypes.data
        self._w.path = self.path.ctypes.data

    def train_from_path(self, points):
        points = np.ascontiguousarray(points, dtype=np.float64)
        N = len(points)
        scenes = np.empty((N, self.H, self.W, 3), np.uint8)
        sc = self._scratch()
        bad = C.c_long(-1)
        rc = lib().nvo_train_from_path(C.byref(self._w), _ptr(points), C.c_long(N), _ptr(scenes),
                                       _ptr(sc), C.byref(bad))
        if rc:
            return rc, bad.value
        self.set_library(scenes, points)
        return 0, -1

    def new_agent(self, x, y, angle):
        a = _Agent()
        a.x, a.y, a.angle = float(x), float(y), float(angle)
        a.navigated_for_frames = 0
        a.nav_err = 0.0
        a.n_nav_err = 0
        cov = np.zeros(self._w.N, np.uint8)
        a.coverage = cov.ctypes.data
        a._cov = cov
        return a

    def step_forward(self, agent, fake=False, want_scene_fam=False):
        """Returns (status, best_idx, angle_familiarity[A], scene_familiarity|None)."""
        af = np.empty(self.A, np.float64)
        sf = np.empty(self._w.N, np.float64) if want_scene_fam else None
        best = C.c_long(-1)
        sfam = C.c_double(0)
        sc = self._scratch()
        tmp = np.empty(self._w.N, np.float64)
        rc = lib().nvo_step_forward(C.byref(self._w), C.byref(agent), C.c_int(int(fake)), _ptr(af),
                                    _ptr(sf) if sf is not None else None, C.byref(best),
                                    C.byref(sfam), _ptr(sc), _ptr(tmp))
        return rc, best.value, af, sf

    def run(self, agent, frames, log_afam=False):
        """Returns dict(status, completed, best_idx[frames], pos[frames,3], afam)."""
        best = np.full(frames, -1, np.int32)
        pos = np.full((frames, 3), np.nan)
        afam = np.full((frames, self.A), np.nan) if log_afam else None
        done = C.c_long(0)
        rc = lib().nvo_run(C.byref(self._w), C.byref(agent), C.c_long(frames), C.byref(done),
                           _ptr(best), _ptr(pos), _ptr(afam) if afam is not None else None)
        return dict(status=rc, completed=done.value, best_idx=best, pos=pos, afam=afam)

    def run_batch(self, poses, frames, log_best=False):
        poses = np.array(poses, dtype=np.float64, copy=True, order="C")
        B = len(poses)
        N = self._w.N
        status = np.zeros(B, np.int32)
        completed = np.zeros(B, np.int64)
        nav_err = np.zeros(B, np.float64)
        n_nav = np.zeros(B, np.int64)
        cov = np.zeros((B, N), np.uint8)
        best = np.full((B, frames), -1, np.int32) if log_best else None
        lib().nvo_run_batch(C.byref(self._w), C.c_long(B), _ptr(poses), C.c_long(frames),
                            _ptr(status), _ptr(completed), _ptr(nav_err), _ptr(n_nav), _ptr(cov),
                            _ptr(best) if best is not None else None)
        return dict(poses=poses, status=status, completed=completed, nav_err=nav_err,
                    n_nav_err=n_nav, coverage=cov, best_idx=best)
