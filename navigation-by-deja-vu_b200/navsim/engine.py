"""NavEngine: the batched, device-resident form of NavBySceneFamiliarity.

One NavEngine is one world (landscape + sensor + heading sweep + library) on
one B200.  Where the reference object (navsim/NavBySceneFamiliarity.py:57-329)
advances one agent by one step per Python call, the engine advances a whole
batch of agents by many steps per call without host interaction; the
single-agent class in NavBySceneFamiliarity.py replays its log.

Host code is Python; every array handed to the C ABI is a NumPy host array
(torch is used by callers only for pinned buffers, streams and NCCL).
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import check, ptr


class NavEngine(object):

    def __init__(self,
                 landscape,
                 sensor_dimensions,
                 step_size,
                 n_test_angles=60,
                 sensor_pixel_dimensions=[1, 1],
                 max_distance_to_training_path=np.inf,
                 n_sensor_levels=5,
                 mask_middle_n=0,
                 threshold_factor=2.,
                 coverage_threshold_factor=0.8,
                 saccade_degrees=180.,
                 chem_weight=0.0,
                 device=None,
                 stream=None):
        self._lib = _cabi.lib()
        self._h = C.c_void_p()
        dev = _cabi.default_device() if device is None else int(device)
        # stream: a cudaStream_t handle (e.g. torch.cuda.Stream().cuda_stream).  None:
        # the engine makes its own stream.  0 is CUDA's legacy default stream, spelled
        # cudaStreamLegacy (0x1) for the C ABI, where NULL means "make your own".
        if stream is not None and int(stream) == 0:
            stream = 1
        check(self._lib.nvb_engine_create(dev, C.c_void_p(int(stream)) if stream is not None else None,
                                          C.byref(self._h)))
        self.device = dev
        self.landscape = None
        self.set_landscape(landscape)
        self.set_world(sensor_dimensions, step_size, n_test_angles=n_test_angles,
                       sensor_pixel_dimensions=sensor_pixel_dimensions,
                       max_distance_to_training_path=max_distance_to_training_path,
                       n_sensor_levels=n_sensor_levels, mask_middle_n=mask_middle_n,
                       threshold_factor=threshold_factor, coverage_threshold_factor=coverage_threshold_factor,
                       saccade_degrees=saccade_degrees, chem_weight=chem_weight)

    def set_landscape(self, landscape):
        """Uploads another landscape (any strides: flipped views are fine).  The sensor stays;
        library and agents must be set again."""
        landscape = np.asarray(landscape)
        if landscape.dtype != np.uint8 or landscape.ndim != 3 or landscape.shape[2] != 3:
            raise ValueError("landscape must be a (rows, cols, 3) uint8 HSV array")
        self.landscape = landscape
        s = landscape.strides
        check(self._lib.nvb_set_landscape(self._h, ptr(landscape), landscape.shape[0],
                                          landscape.shape[1], s[0], s[1], s[2]))
        self.training_path = None
        self.n_views = 0
        self.n_agents = 0
        self._familiar_scenes = None

    # ---- landscape preparation on the device (scripts/run_experiment.py:160-199) ----------
    def label_grains(self, threshold=200, min_chem_grain_diameter=2):
        """Threshold + modal filter + grain labelling of the landscape on the device, as make_nsf
        does them (run_experiment.py:169-180).  Returns the per-grain pixel counts (int32)."""
        w = min_chem_grain_diameter // 2                   # :170-174
        if not w == 0:
            if w % 2 == 0:
                w -= 1
            w = int(w)
        n = C.c_int64(0)
        check(self._lib.nvb_landscape_label_grains(self._h, int(threshold), int(w), C.byref(n)))
        areas = np.zeros(int(n.value), np.int32)
        check(self._lib.nvb_landscape_grains_get(self._h, ptr(areas) if len(areas) else None, None))
        return areas

    def grain_labels(self):
        out = np.empty(self.landscape.shape[:2], np.int64)
        check(self._lib.nvb_landscape_grains_get(self._h, None, ptr(out)))
        return out

    def add_chemistry(self, areas, n_chemicals=2, min_grain_diameter=2, concentration_range=(127, 128), rng=None):
        """add_chemistry (run_experiment.py:126-142) on the device copy of the landscape: per-grain
        hue / saturation tables drawn on the host (same RNG calls), painted by a kernel."""
        rng = np.random.default_rng() if rng is None else rng
        n = len(areas)
        chems = rng.integers(n_chemicals, size=n, dtype=np.uint8) * np.uint8(255 // n_chemicals)
        sats = rng.integers(concentration_range[0], concentration_range[1], size=n, dtype=np.uint8)
        sats[np.sqrt(4.0 * areas / np.pi) < min_grain_diameter] = 0      # regionprops equivalent_diameter
        chems, sats = np.ascontiguousarray(chems, np.uint8), np.ascontiguousarray(sats, np.uint8)
        check(self._lib.nvb_landscape_paint(self._h, ptr(chems), ptr(sats), n))
        self._after_landscape_edit()
        return chems, sats

    def flip_landscape(self, vertical=False, horizontal=False):
        """landscape[::-1] / [:, ::-1] on the device (run_experiment.py:196-199)."""
        check(self._lib.nvb_landscape_flip(self._h, int(bool(vertical)), int(bool(horizontal))))
        self._after_landscape_edit()

    def download_landscape(self):
        out = np.empty(self.landscape.shape[:2] + (3,), np.uint8)
        check(self._lib.nvb_landscape_download(self._h, ptr(out)))
        return out

    def _after_landscape_edit(self):
        self.training_path = None
        self.n_views = 0
        self.n_agents = 0
        self._familiar_scenes = None
        self.landscape = self.download_landscape() if False else self.landscape   # host copy is the caller's; shape only

    def set_world(self, sensor_dimensions, step_size, n_test_angles=60, sensor_pixel_dimensions=[1, 1],
                  max_distance_to_training_path=np.inf, n_sensor_levels=5, mask_middle_n=0,
                  threshold_factor=2., coverage_threshold_factor=0.8, saccade_degrees=180., chem_weight=0.0):
        """(Re)configures sensor, heading sweep and navigation parameters on the landscape the
        engine already holds -- what the reference's constructor does
        (NavBySceneFamiliarity.py:59-116) without a new engine, stream or landscape upload:
        a parameter sweep walks through its worlds on one engine per device."""
        self.sensor_dimensions = np.asarray(sensor_dimensions)
        self.sensor_pixel_dimensions = np.asarray(sensor_pixel_dimensions)
        footprint = self.sensor_dimensions * self.sensor_pixel_dimensions
        assert np.all(footprint % 2 == 0)           # NavBySceneFamiliarity.py:93
        self._sensor_r = np.max(footprint / 2)      # :94
        if not isinstance(n_sensor_levels, tuple):  # :100-104
            n_sensor_levels = (256, 256, n_sensor_levels)
        assert len(n_sensor_levels) == 3
        assert all(2 <= l <= 256 for l in n_sensor_levels)
        self.n_sensor_levels = n_sensor_levels
        self.n_test_angles = int(n_test_angles)
        self.saccade_degrees = saccade_degrees
        sd2 = saccade_degrees / 2                   # :86-88
        self.angle_offsets = np.linspace(-(np.pi * sd2 / 180.), np.pi * sd2 / 180., self.n_test_angles)
        self.step_size = step_size
        self.max_distance_to_training_path = max_distance_to_training_path
        self.mask_middle_n = mask_middle_n
        self.threshold_factor = threshold_factor
        self.coverage_threshold_factor = coverage_threshold_factor
        assert 0 <= chem_weight <= 1                # util.pyx:12
        self.chem_weight = float(chem_weight)
        self.training_path = None
        self.n_views = 0
        self.n_agents = 0
        self._familiar_scenes = None
        lut = np.ascontiguousarray(np.stack([_cabi.quant_lut(n) for n in n_sensor_levels]))
        check(self._lib.nvb_set_sensor(self._h, int(self.sensor_dimensions[0]),
                                       int(self.sensor_dimensions[1]),
                                       int(self.sensor_pixel_dimensions[0]),
                                       int(self.sensor_pixel_dimensions[1]), ptr(lut),
                                       int(mask_middle_n)))
        offs = np.ascontiguousarray(self.angle_offsets, dtype=np.float64)
        check(self._lib.nvb_set_saccade(self._h, self.n_test_angles, ptr(offs)))
        self._push_nav_params()

    def _push_nav_params(self):
        check(self._lib.nvb_set_nav_params(self._h, float(self.step_size),
                                           float(self.max_distance_to_training_path),
                                           float(self.threshold_factor),
                                           float(self.coverage_threshold_factor),
                                           float(self.chem_weight)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.nvb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sensor_shape(self):
        """(H, W, 3): shape of one glimpse (get_sensor_mat's return value)."""
        return (int(self.sensor_dimensions[1]), int(self.sensor_dimensions[0]), 3)

    # ---- glimpses ---------------------------------------------------------
    @staticmethod
    def _rot_cs(angles):
        """cos/sin of -(pi/2 - angle) with the host libm (util.pyx:143-145)."""
        rot = -(0.5 * np.pi - np.asarray(angles, dtype=np.float64))
        return np.ascontiguousarray(np.stack([np.cos(rot), np.sin(rot)], axis=-1))

    def get_sensor_mats(self, poses, host_trig=True):
        """get_sensor_mat (NavBySceneFamiliarity.py:151-192) for G poses
        [(x, y, angle)].  Returns (glimpses (G, H, W, 3) uint8, status (G,)).
        status: 0, -2 (out of landscape bounds) or -3 (IndexError in the gather)."""
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        G = len(poses)
        out = np.zeros((G,) + self.sensor_shape, np.uint8)
        status = np.zeros(G, np.int32)
        cs = self._rot_cs(poses[:, 2]) if host_trig else None
        check(self._lib.nvb_glimpse_batch(self._h, ptr(poses), ptr(cs), G, ptr(out), ptr(status)))
        return out, status

    # ---- library ----------------------------------------------------------
    def train_from_path(self, points):
        """Library build (NavBySceneFamiliarity.py:118-140).  Returns (status,
        bad_index): (0, -1) on success, else the failing point's status."""
        points = np.ascontiguousarray(points, dtype=np.float64)
        N = len(points)
        d = points[1:] - points[0:-1]
        angles = np.empty(N, np.float64)
        angles[:-1] = np.arctan2(d[:, 1], d[:, 0])   # :124-126
        angles[-1] = angles[-2]                      # :132
        cs = self._rot_cs(angles)
        bad = C.c_int(-1)
        rc = check(self._lib.nvb_library_build(self._h, ptr(points), ptr(angles), ptr(cs), N,
                                               C.byref(bad)))
        if rc != 0:
            return rc, bad.value
        self.training_path = points
        self.training_path_length = np.sum(np.linalg.norm(d, axis=1))   # :125
        self.n_views = N
        self._familiar_scenes = None
        return 0, -1

    def set_library(self, scenes, path=None):
        """Binds an existing library (what util.pyx:11-20 captures)."""
        scenes = np.ascontiguousarray(scenes, dtype=np.uint8)
        if scenes.shape[1:] != self.sensor_shape:
            raise ValueError("scenes must be (N,) + %r" % (self.sensor_shape,))
        if path is not None:
            path = np.ascontiguousarray(path, dtype=np.float64)
            assert len(path) == len(scenes)
        check(self._lib.nvb_library_upload(self._h, ptr(scenes), ptr(path), len(scenes)))
        self.training_path = path
        self.n_views = len(scenes)
        self._familiar_scenes = None

    def set_library_shard(self, scenes, view_offset, n_total, path):
        """One rank's slice [view_offset, view_offset+len(scenes)) of a library
        of n_total views; `path` is the WHOLE training path (n_total points)."""
        scenes = np.ascontiguousarray(scenes, dtype=np.uint8)
        path = np.ascontiguousarray(path, dtype=np.float64)
        assert len(path) == n_total
        # the path rides along with the local scenes, then is replaced by the whole one
        check(self._lib.nvb_library_upload(self._h, ptr(scenes), None, len(scenes)))
        check(self._lib.nvb_library_set_shard(self._h, int(view_offset), int(n_total)))
        self._set_path_only(path)
        self.training_path = path
        self.n_views = len(scenes)
        self._familiar_scenes = None

    def p2p_attach(self, rank, world_size, group=None):
        """View shards over NVLink peer memory: exchanges this rank's CUDA IPC handle with
        its peers (torch.distributed, any backend) and attaches them.  Afterwards step()
        runs the sharded sequence on the device without NCCL or host round trips.  Call
        after set_library_shard() and set_agents(); every rank must then call step() with
        the same arguments."""
        import torch.distributed as dist
        handle = (C.c_ubyte * 64)()
        check(self._lib.nvb_p2p_export(self._h, C.cast(handle, C.c_void_p)))
        gathered = [None] * world_size
        dist.all_gather_object(gathered, bytes(handle), group=group)
        blob = (C.c_ubyte * (64 * world_size)).from_buffer_copy(b"".join(gathered))
        check(self._lib.nvb_p2p_attach(self._h, int(rank), int(world_size), C.cast(blob, C.c_void_p)))
        dist.barrier(group=group)   # every rank has mapped every area before anybody steps

    def p2p_error(self):
        return int(self._lib.nvb_p2p_error(self._h))

    def _set_path_only(self, path):
        check(self._lib.nvb_set_training_path(self._h, ptr(path), len(path)))

    @property
    def familiar_scenes(self):
        """(N, H, W, 3) uint8, NavBySceneFamiliarity.py:122."""
        if self._familiar_scenes is None and self.n_views > 0:
            out = np.empty((self.n_views,) + self.sensor_shape, np.uint8)
            check(self._lib.nvb_library_download(self._h, ptr(out)))
            self._familiar_scenes = out
        return self._familiar_scenes

    # ---- distance ---------------------------------------------------------
    def familiarity(self, scenes_q):
        """sads_hsv_metric (util.pyx:28-73) for G query scenes: fam (G, N) float64,
        bit-identical operation order."""
        scenes_q = np.ascontiguousarray(scenes_q, dtype=np.uint8).reshape((-1,) + self.sensor_shape)
        fam = np.empty((len(scenes_q), self.n_views), np.float64)
        check(self._lib.nvb_familiarity(self._h, ptr(scenes_q), len(scenes_q), ptr(fam)))
        return fam

    def familiarity_min(self, scenes_q):
        """The hot kernel on host inputs: (min difference (G,), view index (G,))."""
        scenes_q = np.ascontiguousarray(scenes_q, dtype=np.uint8).reshape((-1,) + self.sensor_shape)
        G = len(scenes_q)
        md = np.empty(G, np.float64)
        vi = np.empty(G, np.int64)
        check(self._lib.nvb_familiarity_min(self._h, ptr(scenes_q), G, ptr(md), ptr(vi)))
        return md, vi

    # ---- resident stepping loop --------------------------------------------
    def set_agents(self, poses, frame_budget=None):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        fb = None
        if frame_budget is not None:
            fb = np.ascontiguousarray(np.broadcast_to(np.asarray(frame_budget, dtype=np.int32),
                                                      (len(poses),)))
        self._push_nav_params()
        check(self._lib.nvb_agents_set(self._h, ptr(poses), ptr(fb), len(poses)))
        self.n_agents = len(poses)

    def step(self, nsteps=1, fake=False, log_afam=False):
        """Queues nsteps step-batches (asynchronous)."""
        check(self._lib.nvb_agents_step(self._h, int(nsteps), int(bool(fake)), int(bool(log_afam))))

    def rewind(self):
        """Back to the start poses of the last set_agents(), device side only."""
        check(self._lib.nvb_agents_rewind(self._h))

    def step_io(self, poses_in, nsteps, best_idx, poses_out, step_fam):
        """Host poses in -> nsteps step-batches -> last step's results out, one
        synchronisation.  Arrays are caller-owned (pinned torch tensors' numpy views
        in bench.py); poses_in may be None."""
        check(self._lib.nvb_agents_step_io(self._h, ptr(poses_in), int(nsteps), ptr(best_idx),
                                           ptr(poses_out), ptr(step_fam)))

    def bind_step_io(self, poses_in, best_idx, poses_out, step_fam, nsteps=1):
        """step_io for a host-driven loop that reuses its buffers: resolves the four array
        pointers once and returns a zero-argument callable doing one call each time."""
        fn, h = self._lib.nvb_agents_step_io, self._h
        args = (ptr(poses_in), int(nsteps), ptr(best_idx), ptr(poses_out), ptr(step_fam))
        keep = (poses_in, best_idx, poses_out, step_fam)

        def call(_keep=keep):
            rc = fn(h, *args)
            if rc:
                check(rc)
        return call

    def set_options(self, use_graph=True, kernel_timing=False):
        check(self._lib.nvb_set_options(self._h, int(bool(use_graph)), int(bool(kernel_timing))))

    def set_distance_kernel(self, simd_only=False):
        """False (default): tensor-core distance kernel where it applies; True: byte SIMD everywhere."""
        check(self._lib.nvb_set_distance_kernel(self._h, int(bool(simd_only))))

    @property
    def distance_kernel(self):
        """Name of the kernel that scores the current agent batch."""
        k = self._lib.nvb_distance_kernel(self._h)
        return "k2_tc" if k == 1 else "k2_stream" if k == 2 else ("k2_sad_hsv" if self.chem_weight else "k2_sad_v")

    def kernel_time_ms(self):
        """(summed ms, launches) of the distance kernel since set_options(kernel_timing=True)."""
        n = C.c_int64(0)
        ms = float(self._lib.nvb_kernel_time_ms(self._h, C.byref(n)))
        return ms, int(n.value)

    def phase(self, which, fake=False, log_afam=False):
        check(self._lib.nvb_agents_phase(self._h, int(which), int(bool(fake)), int(bool(log_afam))))

    def sync(self):
        check(self._lib.nvb_sync(self._h))

    @property
    def stream_handle(self):
        """cudaStream_t of the engine's stream as an integer."""
        return int(self._lib.nvb_stream_handle(self._h) or 0)

    @property
    def steps_done(self):
        return self._lib.nvb_agents_steps_done(self._h)

    def state(self, coverage=True):
        B = self.n_agents
        npath = 0 if self.training_path is None else len(self.training_path)
        out = dict(poses=np.empty((B, 3)), status=np.empty(B, np.int32),
                   completed=np.empty(B, np.int32), nav_frames=np.empty(B, np.int32),
                   err_sum=np.empty(B), err_n=np.empty(B, np.int32))
        cov = np.zeros((B, npath), np.uint8) if (coverage and npath) else None
        check(self._lib.nvb_agents_get(self._h, ptr(out["poses"]), ptr(out["status"]),
                                       ptr(out["completed"]), ptr(out["nav_frames"]),
                                       ptr(out["err_sum"]), ptr(out["err_n"]), ptr(cov)))
        out["coverage"] = cov
        return out

    def log(self, step0=0, nsteps=None, afam=False):
        if nsteps is None:
            nsteps = self.steps_done - step0
        B, A = self.n_agents, self.n_test_angles
        out = dict(best_idx=np.empty((nsteps, B), np.int16), poses=np.empty((nsteps, B, 3)),
                   step_fam=np.empty((nsteps, B)))
        af = np.empty((nsteps, B, A)) if afam else None
        check(self._lib.nvb_agents_log(self._h, int(step0), int(nsteps), ptr(out["best_idx"]),
                                       ptr(out["poses"]), ptr(out["step_fam"]), ptr(af)))
        out["afam"] = af
        return out

    def run(self, poses, frames, chunk=64):
        """run_experiment's loop (scripts/run_experiment.py:235-258) for a batch
        of start poses: steps every agent until it stops or `frames` (scalar or
        per agent) are done.  Returns per-agent result arrays."""
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        budget = np.broadcast_to(np.asarray(frames, dtype=np.int32), (len(poses),))
        self.set_agents(poses, budget)
        total = int(budget.max()) if len(budget) else 0
        done = 0
        while done < total:
            n = min(chunk, total - done)
            self.step(n)
            done += n
            st = self.state(coverage=False)
            if not np.any((st["status"] == 0) & (st["completed"] < budget)):
                break
        return self.results()

    def results(self, n_consecutive_scenes=0.05):
        """The result record of scripts/run_experiment.py:251-258 per agent."""
        st = self.state()
        cov = st["coverage"].astype(bool)
        with np.errstate(invalid="ignore", divide="ignore"):
            rmsd = np.sqrt(st["err_sum"] / st["err_n"])          # NavBySceneFamiliarity.py:211
        out = dict(st)
        out["stop_status"] = st["status"]
        out["completed_frames"] = st["completed"]
        out["rmsd_error"] = rmsd
        out["path_coverage"] = cov.sum(axis=1) / cov.shape[1]     # :215
        out["percent_forgiving"] = np.array([percent_recapitulated_forgiving(c, n_consecutive_scenes)
                                             for c in cov])
        out["n_captures"] = np.array([n_captures(c, n_consecutive_scenes) for c in cov])
        return out

    # ---- instrumentation ---------------------------------------------------
    @property
    def launch_count(self):
        return int(self._lib.nvb_launch_count(self._h))

    def probe_sad_peak(self, iters=4096):
        return float(self._lib.nvb_probe_sad_peak(self._h, int(iters)))

    def device_sincos(self, x):
        """(sin, cos) of a float64 array as the stepping loop computes them on the device."""
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        s, c = np.empty_like(x), np.empty_like(x)
        check(self._lib.nvb_debug_sincos(self._h, ptr(x), len(x), ptr(s), ptr(c)))
        return s, c

    def probe_mma_peak(self, iters=2048):
        return float(self._lib.nvb_probe_mma_peak(self._h, int(iters)))

    @property
    def tc_planes(self):
        return int(self._lib.nvb_tc_planes(self._h))

    def time_distance_kernel(self, reps=20):
        return float(self._lib.nvb_time_distance_kernel(self._h, int(reps)))

    def timeline(self, nsteps=8, raw=False):
        """Tuning aid: runs `nsteps` step-batches as step() does and returns, for the last one,
        {kernel: (first CTA resident, first dependency met, last CTA done)} in microseconds
        since the step-batch's first stamp, plus 'span' = the whole step-batch.
        raw=True: the per-CTA stamps [kernel][cta][3] in ns (0 = not stamped)."""
        out = np.zeros((6, 2048, 3), np.int64)
        check(self._lib.nvb_debug_timeline(self._h, int(nsteps), ptr(out)))
        if raw:
            return out
        # (slots 4 and 5 hold per-CTA checkpoints of the single-launch step kernel, k3_step_tm)
        names = ["k2", "k3_decide", "k3_ties", "k3_step_tm" if out[4].any() else "k3_move_sample"]
        res = {}
        t0 = None
        for k, name in enumerate(names):
            d = out[k]
            ok = d[:, 0] > 0
            if not ok.any():
                continue
            done = d[ok][:, 2]
            done = done[done > 0]
            res[name] = (int(d[ok][:, 0].min()), int(d[ok][:, 1].min()), int(done.max()) if len(done) else 0)
            t0 = res[name][0] if t0 is None else min(t0, res[name][0])
        out_us = {k: tuple((x - t0) / 1e3 for x in v) for k, v in res.items()}
        if out_us:
            out_us["span"] = max(v[2] for v in out_us.values())
        return out_us

    def device_ptr(self, which):
        return self._lib.nvb_device_ptr(self._h, int(which))


def _window_all(coverage, k):
    """w[i] = all(coverage[i:i + k]) for i in 0 .. n - k (an empty window is all-true)."""
    c = np.asarray(coverage).astype(bool)
    n = len(c)
    if k <= 0:
        return np.ones(n + 1, bool)
    cs = np.concatenate([[0], np.cumsum(c, dtype=np.int64)])
    return (cs[k:] - cs[:n - k + 1]) == k


def percent_recapitulated_forgiving(coverage, n_consecutive_scenes=0.05):
    """NavBySceneFamiliarity.py:218-232 on a coverage bitmap: the end i of the LAST window of
    k consecutive covered scenes, as a fraction of the path (0 if there is none); the
    reference's backwards loop with one cumulative sum."""
    n = len(coverage)
    k = int(n_consecutive_scenes * n)
    if n == 0:
        return 0.
    w = _window_all(coverage, k)          # window coverage[i - k:i] for i = k .. n  <->  w[i - k]
    hits = np.nonzero(w)[0]
    return (int(hits[-1]) + k) / n if len(hits) else 0.


def n_captures(coverage, n_consecutive_scenes=0.05):
    """NavBySceneFamiliarity.py:235-249 on a coverage bitmap: positions i < n - k that are not
    covered while the k scenes after them all are."""
    c = np.asarray(coverage).astype(bool)
    n = len(c)
    k = int(n_consecutive_scenes * n)
    if n - k <= 0:
        return 0
    w = _window_all(c, k)                 # w[j] = all(c[j:j + k])
    return int(np.sum(~c[:n - k] & w[1:n - k + 1]))
