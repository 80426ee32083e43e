"""Single-agent front end with the reference's API, served by the batched engine.

Mirrors class NavBySceneFamiliarity and its exceptions
(navsim/NavBySceneFamiliarity.py:22-49, 57-329) so that callers such as
scripts/run_experiment.py's make_nsf / run_experiment (:146-258) work
unchanged: same constructor keywords, same public attributes (position, angle,
angle_familiarity, step_familiarity, scene_familiarity, familiar_scenes,
training_path, navigated_for_frames, ...), same exceptions raised at the same
step.

How it is served: a trajectory is a deterministic function of the state, so
step_forward() does not launch per call.  The engine runs a chunk of steps
ahead on the device (one agent, all headings, resident loop) and step_forward()
replays the log entry by entry, doing the cheap host bookkeeping
(update_error, :252-276) in NumPy exactly like the reference.  Assigning
`position` or `angle` (make_nsf does, run_experiment.py:223-229) discards the
run-ahead.

Plotting (_plot_landscape, compass_plot, animate; :332-661) is visualisation
only and out of scope here.
"""
import warnings

import numpy as np

from .engine import NavEngine
from .util import sads_familiarity, downscale_chem, fill_sensor_from
from . import _cabi


class StopNavigationException(Exception):
    def get_reason(self):
        raise NotImplementedError()

    def get_code(self):
        raise NotImplementedError()

    def __str__(self):
        return self.get_reason()


class ReachedEndOfTrainingPathException(StopNavigationException):
    def get_reason(self):
        return "agent reached end of training path"

    def get_code(self):
        return 1


class NavigatingFailedException(StopNavigationException):
    pass


class TooFarFromTrainingPathException(NavigatingFailedException):
    def get_reason(self):
        return "agent went too far from training path"

    def get_code(self):
        return -1


class OutOfLandscapeBoundsException(NavigatingFailedException):
    def get_reason(self):
        return "agent went too close to boundary of landscape"

    def get_code(self):
        return -2


RUN_AHEAD_STEPS = 64


class NavBySceneFamiliarity(object):

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
                 sensor_px_per_mm=None,
                 familiarity_model=None):
        if familiarity_model is None:
            familiarity_model = sads_familiarity()
        cw = getattr(familiarity_model, "chem_weight", None)
        if cw is None:
            raise TypeError("familiarity_model must come from navsim.util.sads_familiarity(); "
                            "other models have no device implementation")
        self.familiarity_model = familiarity_model
        self._engine = NavEngine(landscape, sensor_dimensions, step_size,
                                 n_test_angles=n_test_angles,
                                 sensor_pixel_dimensions=sensor_pixel_dimensions,
                                 max_distance_to_training_path=max_distance_to_training_path,
                                 n_sensor_levels=n_sensor_levels, mask_middle_n=mask_middle_n,
                                 threshold_factor=threshold_factor,
                                 coverage_threshold_factor=coverage_threshold_factor,
                                 saccade_degrees=saccade_degrees, chem_weight=cw)
        e = self._engine
        self.landscape = e.landscape
        self._position = (0., 0.)
        self._angle = 0.
        self.n_test_angles = n_test_angles
        self.mask_middle_n = mask_middle_n
        self.threshold_factor = threshold_factor
        self.coverage_threshold_factor = coverage_threshold_factor
        self.sensor_px_per_mm = sensor_px_per_mm
        self.saccade_degrees = saccade_degrees
        self.angle_offsets = e.angle_offsets
        self.sensor_dimensions = e.sensor_dimensions
        self.sensor_pixel_dimensions = e.sensor_pixel_dimensions
        self._sensor_r = e._sensor_r
        self.n_sensor_pixels = np.prod(self.sensor_dimensions)
        self.n_sensor_levels = e.n_sensor_levels
        self.step_size = step_size
        self.angle_familiarity = np.empty(shape=n_test_angles)
        self.step_familiarity = np.inf
        self.max_distance_to_training_path = max_distance_to_training_path
        self._ahead = None
        self.replay_restarts = 0      # run-aheads recomputed because the device and the host bookkeeping disagreed
        self.clear_training()
        self.reset_error()

    # ---- pose: plain attributes in the reference; assigning drops the run-ahead
    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, value):
        self._position = value
        self._ahead = None

    @property
    def angle(self):
        return self._angle

    @angle.setter
    def angle(self, value):
        self._angle = value
        self._ahead = None

    # ---- training ----------------------------------------------------------
    def train_from_path(self, points):
        if self.training_path is not None:
            raise ValueError("Tried to train NavBySceneFamiliarity more than once.")
        points = np.asarray(points, dtype=np.float64)
        rc, bad = self._engine.train_from_path(points)
        if rc == _cabi.OUT_OF_BOUNDS:
            raise OutOfLandscapeBoundsException()
        if rc == _cabi.INDEX_ERROR:
            raise IndexError("Index out of bounds (axis 0)")
        self.training_path = points
        self.training_path_length = self._engine.training_path_length
        self._scene_familiarity = np.zeros(shape=len(points), dtype=float)
        self._scene_fam_pose = None
        self.reset_error()
        self._familiarity_func = _FamiliarityFunc(self)
        self._ahead = None

    def clear_training(self):
        self.training_path = None
        self._familiarity_func = None
        self._scene_familiarity = None
        self._scene_fam_pose = None
        self.training_path_length = None

    @property
    def familiar_scenes(self):
        return None if self.training_path is None else self._engine.familiar_scenes

    def get_sensor_mat(self, position, angle):
        out, status = self._engine.get_sensor_mats([(position[0], position[1], angle)])
        if status[0] == _cabi.OUT_OF_BOUNDS:
            raise OutOfLandscapeBoundsException()
        if status[0] == _cabi.INDEX_ERROR:
            raise IndexError("Index out of bounds (axis 0)")
        return out[0]

    # ---- error bookkeeping (host NumPy, same arithmetic as :195-276) --------
    def reset_error(self):
        self.stopped_with_exception = None
        self.navigated_for_frames = 0
        self._navigation_error = 0.0
        self._n_navigation_error = 0
        if self.training_path is not None:
            self._coverage_array = np.zeros(len(self.training_path), dtype=bool)
        self._ahead = None

    @property
    def navigation_error(self):
        return np.sqrt(self._navigation_error / self._n_navigation_error)

    @property
    def percent_recapitulated(self):
        return np.sum(self._coverage_array) / len(self._coverage_array)

    def percent_recapitulated_forgiving(self, n_consecutive_scenes=0.05):
        from .engine import percent_recapitulated_forgiving
        return percent_recapitulated_forgiving(self._coverage_array, n_consecutive_scenes)

    def n_captures(self, n_consecutive_scenes=0.05):
        from .engine import n_captures
        return n_captures(self._coverage_array, n_consecutive_scenes)

    def update_error(self):
        self.navigated_for_frames += 1
        d = self.training_path - self._position
        d *= d
        dist = np.sqrt(np.sum(d, axis=1))
        diff = np.min(dist)
        if diff > self.max_distance_to_training_path:
            raise TooFarFromTrainingPathException()
        self._navigation_error += diff * diff
        self._n_navigation_error += 1
        cvge_thresh = self.coverage_threshold_factor * self.step_size
        if diff <= cvge_thresh:
            self._coverage_array |= (dist <= cvge_thresh)

    # ---- scene_familiarity: only plotting reads it, so it is evaluated on demand
    @property
    def scene_familiarity(self):
        if self.training_path is None:
            return None
        if self._scene_fam_pose is not None:
            (x, y), ang = self._scene_fam_pose
            angles = (ang + self.angle_offsets) % (2 * np.pi)
            mats, status = self._engine.get_sensor_mats(
                [(x, y, a) for a in angles])
            if np.all(status == 0):
                self._scene_familiarity = self._engine.familiarity(mats).min(axis=0)
            else:
                self._scene_familiarity = np.full(len(self.training_path), np.inf)
            self._scene_fam_pose = None
        return self._scene_familiarity

    # ---- stepping -----------------------------------------------------------
    def _nav_params(self):
        """The plain attributes the reference reads at every step (:263,271,328): the caller may
        change them between steps, which invalidates a run-ahead computed with the old values."""
        return (self.max_distance_to_training_path, self.step_size, self.threshold_factor,
                self.coverage_threshold_factor)

    def _run_ahead(self, fake):
        e = self._engine
        (e.max_distance_to_training_path, e.step_size, e.threshold_factor,
         e.coverage_threshold_factor) = params = self._nav_params()
        e.set_agents([(self._position[0], self._position[1], self._angle)])
        n = 1 if fake else RUN_AHEAD_STEPS
        e.step(n, fake=fake, log_afam=True)
        log = e.log(0, n, afam=True)
        st = e.state(coverage=False)
        self._ahead = dict(i=0, n=n, log=log, status=int(st["status"][0]), fake=fake, params=params)

    def step_forward(self, fake=False):
        if self.training_path is None:
            raise TypeError("'NoneType' object is not callable")   # untrained, as in the reference
        ah = self._ahead
        if ah is None or ah["i"] >= ah["n"] or ah["fake"] != bool(fake) or ah["params"] != self._nav_params():
            self._run_ahead(bool(fake))
            ah = self._ahead
        i = ah["i"]
        log = ah["log"]
        best = int(log["best_idx"][i, 0])
        self._scene_fam_pose = (tuple(self._position), self._angle)
        if best < 0:
            # the device agent stopped before taking this step
            status = ah["status"]
            self._ahead = None
            if status == _cabi.OUT_OF_BOUNDS:
                self.angle_familiarity[:] = np.nan
                self._scene_fam_pose = None
                self._scene_familiarity = np.full(len(self.training_path), np.inf)
                raise OutOfLandscapeBoundsException()
            if status == _cabi.INDEX_ERROR:
                raise IndexError("Index out of bounds (axis 0)")
            # The device stopped this agent (too far / end of path) at a step where the host-side
            # bookkeeping below did not.  Positions are bit-identical, so this needs a threshold
            # compared differently in the last bit (np.linalg.norm's BLAS dot may fuse where the
            # kernel does not, :328): the host state is the reference's, so the run-ahead is
            # recomputed from it -- counted, never silent.
            self.replay_restarts += 1
            if self.replay_restarts > 8 + self.navigated_for_frames:
                raise RuntimeError("device log and host bookkeeping keep disagreeing (status %d)" % status)
            self._run_ahead(bool(fake))
            return self.step_forward(fake)
        self.angle_familiarity[:] = log["afam"][i, 0]
        self.step_familiarity = float(log["step_fam"][i, 0])
        x, y, ang = log["poses"][i, 0]
        self._position = (x, y)
        self._angle = ang
        ah["i"] = i + 1
        if fake:
            self._ahead = None
            return
        self.update_error()
        if np.linalg.norm(self.training_path[-1] - self._position) <= self.threshold_factor * self.step_size:
            raise ReachedEndOfTrainingPathException()

    # ---- visualisation -------------------------------------------------------
    # The reference's plotting methods (NavBySceneFamiliarity.py:333-661: _plot_landscape,
    # compass_plot, animate) are matplotlib code that only reads public state this class has
    # as well (position, angle, angle_familiarity, scene_familiarity, angle_offsets,
    # training_path, sensor geometry, step_forward, the error properties).  They are not
    # rebuilt here; a user who has the reference installed lends them to this navigator:
    #     viz = nsf.plotting(reference_navsim.NavBySceneFamiliarity)
    #     (fig, ax), stopped_for, path = viz.compass_plot(frames=40, show_navpath=True)
    # and every frame they draw is stepped by the device-resident loop.
    @property
    def _landscape_rgb(self):
        """RGB view of the HSV landscape for imshow (NavBySceneFamiliarity.py:74), made on first use."""
        rgb = self.__dict__.get("_landscape_rgb_cache")
        if rgb is None:
            from PIL import Image
            rgb = np.asarray(Image.fromarray(np.ascontiguousarray(self.landscape), mode='HSV').convert('RGB'))
            self.__dict__["_landscape_rgb_cache"] = rgb
        return rgb

    def plotting(self, reference_class):
        """The reference class's visualisation methods bound to this navigator (see above).
        `reference_class` is the reference's NavBySceneFamiliarity; stop exceptions raised by
        step_forward() reach its code as the reference's own exception classes."""
        return _ReferencePlotting(self, reference_class)

    def _plot_landscape(self, *a, **k):
        raise NotImplementedError("plotting lives in the reference: use nsf.plotting(reference_class) "
                                  "(matplotlib is needed; DESIGN.md, SURVEY 8(f) N4)")

    compass_plot = animate = _plot_landscape


class _ReferencePlotting(object):
    """Proxy handed to the reference's plotting functions as `self`: attribute reads and writes
    go to the navigator, the three plotting methods come from the reference class, and
    step_forward() translates this package's stop exceptions into the reference module's classes of
    the same name (its `except StopNavigationException` clauses would not match ours)."""

    _METHODS = ("_plot_landscape", "compass_plot", "animate")

    def __init__(self, nsf, reference_class):
        import sys
        object.__setattr__(self, "_nsf", nsf)
        ref_globals = None
        for name in self._METHODS:
            fn = getattr(reference_class, name)
            fn = getattr(fn, "__func__", fn)
            object.__setattr__(self, "_fn_" + name, fn)
            if ref_globals is None:
                ref_globals = getattr(fn, "__globals__", None)   # the namespace its `except` clauses look names up in
        if ref_globals is None:
            mod = sys.modules.get(getattr(reference_class, "__module__", ""), None)
            ref_globals = vars(mod) if mod is not None else {}
        object.__setattr__(self, "_ref_globals", ref_globals)

    def __getattr__(self, name):
        if name in _ReferencePlotting._METHODS:
            fn = object.__getattribute__(self, "_fn_" + name)
            return lambda *a, **k: fn(self, *a, **k)
        return getattr(object.__getattribute__(self, "_nsf"), name)

    def __setattr__(self, name, value):
        setattr(object.__getattribute__(self, "_nsf"), name, value)

    def step_forward(self, fake=False):
        nsf = object.__getattribute__(self, "_nsf")
        try:
            return nsf.step_forward(fake)
        except StopNavigationException as e:
            ref_exc = object.__getattribute__(self, "_ref_globals").get(type(e).__name__)
            if ref_exc is None or ref_exc is type(e):
                raise
            raise ref_exc() from e


class _FamiliarityFunc(object):
    """What familiarity_model(familiar_scenes) returns in the reference
    (util.pyx:14-24): callable (scene, fambuf) with .max_familiarity."""

    def __init__(self, nsf):
        self._nsf = nsf
        dims = nsf.sensor_dimensions
        self.max_familiarity = int(dims[0]) * int(dims[1])

    def __call__(self, scene, fambuf):
        fambuf[:] = self._nsf._engine.familiarity(np.asarray(scene)[None])[0]
