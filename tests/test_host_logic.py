"""Host-side logic of the drop-in (CPU only): quantisation tables, synthetic path
construction, coverage statistics, off-hot-path util functions -- against the
oracle and, when oracle/_ref has been built, the reference itself."""
import numpy as np
import pytest

from oracle import oracle as O, ref_loader


def test_quant_lut_matches_oracle_for_every_level_count():
    from navsim import _cabi
    for n in range(2, 257):
        assert np.array_equal(_cabi.quant_lut(n), O.quant_lut(n)), n
    assert list(_cabi.quant_lut(5)[[0, 31, 32, 95, 96, 159, 160, 223, 224, 255]]) == \
        [0, 0, 63, 63, 127, 127, 191, 191, 255, 255]          # SURVEY.md A3 bin edges
    assert np.array_equal(_cabi.quant_lut(256), np.arange(256))


def test_training_path_shapes():
    from navsim import synthetic
    p = synthetic.training_path_for((2000, 2000), 10.0, 10, 0.0)
    assert p.shape == (1414, 2)                                # N = 1414
    seg = np.linalg.norm(p[1:] - p[:-1], axis=1)
    assert np.sum(seg) <= np.sqrt(2 * 1000.0 ** 2)
    assert synthetic.default_frames(p, 10.0) == int(3.0 * np.sum(seg) / 10.0)
    x, y, a = synthetic.start_pose(p, (0.0, 0.0), 80)
    assert np.allclose([x, y], p[1]) and np.isclose(a, np.pi / 4)


def test_coverage_statistics_match_reference_methods():
    from navsim.engine import n_captures, percent_recapitulated_forgiving
    ref = ref_loader.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(0)
    nsf = ref.NavBySceneFamiliarity(np.zeros((64, 64, 3), np.uint8), (8, 2), 1.0,
                                    sensor_pixel_dimensions=[2, 2])
    for n in (40, 200, 707):
        nsf.training_path = np.zeros((n, 2))
        for _ in range(20):
            cov = rng.random(n) < rng.choice([0.5, 0.9, 0.99])
            if rng.random() < 0.3:
                cov[n // 3:] = True
            nsf._coverage_array = cov
            assert percent_recapitulated_forgiving(cov, 0.05) == nsf.percent_recapitulated_forgiving(0.05)
            assert n_captures(cov, 0.05) == nsf.n_captures(0.05)


def test_off_hot_path_util_matches_reference():
    from navsim import util
    ref = ref_loader.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1)
    labels = rng.integers(0, 5, (30, 40)).astype(np.int64)
    H = rng.integers(0, 256, 4, dtype=np.uint8)
    S = rng.integers(0, 256, 4, dtype=np.uint8)
    a = rng.integers(0, 256, (30, 40, 3), dtype=np.uint8)
    b = a.copy()
    util.set_HS_where_equal(labels, a, H, S)
    ref.util.set_HS_where_equal(labels, b, H, S)
    assert np.array_equal(a, b)
    x, y = rng.random((6, 7)), rng.random((6, 7))
    assert np.isclose(util.ssds(x, y), ref.util.ssds(x, y), rtol=1e-14)
    m = rng.random((16, 16))
    assert np.array_equal(util.diffuse_host(m, 5), ref.util.diffuse(m, 5))   # (the product's diffuse runs on the device: tests/test_gpu_landscape_prep.py)


def test_reference_plotting_proxy_mechanics():
    """nsf.plotting(reference_class) without a device: the proxy hands the
    reference's plotting functions this navigator's state, forwards attribute writes, and turns the
    package's stop exceptions into the classes the reference's `except` clauses look up in their own
    module namespace (the GPU test runs the compiled reference's real compass_plot this way)."""
    from navsim.NavBySceneFamiliarity import (_ReferencePlotting, StopNavigationException,
                                              TooFarFromTrainingPathException)

    class RefStop(Exception):
        pass

    class RefTooFar(RefStop):
        pass

    ns = {"StopNavigationException": RefStop, "TooFarFromTrainingPathException": RefTooFar}
    exec("def _plot_landscape(self, ax, training_path=True):\n"
         "    ax.append(('landscape', self.position, self.training_path is not None and training_path))\n"
         "def compass_plot(self, ax=None, frames=3):\n"
         "    self._plot_landscape(ax)\n"
         "    self._anim_stop_cond = False\n"
         "    stopped = None\n"
         "    for i in range(frames):\n"
         "        try:\n"
         "            self.step_forward()\n"
         "        except StopNavigationException as e:\n"
         "            stopped = e\n"
         "            break\n"
         "    return ax, stopped, self.position\n"
         "def animate(self, frames):\n"
         "    return frames\n", ns)
    Ref = type("NavBySceneFamiliarity", (), {k: ns[k] for k in ("_plot_landscape", "compass_plot", "animate")})

    class Nsf(object):
        def __init__(self):
            self.position, self.training_path, self.n = (0.0, 0.0), [1], 0

        def step_forward(self, fake=False):
            self.n += 1
            self.position = (float(self.n), 0.0)
            if self.n == 3:
                raise TooFarFromTrainingPathException()

    nsf = Nsf()
    viz = _ReferencePlotting(nsf, Ref)
    ax, stopped, pos = viz.compass_plot(ax=[], frames=5)
    assert ax == [("landscape", (0.0, 0.0), True)]
    assert type(stopped) is RefTooFar and isinstance(stopped.__cause__, StopNavigationException)
    assert nsf.n == 3 and pos == (3.0, 0.0) and nsf._anim_stop_cond is False
    assert viz.animate(7) == 7 and viz.n == 3
