"""The drop-in seam exercised the way a reference user would hit it:

* INTEGRATION.md option B: the REFERENCE's own NavBySceneFamiliarity class (compiled,
  unmodified, from oracle/_ref) running on top of this repo's navsim.util functions;
* two landscapes of the same shape one after the other (the reference driver makes a fresh
  landscape.copy() per trial, scripts/run_experiment.py:186-193);
* single-call APIs interleaved with the resident stepping loop;
* scene_familiarity compared value for value with the oracle.
"""
import numpy as np
import pytest

from cases import agent_grid, build_case
from golden_util import trajectory

pytestmark = pytest.mark.gpu

from test_gpu_parity import FAM_RTOL, POS_TOL  # noqa: E402


@pytest.fixture()
def ref_on_product_util(gpu):
    """The compiled reference class with its three hot-path imports
    (NavBySceneFamiliarity.py:20) re-bound to the product's navsim.util."""
    from oracle import ref_loader
    from navsim import util
    ref = ref_loader.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent when build() ran)")
    saved = (ref.module.fill_sensor_from, ref.module.downscale_chem, ref.module.sads_familiarity)
    ref.module.fill_sensor_from = util.fill_sensor_from
    ref.module.downscale_chem = util.downscale_chem
    ref.module.sads_familiarity = util.sads_familiarity
    util.invalidate_landscape_cache()
    try:
        yield ref, util
    finally:
        (ref.module.fill_sensor_from, ref.module.downscale_chem, ref.module.sads_familiarity) = saved
        util.invalidate_landscape_cache()


@pytest.mark.parametrize("name", ["c1", "chem"])
def test_option_b_reference_class_on_product_util(ref_on_product_util, name):
    """Golden trajectory replayed by the reference's class calling the CUDA path per call:
    everything the reference computed, bit for bit (trig and bookkeeping are the reference's)."""
    ref, util = ref_on_product_util
    g, world = trajectory(name)
    frames = min(int(g["frames"]), 60)
    nsf = ref.NavBySceneFamiliarity(g["landscape"], familiarity_model=util.sads_familiarity(float(g["chem_weight"])),
                                    **world)
    nsf.train_from_path(g["tpath"])
    assert np.array_equal(nsf.familiar_scenes, g["familiar_scenes"])
    nsf.position = (g["pose"][0], g["pose"][1])
    nsf.angle = g["pose"][2]
    done = 0
    try:
        for f in range(frames):
            nsf.step_forward()
            done += 1
            assert np.array_equal(nsf.angle_familiarity, g["afam"][f])
            assert int(np.argmax(nsf.angle_familiarity)) == int(g["best_idx"][f])
            assert (nsf.position[0], nsf.position[1], nsf.angle) == tuple(g["pos"][f])
    except ref.StopNavigationException:
        pass
    assert done == min(frames, int(g["completed"]))


def test_option_b_two_landscapes_same_shape(ref_on_product_util):
    """A second landscape of the same shape (and, as CPython's allocator likes to arrange, the
    same address) must replace the device copy of the first."""
    ref, util = ref_on_product_util
    from navsim import synthetic
    from oracle import oracle as O
    shape_seed = [(7101, 6.0), (7102, 3.0), (7103, 6.0)]
    for seed, sigma in shape_seed:
        L = synthetic.make_landscape(seed, 300, sigma=sigma).copy()     # fresh array each trial
        buf = np.zeros((8, 80, 3), np.uint8)
        want = np.zeros_like(buf)
        util.fill_sensor_from(buf, 150.3, 149.1, 0.4, L)
        assert O.fill_sensor(want, 150.3, 149.1, 0.4, L) == 0
        assert np.array_equal(buf, want), seed
        del L
    # in-place edit + explicit invalidate
    L = synthetic.make_landscape(7104, 300)
    buf = np.zeros((8, 80, 3), np.uint8)
    util.fill_sensor_from(buf, 150.3, 149.1, 0.4, L)
    L[100:200, 100:200, 2] = 255 - L[100:200, 100:200, 2]
    util.invalidate_landscape_cache()
    want = np.zeros_like(buf)
    util.fill_sensor_from(buf, 150.3, 149.1, 0.4, L)
    O.fill_sensor(want, 150.3, 149.1, 0.4, L)
    assert np.array_equal(buf, want)


@pytest.mark.parametrize("name", ["c1_small", "ties"])
def test_single_calls_between_steps(gpu, name):
    """get_sensor_mats / familiarity / familiarity_min reuse the engine's glimpse buffers: a
    call between two step() calls must not leak into the next step (the glimpses sampled
    ahead are dropped and sampled again)."""
    import navsim
    from oracle import oracle as O
    L, w, tpath, pose, frames = build_case(name)
    eng = navsim.NavEngine(L, **w)
    ow = O.World(L, **w)
    assert eng.train_from_path(tpath) == (0, -1) and ow.train_from_path(tpath) == (0, -1)
    poses = np.vstack([np.asarray(pose)[None], agent_grid(tpath, w)])
    K = 24
    eng.set_agents(poses, K)
    rng = np.random.default_rng(5)
    done = 0
    for chunk, what in ((3, "mats"), (5, "fam"), (1, "min"), (7, "mats"), (8, None)):
        eng.step(chunk)
        done += chunk
        q = np.stack([rng.uniform(100, L.shape[1] - 100, 40), rng.uniform(100, L.shape[0] - 100, 40),
                      rng.uniform(-3, 3, 40)], axis=1)
        if what == "mats":
            mats, status = eng.get_sensor_mats(q)
            assert np.all(status == 0)
        elif what == "fam":
            eng.familiarity(ow.scenes[:7])
        elif what == "min":
            eng.familiarity_min(ow.scenes[:33])
    assert done == K
    log = eng.log(0, K)
    st = eng.state()
    ref = ow.run_batch(poses, K, log_best=True)
    assert np.array_equal(st["status"], ref["status"]) and np.array_equal(st["completed"], ref["completed"])
    for b in range(len(poses)):
        n = int(ref["completed"][b]) + (1 if ref["status"][b] in (1, -1) else 0)
        assert np.array_equal(log["best_idx"][:n, b].astype(np.int32), ref["best_idx"][b, :n]), (name, b)
    assert np.allclose(st["poses"], ref["poses"], rtol=0, atol=POS_TOL)


@pytest.mark.parametrize("name", ["c1_small", "chem", "ties"])
def test_scene_familiarity_values(gpu, name):
    """scene_familiarity[n] = min over headings of fam[heading][n] (NavBySceneFamiliarity.py:287,
    301-303) of the step just taken: evaluated lazily by the drop-in class, every double equal to
    the oracle's."""
    import navsim
    from oracle import oracle as O
    L, w, tpath, pose, frames = build_case(name)
    kw = dict(w)
    cw = kw.pop("chem_weight", 0.0)
    nsf = navsim.NavBySceneFamiliarity(L, familiarity_model=navsim.sads_familiarity(cw), **kw)
    nsf.train_from_path(tpath)
    nsf.position = (pose[0], pose[1])
    nsf.angle = pose[2]
    ow = O.World(L, **w)
    ow.train_from_path(tpath)
    ag = ow.new_agent(*pose)
    for f in range(6):
        rc, best, af, sf = ow.step_forward(ag, want_scene_fam=True)
        assert rc == 0
        nsf.step_forward()
        got = nsf.scene_familiarity
        assert got.shape == sf.shape
        assert np.array_equal(got, sf), (name, f)
        assert np.allclose(nsf.angle_familiarity, af, rtol=FAM_RTOL, atol=0)


def test_dropin_parameters_changed_mid_run(gpu):
    """max_distance_to_training_path is a plain attribute the reference reads at every step
    (NavBySceneFamiliarity.py:263): changing it between steps must take effect at the next step
    even though the device has already run ahead with the old value."""
    import navsim
    from oracle import oracle as O
    L, w, tpath, pose, frames = build_case("c1_small")
    kw = dict(w)
    kw["max_distance_to_training_path"] = 450
    nsf = navsim.NavBySceneFamiliarity(L, familiarity_model=navsim.sads_familiarity(0.0), **kw)
    nsf.train_from_path(tpath)
    off = (pose[0] + 25.0, pose[1] - 25.0, pose[2] + 0.5)      # starts ~35 px off the path
    nsf.position = off[:2]
    nsf.angle = off[2]
    ow = O.World(L, **kw)
    ow.train_from_path(tpath)
    ag = ow.new_agent(*off)
    for _ in range(5):
        nsf.step_forward()
        rc, best, af, _sf = ow.step_forward(ag)
        assert rc == 0
    assert nsf.position == (ag.x, ag.y)
    nsf.max_distance_to_training_path = 1.0                    # now every step is "too far"
    ow._w.max_dist = 1.0
    rc, best, af, _sf = ow.step_forward(ag)
    assert rc == O.TOO_FAR
    with pytest.raises(navsim.TooFarFromTrainingPathException):
        nsf.step_forward()
    assert nsf.position == (ag.x, ag.y) and nsf.navigated_for_frames == ag.navigated_for_frames
    assert nsf.replay_restarts == 0


def test_reference_plotting_methods_run_on_the_product(gpu):
    """the reference's compass_plot / _plot_landscape (NavBySceneFamiliarity.py:333-473,
    matplotlib code that only reads public navigator state) lent to the product's navigator through
    nsf.plotting(reference_class): every frame is stepped by the device-resident loop, the path they
    return equals a plain run, a stop reaches their `except StopNavigationException` as the
    reference's own class.  matplotlib is absent here: the reference module's plotting globals are
    mocks that only record calls."""
    from unittest import mock
    import navsim
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent when build() ran)")
    L, w, tpath, pose, frames = build_case("c1_small")
    kw = dict(w)
    kw.pop("chem_weight", None)

    def make():
        nsf = navsim.NavBySceneFamiliarity(L, sensor_px_per_mm=2.0, **kw)
        nsf.train_from_path(tpath)
        nsf.position = (pose[0], pose[1])
        nsf.angle = pose[2]
        return nsf

    plain = make()
    want = [tuple(plain.position)]
    for _ in range(12):
        plain.step_forward()
        want.append(tuple(plain.position))

    nsf = make()
    viz = nsf.plotting(ref.NavBySceneFamiliarity)
    ax = mock.MagicMock()
    ax.plot.return_value = [mock.MagicMock()]   # `path_ln, = main_ax.plot(...)`, NavBySceneFamiliarity.py:463
    names = ("plt", "matplotlib", "fm", "AnchoredSizeBar")
    with mock.patch.multiple(ref.module, **{n: mock.MagicMock() for n in names if hasattr(ref.module, n)}):
        (fig, main_ax), stopped_for, path = viz.compass_plot(ax=ax, frames=12, show_every=3, show_navpath=True)
    assert fig is None and main_ax is ax and stopped_for is None
    assert np.array_equal(path, np.array(want))
    assert ax.imshow.call_count == 1 and ax.imshow.call_args[0][0].shape == L.shape   # the RGB landscape
    assert ax.add_patch.call_count == 8 and ax.arrow.call_count == 4                  # frames 0, 3, 6, 9
    assert np.array_equal(nsf.position, plain.position) and nsf.angle == plain.angle

    # a navigator that stops: too far from the training path after a few frames
    far = navsim.NavBySceneFamiliarity(L, max_distance_to_training_path=1.0, **{k: v for k, v in kw.items()
                                                                              if k != "max_distance_to_training_path"})
    far.train_from_path(tpath)
    far.position = (pose[0] + 30.0, pose[1] + 30.0)
    far.angle = pose[2]
    with mock.patch.multiple(ref.module, **{n: mock.MagicMock() for n in names if hasattr(ref.module, n)}):
        _, stopped_for, path = far.plotting(ref.NavBySceneFamiliarity).compass_plot(ax=mock.MagicMock(), frames=5,
                                                                                   show_scalebar=False)
    assert isinstance(stopped_for, ref.StopNavigationException) and stopped_for.get_code() == -1
    assert len(path) == 2
