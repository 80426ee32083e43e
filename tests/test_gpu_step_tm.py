"""The single-launch step (csrc/step_tm.cuh, the default after the tensor-core distance kernel)
on worlds built to stress its tie resolution: every view of the library attains
every heading's minimum, so the kernel walks each of its three tie paths --
  * two candidates per view tile evaluated in FP64,
  * tiles whose runner-up ties as well rescanned behind it,
  * more (heading, tile) pairs than the job list holds (the plain loop over all of them) --
and must still take the reference's first maximum (NavBySceneFamiliarity.py:313-315).
Oracle comparison: heading sequence, stop status, frame counts and positions exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(gpu):
    import navsim
    from navsim import synthetic
    from oracle import oracle as O
    return navsim, synthetic, O


def _run(navsim, O, L, w, tpath, poses, frames):
    eng = navsim.NavEngine(L, **w)
    ow = O.World(L, **w)
    assert eng.train_from_path(tpath) == (0, -1)
    assert ow.train_from_path(tpath) == (0, -1)
    eng.set_agents(poses, frames)
    eng.step(frames, log_afam=True)
    assert eng.distance_kernel == "k2_tc"   # (the single-launch step follows the tensor-core kernel only)
    log = eng.log(0, frames, afam=True)
    st = eng.state()
    ref = ow.run_batch(poses, frames, log_best=True)
    assert np.array_equal(st["status"], ref["status"])
    assert np.array_equal(log["best_idx"].T.astype(np.int32), ref["best_idx"])
    assert np.array_equal(st["poses"], ref["poses"])
    eng.close()
    return log, ref


@pytest.mark.parametrize("n_angles,side,step", [(10, 600, 5.0), (20, 2000, 2.0)])
def test_every_view_at_the_minimum(mods, n_angles, side, step):
    """A landscape of one grey level: all glimpses and all views are identical, every heading ties
    with every view (10 headings x 1 tile: runner-up path + rescan; 20 headings x 6 tiles = 120
    pairs: more than the job list holds)."""
    navsim, synthetic, O = mods
    L = np.zeros((side, side, 3), np.uint8)
    L[..., 2] = 128
    w = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=step, n_test_angles=n_angles,
             n_sensor_levels=5, max_distance_to_training_path=450)
    tpath = synthetic.training_path_for(L.shape, step, n_angles, 0.0)
    if n_angles == 20:
        assert n_angles * ((len(tpath) + 255) // 256) > 64   # the overflow path really is taken
    poses = synthetic.start_pose_grid(tpath, 80, n_lat=4, n_deg=4)
    log, ref = _run(navsim, O, L, w, tpath, poses, 12)
    assert np.all(ref["best_idx"][ref["best_idx"] >= 0] == 0)   # all equal: the first maximum is heading 0


def test_two_level_landscape_many_duplicate_views(mods):
    """Broad stripes of two grey levels: long runs of identical views (duplicates inside a tile and
    across tiles) next to genuine differences, ties on most steps."""
    navsim, synthetic, O = mods
    side = 1200
    L = np.zeros((side, side, 3), np.uint8)
    L[..., 2] = np.where((np.arange(side)[None, :] // 150) % 2 == 0, 64, 192)
    w = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=3.0, n_test_angles=10,
             n_sensor_levels=5, max_distance_to_training_path=450)
    tpath = synthetic.training_path_for(L.shape, 3.0, 10, 0.3)
    poses = synthetic.start_pose_grid(tpath, 80, n_lat=4, n_deg=4)
    _run(navsim, O, L, w, tpath, poses, 40)
