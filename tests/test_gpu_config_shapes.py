"""The other BASELINE.json configs at sizes the oracle finishes in seconds:
C3 (64x64 sensor, 360-heading sweep) and C4 (very large library, few agents)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POS_TOL = 0.0    # px: exact -- the device computes sin / cos with glibc's own algorithm (csrc/glibc_trig.cuh)
FAM_RTOL = 1e-12


def _compare(eng, ow, poses, frames):
    eng.set_agents(poses, frames)
    eng.step(frames, log_afam=True)
    log = eng.log(0, frames, afam=True)
    st = eng.state()
    for b, p in enumerate(poses):
        ag = ow.new_agent(*p)
        r = ow.run(ag, frames, log_afam=True)
        n = r["completed"] + (1 if r["status"] in (1, -1) else 0)
        assert st["status"][b] == r["status"] and st["completed"][b] == r["completed"]
        assert np.array_equal(log["best_idx"][:n, b], r["best_idx"][:n])
        assert np.allclose(log["afam"][:n, b], r["afam"][:n], rtol=FAM_RTOL, atol=0)
        assert np.allclose(log["poses"][:n, b], r["pos"][:n], rtol=0, atol=POS_TOL)
        assert np.array_equal(st["coverage"][b], ag._cov)


def test_c3_large_sensor_fine_sweep(gpu):
    """64x64 sensor pixels of 1x1 (P = 4096: the K loop of the distance kernel), 360 headings
    over 360 degrees (first and last heading coincide: the first must win)."""
    import navsim
    from navsim import synthetic
    from oracle import oracle as O
    L = synthetic.make_landscape(5001, 400, sigma=8.0)
    kw = dict(sensor_dimensions=(64, 64), sensor_pixel_dimensions=(1, 1), step_size=6.0, n_test_angles=360,
              n_sensor_levels=5, saccade_degrees=360., max_distance_to_training_path=450)
    tpath = synthetic.training_path_for(L.shape, 6.0, 1, 0.5)[::6][:48]      # 48 views
    eng = navsim.NavEngine(L, **kw)
    ow = O.World(L, **kw)
    assert eng.train_from_path(tpath) == (0, -1) and ow.train_from_path(tpath) == (0, -1)
    assert np.array_equal(eng.familiar_scenes, ow.scenes)
    poses = np.array([[tpath[1][0] + 2.0, tpath[1][1] - 1.5, 0.9], [tpath[5][0], tpath[5][1], 0.3]])
    _compare(eng, ow, poses, 3)


def test_c4_large_library_few_agents(gpu):
    """70 000 views (beyond the fused-step limit: K1, K2, decide, grid-wide ties, move),
    10 headings, 2 agents: the library-streaming configuration of the distance kernel."""
    import navsim
    from navsim import synthetic
    from oracle import oracle as O
    L = synthetic.make_landscape(5002, 500, sigma=6.0)
    kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=5.0, n_test_angles=10,
              n_sensor_levels=5, max_distance_to_training_path=450)
    tpath = synthetic.training_path_for(L.shape, 5.0, 10, 0.0)            # ~700 genuine views
    eng = navsim.NavEngine(L, **kw)
    ow = O.World(L, **kw)
    assert ow.train_from_path(tpath) == (0, -1)
    rng = np.random.default_rng(7)
    N = 70000
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    scenes = np.zeros((N, 2, 40, 3), np.uint8)
    scenes[..., 2] = levels[rng.integers(0, 5, (N, 2, 40))]
    scenes[:len(tpath)] = ow.scenes                                         # genuine views first
    scenes[N - 5] = ow.scenes[7]                                            # a duplicate far away: lower index must win
    path = np.vstack([tpath, np.repeat(tpath[-1:], N - len(tpath), axis=0)])
    eng.set_library(scenes, path)
    ow.set_library(scenes, path)
    from navsim.synthetic import start_pose
    poses = np.array([start_pose(tpath, (0.05, 3.0), 80), start_pose(tpath, (-0.1, -4.0), 80)])
    md, vi = eng.familiarity_min(ow.scenes[7][None])
    assert md[0] == 0 and vi[0] <= 7
    _compare(eng, ow, poses, 3)
