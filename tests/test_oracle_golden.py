"""The oracle (oracle/navsim_oracle.c) against golden vectors produced by the
reference's own code (tests/golden/make_golden.py).  CPU only.  Everything is
compared bit for bit: the oracle restates the reference's arithmetic exactly."""
import numpy as np
import pytest

from golden_util import TRAJ, primitives, trajectory
from oracle import oracle as O


def test_fill_sensor_golden():
    g = primitives()
    L = g["landscape"]
    for pose, want in zip(g["fill_poses"], g["fill_out"]):
        buf = np.zeros_like(want)
        assert O.fill_sensor(buf, pose[0], pose[1], pose[2], L) == 0
        assert np.array_equal(buf, want)
    buf = np.zeros_like(g["fill_wrap_out"])
    assert O.fill_sensor(buf, *g["fill_wrap_pose"], L) == 0      # negative indices wrap once
    assert np.array_equal(buf, g["fill_wrap_out"])
    buf = np.zeros_like(g["fill_flip_out"])
    assert O.fill_sensor(buf, *g["fill_flip_pose"], L[::-1, ::-1]) == 0   # negative strides
    assert np.array_equal(buf, g["fill_flip_out"])


def test_downscale_chem_golden():
    g = primitives()
    for key, (fr, fc) in (("down_out_2x4", (2, 4)), ("down_out_4x2", (4, 2)), ("down_out_3x5", (3, 5))):
        for im, want in zip(g["down_in"], g[key]):
            assert np.array_equal(O.downscale_chem(im, fr, fc), want)


def test_quantisation_tables_golden():
    g = primitives()
    for n, want in zip(g["lut_levels"], g["lut_tables"]):
        assert np.array_equal(O.quant_lut(int(n)), want)


@pytest.mark.parametrize("cw", [0.0, 0.3, 1.0])
def test_sads_golden(cw):
    g = primitives()
    want = g["sads_fam_cw%02d" % int(cw * 10)]
    for q, w in zip(g["sads_queries"], want):
        assert np.array_equal(O.sads_hsv(g["sads_scenes"], q, cw), w)     # every double identical


@pytest.mark.parametrize("name", TRAJ)
def test_trajectory_golden(name):
    g, world = trajectory(name)
    w = O.World(g["landscape"], chem_weight=float(g["chem_weight"]), **world)
    assert w.train_from_path(g["tpath"]) == (0, -1)
    assert np.array_equal(w.scenes, g["familiar_scenes"])
    ag = w.new_agent(*g["pose"])
    r = w.run(ag, int(g["frames"]), log_afam=True)
    n = len(g["best_idx"])
    assert r["status"] == int(g["status"]) and r["completed"] == int(g["completed"])
    assert np.array_equal(r["best_idx"][:n], g["best_idx"])
    assert np.array_equal(r["afam"][:n], g["afam"])
    assert np.array_equal(r["pos"][:n], g["pos"])
    assert ag.navigated_for_frames == int(g["navigated_for_frames"])
    assert ag.nav_err == float(g["nav_err"]) and ag.n_nav_err == int(g["n_nav_err"])
    assert np.array_equal(ag._cov, g["coverage"])
