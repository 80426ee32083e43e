"""The CUDA path (through the C ABI) against golden vectors produced by the
reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from golden_util import TRAJ, primitives, trajectory

pytestmark = pytest.mark.gpu

POS_TOL = 0.0    # px: exact -- the device computes sin / cos with glibc's own algorithm (csrc/glibc_trig.cuh)
FAM_RTOL = 1e-12  # angle_familiarity where several views tie at a heading's minimum


def test_util_functions_golden(gpu):
    from navsim import util
    g = primitives()
    L = g["landscape"]
    util.invalidate_landscape_cache()
    for pose, want in zip(g["fill_poses"], g["fill_out"]):
        buf = np.zeros_like(want)
        util.fill_sensor_from(buf, pose[0], pose[1], pose[2], L)
        assert np.array_equal(buf, want)
    buf = np.zeros_like(g["fill_wrap_out"])
    util.fill_sensor_from(buf, *g["fill_wrap_pose"], L)
    assert np.array_equal(buf, g["fill_wrap_out"])
    buf = np.zeros_like(g["fill_flip_out"])
    util.fill_sensor_from(buf, *g["fill_flip_pose"], L[::-1, ::-1])
    assert np.array_equal(buf, g["fill_flip_out"])
    for key, (fr, fc) in (("down_out_2x4", (2, 4)), ("down_out_4x2", (4, 2)), ("down_out_3x5", (3, 5))):
        for im, want in zip(g["down_in"], g[key]):
            assert np.array_equal(util.downscale_chem(im, fr, fc), want)
    for cw in (0.0, 0.3, 1.0):
        func = util.sads_familiarity(cw)(g["sads_scenes"])
        for q, want in zip(g["sads_queries"], g["sads_fam_cw%02d" % int(cw * 10)]):
            fam = np.empty(len(g["sads_scenes"]))
            func(q, fam)
            assert np.array_equal(fam, want)


def test_quantisation_tables_golden(gpu):
    from navsim import _cabi
    g = primitives()
    for n, want in zip(g["lut_levels"], g["lut_tables"]):
        assert np.array_equal(_cabi.quant_lut(int(n)), want)


@pytest.mark.parametrize("name", TRAJ)
def test_trajectory_golden(gpu, name):
    import navsim
    g, world = trajectory(name)
    frames = int(g["frames"])
    eng = navsim.NavEngine(g["landscape"], chem_weight=float(g["chem_weight"]), **world)
    assert eng.train_from_path(g["tpath"]) == (0, -1)
    assert np.array_equal(eng.familiar_scenes, g["familiar_scenes"])
    eng.set_agents([g["pose"]], frames)
    eng.step(frames, log_afam=True)
    log = eng.log(0, frames, afam=True)
    st = eng.state()
    n = len(g["best_idx"])
    assert st["status"][0] == int(g["status"]) and st["completed"][0] == int(g["completed"])
    assert np.array_equal(log["best_idx"][:n, 0], g["best_idx"])          # heading sequence: exact
    assert np.all(log["best_idx"][n:, 0] == -1)
    assert np.allclose(log["afam"][:n, 0], g["afam"], rtol=FAM_RTOL, atol=0)
    assert np.allclose(log["poses"][:n, 0], g["pos"], rtol=0, atol=POS_TOL)
    assert st["nav_frames"][0] == int(g["navigated_for_frames"])
    assert st["err_n"][0] == int(g["n_nav_err"])
    assert np.isclose(st["err_sum"][0], float(g["nav_err"]), rtol=1e-12)
    assert np.array_equal(st["coverage"][0], g["coverage"])
    res = eng.results()
    fmt = "{:6f}".format                                                   # run_experiment.py:44
    assert fmt(res["path_coverage"][0]) == fmt(float(g["path_coverage"]))
    assert fmt(res["rmsd_error"][0]) == fmt(float(g["rmsd"]))
    assert fmt(res["percent_forgiving"][0]) == fmt(float(g["percent_forgiving"]))
    assert int(res["n_captures"][0]) == int(g["n_captures"])


@pytest.mark.parametrize("name", ["c1", "chem"])
def test_dropin_class_golden(gpu, name):
    """The reference-API class, driven like scripts/run_experiment.py:235-258."""
    import navsim
    g, world = trajectory(name)
    nsf = navsim.NavBySceneFamiliarity(g["landscape"],
                                       familiarity_model=navsim.sads_familiarity(float(g["chem_weight"])),
                                       **world)
    nsf.train_from_path(g["tpath"])
    nsf.position = (g["pose"][0], g["pose"][1])
    nsf.angle = g["pose"][2]
    status, done = 0, 0
    try:
        for _ in range(int(g["frames"])):
            nsf.step_forward()
            done += 1
    except navsim.StopNavigationException as e:
        status = e.get_code()
    assert (status, done) == (int(g["status"]), int(g["completed"]))
    assert "{:6f}".format(nsf.percent_recapitulated) == "{:6f}".format(float(g["path_coverage"]))
    assert "{:6f}".format(nsf.navigation_error) == "{:6f}".format(float(g["rmsd"]))
    assert nsf.n_captures(0.05) == int(g["n_captures"])
