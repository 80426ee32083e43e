"""Parity at the sizes bench.py and BASELINE.json quote, not at shrunken ones:

C2  bench.py's own world (WORKLOAD / build_world_inputs): all 1024 agents for the whole frame
    budget on the GPU, the oracle on a fixed 32-agent subsample.
C3  64x64 sensor, 360 headings, 8192 views.
C4  10^6-view library, 10 headings.

Heading sequences, stop status, frame counts and coverage are compared exactly; positions as
in test_gpu_parity.py (POS_TOL).  The oracle runs one agent on C3 / C4 (seconds).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_gpu_parity import FAM_RTOL, POS_TOL  # noqa: E402


def test_c2_bench_world_all_agents_full_budget(gpu):
    import bench
    import navsim
    from navsim import synthetic
    from oracle import oracle as O
    wl = bench.WORKLOAD
    L, tpath, poses, kw = bench.build_world_inputs(wl)
    assert len(poses) == 1024 and L.shape[:2] == (2000, 2000)
    frames = synthetic.default_frames(tpath, wl["step_size"])       # run_experiment.py:238
    eng = navsim.NavEngine(L, **kw)
    ow = O.World(L, **kw)
    assert eng.train_from_path(tpath) == (0, -1) and ow.train_from_path(tpath) == (0, -1)
    assert np.array_equal(eng.familiar_scenes, ow.scenes)
    eng.set_agents(poses, frames)
    eng.step(frames)
    log = eng.log(0, frames)
    st = eng.state()
    sub = np.linspace(0, len(poses) - 1, 32).astype(int)
    ref = ow.run_batch(poses[sub], frames, log_best=True)
    assert np.array_equal(st["status"][sub], ref["status"])
    assert np.array_equal(st["completed"][sub], ref["completed"])
    for i, b in enumerate(sub):
        n = int(ref["completed"][i]) + (1 if ref["status"][i] in (1, -1) else 0)
        assert np.array_equal(log["best_idx"][:n, b].astype(np.int32), ref["best_idx"][i, :n]), b
        assert np.all(log["best_idx"][n:, b] == -1)
    assert np.array_equal(st["coverage"][sub], ref["coverage"])
    assert np.array_equal(st["err_n"][sub], ref["n_nav_err"])
    assert np.allclose(st["err_sum"][sub], ref["nav_err"], rtol=1e-12, atol=0)
    assert np.allclose(st["poses"][sub], ref["poses"], rtol=0, atol=POS_TOL)
    # every agent either used its whole budget or stopped for a reason the reference knows
    assert np.all(np.isin(st["status"], (0, 1, -1, -2)))
    assert np.all((st["status"] != 0) | (st["completed"] == frames))


def _one_agent_steps(eng, ow, pose, frames):
    eng.set_agents([pose], frames)
    eng.step(frames, log_afam=True)
    log = eng.log(0, frames, afam=True)
    st = eng.state()
    ag = ow.new_agent(*pose)
    r = ow.run(ag, frames, log_afam=True)
    n = r["completed"] + (1 if r["status"] in (1, -1) else 0)
    assert n > 0
    assert st["status"][0] == r["status"] and st["completed"][0] == r["completed"]
    assert np.array_equal(log["best_idx"][:n, 0], r["best_idx"][:n])
    assert np.allclose(log["afam"][:n, 0], r["afam"][:n], rtol=FAM_RTOL, atol=0)
    assert np.allclose(log["poses"][:n, 0], r["pos"][:n], rtol=0, atol=POS_TOL)
    assert np.array_equal(st["coverage"][0], ag._cov)


def test_c3_8192_views(gpu):
    """BASELINE configs[2] at its stated size: P = 4096, 360 headings over 180 degrees, 8192 views."""
    import navsim
    from navsim import synthetic
    from oracle import oracle as O
    L = synthetic.make_landscape(3003, 2000, sigma=8.0)
    kw = dict(sensor_dimensions=(64, 64), sensor_pixel_dimensions=(1, 1), step_size=10.0, n_test_angles=360,
              n_sensor_levels=5, saccade_degrees=180., max_distance_to_training_path=450)
    # the reference's path generator at arclen 10/360 gives ~51 000 points: 8192 evenly spaced ones
    full = synthetic.training_path_for(L.shape, 10.0, 360, 0.5)
    tpath = np.ascontiguousarray(full[np.linspace(0, len(full) - 1, 8192).astype(int)])
    eng = navsim.NavEngine(L, **kw)
    ow = O.World(L, **kw)
    assert eng.train_from_path(tpath) == (0, -1) and ow.train_from_path(tpath) == (0, -1)
    assert np.array_equal(eng.familiar_scenes, ow.scenes)
    pose = (tpath[40][0] + 3.0, tpath[40][1] - 2.0, float(np.arctan2(*(tpath[41] - tpath[40])[::-1])) + 0.1)
    _one_agent_steps(eng, ow, pose, 2)


def test_c4_million_views(gpu):
    """BASELINE configs[3] on one GPU: 10^6 views (1414 genuine + random filler),
    one agent, two steps (the oracle scans 10 x 10^6 views per step)."""
    import navsim
    from navsim import synthetic
    from oracle import oracle as O
    L = synthetic.make_landscape(4004, 2000, sigma=6.0)
    kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=10.0, n_test_angles=10,
              n_sensor_levels=5, max_distance_to_training_path=450)
    tpath = synthetic.training_path_for(L.shape, 10.0, 10, 0.0)
    eng = navsim.NavEngine(L, **kw)
    ow = O.World(L, **kw)
    assert ow.train_from_path(tpath) == (0, -1)
    N = 1000000
    rng = np.random.default_rng(4704)
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    scenes = np.zeros((N, 2, 40, 3), np.uint8)
    scenes[..., 2] = levels[rng.integers(0, 5, (N, 2, 40), dtype=np.uint8)]
    scenes[:len(tpath)] = ow.scenes
    scenes[N - 3] = ow.scenes[11]          # duplicate of a genuine view at the far end: the lower index wins
    path = np.vstack([tpath, np.repeat(tpath[-1:], N - len(tpath), axis=0)])
    eng.set_library(scenes, path)
    ow.set_library(scenes, path)
    md, vi = eng.familiarity_min(ow.scenes[11][None])
    assert md[0] == 0 and vi[0] <= 11
    pose = synthetic.start_pose(tpath, (0.05, 3.0), 80)
    _one_agent_steps(eng, ow, pose, 2)


@pytest.mark.parametrize("dup_at", ["far", "near"])
def test_long_path_two_level_prefilter(gpu, dup_at):
    """update_error over a 40 000-point path inside the agent's own CTA: two levels of bounding
    circles prune it ("far": the filler points sit at the end of the path, away from the agents);
    thousands of points on one spot next to the agents overflow the candidate lists and fall
    back to the full scan ("near").  Distances, stops and coverage as the oracle's."""
    import navsim
    from navsim import synthetic
    from oracle import oracle as O
    L = synthetic.make_landscape(5003, 500, sigma=6.0)
    kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=5.0, n_test_angles=10,
              n_sensor_levels=5, max_distance_to_training_path=450)
    tpath = synthetic.training_path_for(L.shape, 5.0, 10, 0.0)
    eng = navsim.NavEngine(L, **kw)
    ow = O.World(L, **kw)
    assert ow.train_from_path(tpath) == (0, -1)
    N = 40000
    rng = np.random.default_rng(11)
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    scenes = np.zeros((N, 2, 40, 3), np.uint8)
    scenes[..., 2] = levels[rng.integers(0, 5, (N, 2, 40))]
    scenes[:len(tpath)] = ow.scenes
    spot = tpath[-1] if dup_at == "far" else tpath[6] + np.array([1.5, -0.5])
    path = np.vstack([tpath, np.repeat(spot[None], N - len(tpath), axis=0)])
    eng.set_library(scenes, path)
    ow.set_library(scenes, path)
    poses = synthetic.start_pose_grid(tpath, 80, n_lat=3, n_deg=3, lat=0.2, deg=8.0)
    frames = 6
    eng.set_agents(poses, frames)
    eng.step(frames)
    st = eng.state()
    ref = ow.run_batch(poses, frames, log_best=True)
    log = eng.log(0, frames)
    assert np.array_equal(st["status"], ref["status"]) and np.array_equal(st["completed"], ref["completed"])
    assert np.array_equal(st["coverage"], ref["coverage"])
    assert np.array_equal(st["err_n"], ref["n_nav_err"]) and np.array_equal(st["err_sum"], ref["nav_err"])
    assert np.array_equal(st["poses"], ref["poses"])
    assert st["coverage"].sum() > 0
    for b in range(len(poses)):
        n = int(ref["completed"][b]) + (1 if ref["status"][b] in (1, -1) else 0)
        assert np.array_equal(log["best_idx"][:n, b].astype(np.int32), ref["best_idx"][b, :n])
