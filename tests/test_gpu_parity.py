"""Parity of the CUDA path (through the C ABI) against the oracle.

Bit-exact for bytes, indices and integer sums.  FP64 familiarity values: the
single-call closure (sads_familiarity) is bit-exact; in the stepping loop
angle_familiarity is bit-exact unless several views tie at a heading's integer
minimum, where the reference's max over doubles picks one by rounding noise
(bound: FAM_RTOL; > 90 % of the values are checked to be bit-exact).  Headings
that tie for the step's best integer minimum are always resolved in exact FP64
over every tied view, so the chosen heading index is exact.  Agent positions are bit-identical (POS_TOL = 0): the
device's sin / cos restate glibc's algorithm (csrc/glibc_trig.cuh, tests/test_glibc_trig.py).
"""
import numpy as np
import pytest

from cases import CASES, agent_grid, build_case

pytestmark = pytest.mark.gpu

POS_TOL = 0.0    # px: exact -- the device computes sin / cos with glibc's own algorithm (csrc/glibc_trig.cuh)
FAM_RTOL = 1e-12  # relative, angle_familiarity of non-winning headings (north_star asks for FP32-level 1e-6)


@pytest.fixture(scope="module")
def mods(gpu):
    import navsim
    from navsim import util
    from oracle import oracle as O
    return navsim, util, O


def test_no_cpu_fallback(mods):
    navsim, util, O = mods
    from navsim import _cabi
    assert _cabi.lib().nvb_version().decode().startswith("navsim_b200")


def test_fill_sensor_from(mods):
    navsim, util, O = mods
    from navsim import synthetic
    L = synthetic.make_landscape(11, 300, n_chemicals=3)
    rng = np.random.default_rng(0)
    for land in (L, L[::-1], L[:, ::-1], L[::-1, ::-1], np.ascontiguousarray(L[::2, ::2])):
        util.invalidate_landscape_cache()
        for _ in range(25):
            Hpx, Wpx = int(rng.choice([2, 4, 8, 16, 64])), int(rng.choice([2, 8, 16, 40, 80]))
            x, y = rng.uniform(60, land.shape[1] - 60), rng.uniform(60, land.shape[0] - 60)
            ang = rng.uniform(-7, 7)
            a = np.zeros((Hpx, Wpx, 3), np.uint8)
            b = np.zeros_like(a)
            util.fill_sensor_from(a, x, y, ang, land)
            assert O.fill_sensor(b, x, y, ang, land) == 0
            assert np.array_equal(a, b)


def test_fill_sensor_wraparound_and_index_error(mods):
    navsim, util, O = mods
    from navsim import synthetic
    L = synthetic.make_landscape(12, 300)
    util.invalidate_landscape_cache()
    a = np.zeros((64, 64, 3), np.uint8)
    b = np.zeros_like(a)
    # negative indices wrap once (util.pyx:137 has wraparound on)
    util.fill_sensor_from(a, 33., 33., 0.7, L)
    assert O.fill_sensor(b, 33., 33., 0.7, L) == 0
    assert np.array_equal(a, b)
    # overshoot on the high side raises IndexError
    assert O.fill_sensor(b, 267., 267., 0.7, L) == O.INDEX_ERROR
    with pytest.raises(IndexError):
        util.fill_sensor_from(a, 267., 267., 0.7, L)
    # non-contiguous output buffer
    big = np.zeros((8, 80, 6), np.uint8)
    view = big[:, :, ::2]
    ref = np.zeros((8, 80, 3), np.uint8)
    util.fill_sensor_from(view, 150.2, 140.9, 2.0, L)
    O.fill_sensor(ref, 150.2, 140.9, 2.0, L)
    assert np.array_equal(view, ref)
    with pytest.raises(ValueError):
        util.fill_sensor_from(np.zeros((8, 80, 3), np.float32), 1., 1., 0., L)


def test_downscale_chem(mods):
    navsim, util, O = mods
    rng = np.random.default_rng(1)
    for t in range(60):
        fr, fc = int(rng.choice([1, 2, 3, 4])), int(rng.choice([1, 2, 4, 5]))
        img = rng.integers(0, 256, (fr * int(rng.integers(1, 6)), fc * int(rng.integers(1, 9)), 3),
                           dtype=np.uint8)
        if t % 2:
            img[..., 0] = rng.integers(0, 3, img.shape[:2]) * 85   # few hues: real votes and ties
        if t % 3 == 0:
            img[..., 1] = 127                                       # S sums that wrap past 255
        assert np.array_equal(util.downscale_chem(img, fr, fc), O.downscale_chem(img, fr, fc))
    # ragged: rows/cols not a multiple of the factor are dropped (util.pyx:102)
    img = rng.integers(0, 256, (7, 11, 3), dtype=np.uint8)
    assert np.array_equal(util.downscale_chem(img, 2, 4), O.downscale_chem(img, 2, 4))


@pytest.mark.parametrize("name", ["c1_small", "chem", "square", "ties"])
def test_get_sensor_mat_and_library(mods, name):
    navsim, util, O = mods
    L, w, tpath, pose, frames = build_case(name)
    eng = navsim.NavEngine(L, **w)
    ow = O.World(L, **w)
    rng = np.random.default_rng(2)
    poses = np.stack([rng.uniform(0, L.shape[1], 200), rng.uniform(0, L.shape[0], 200),
                      rng.uniform(-7, 7, 200)], axis=1)
    mats, status = eng.get_sensor_mats(poses)
    n_ok = 0
    for i, p in enumerate(poses):
        rc, ref = ow.get_sensor_mat(p[:2], p[2])
        assert status[i] == rc
        if rc == 0:
            n_ok += 1
            assert np.array_equal(mats[i], ref)
    assert n_ok > 50 and n_ok < 200       # both in-bounds and out-of-bounds poses were exercised
    assert eng.train_from_path(tpath) == (0, -1)
    assert ow.train_from_path(tpath) == (0, -1)
    assert np.array_equal(eng.familiar_scenes, ow.scenes)
    # a path that leaves the landscape fails at the same point
    bad = tpath.copy()
    bad[5] = (1.0, 1.0)
    eng2 = navsim.NavEngine(L, **w)
    assert eng2.train_from_path(bad) == O.World(L, **w).train_from_path(bad) == (O.OUT_OF_BOUNDS, 5)


@pytest.mark.parametrize("cw", [0.0, 0.3, 1.0])
def test_familiarity_exact_fp64(mods, cw):
    """sads_familiarity(cw)(scenes)(scene, fambuf): every double bit-identical."""
    navsim, util, O = mods
    rng = np.random.default_rng(3)
    N, H, W = 257, 3, 21
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    scenes = levels[rng.integers(0, 5, (N, H, W, 3))]
    scenes[..., 0] = rng.integers(0, 2, (N, H, W)) * 127
    func = util.sads_familiarity(cw)(scenes)
    assert func.max_familiarity == H * W
    for t in range(4):
        scene = scenes[rng.integers(N)].copy()
        scene[rng.integers(H), rng.integers(W)] = 17
        fam = np.full(N, np.nan)
        func(scene, fam)
        ref = O.sads_hsv(scenes, scene, cw)
        assert np.array_equal(fam, ref)


# P (sensor pixels) values chosen to hit every shared-memory chunking mode of the
# distance kernel: 16-B chunks per row = 1, 2, 3, 4, 5, 6->7, 7, 8, and >8 (K loop).
@pytest.mark.parametrize("W,H", [(8, 2), (16, 2), (20, 2), (16, 4), (40, 2), (24, 4), (28, 4),
                                 (32, 4), (50, 5), (64, 64)])
@pytest.mark.parametrize("G", [1, 10, 40, 300])
def test_distance_kernel_min_argmin(mods, W, H, G):
    """K2 vs the oracle's integer sums: minimum and LOWEST view index, exact."""
    navsim, util, O = mods
    if W * H > 1000 and G > 40:
        pytest.skip("oracle too slow at this size")
    rng = np.random.default_rng(W * 1000 + H * 10 + G)
    N = int(rng.integers(130, 700)) if W * H < 1000 else 150
    L = np.zeros((64, 64, 3), np.uint8)
    eng = navsim.NavEngine(L, (W, H), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2))
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    scenes = levels[rng.integers(0, 5, (N, H, W, 3))]
    scenes[N // 2] = scenes[3]          # duplicate views: the lower index must win
    scenes[N - 1] = scenes[3]
    eng.set_library(scenes)
    q = levels[rng.integers(0, 5, (G, H, W, 3))]
    q[0] = scenes[3]
    md, vi = eng.familiarity_min(q)
    for g in range(G):
        _, vt = O.sad_int(scenes, q[g])
        assert md[g] == vt.min()
        assert vi[g] == int(np.argmin(vt))
    assert md[0] == 0 and vi[0] == 3


@pytest.mark.parametrize("name", list(CASES))
def test_trajectories_match_oracle(mods, name):
    """Resident stepping loop vs the oracle's step_forward loop: heading index
    sequence, stop status and frame counts exact; angle_familiarity bit-exact;
    positions within POS_TOL."""
    navsim, util, O = mods
    L, w, tpath, pose, frames = build_case(name)
    frames = min(frames, 150)
    eng = navsim.NavEngine(L, **w)
    ow = O.World(L, **w)
    assert eng.train_from_path(tpath) == (0, -1)
    assert ow.train_from_path(tpath) == (0, -1)
    poses = np.vstack([np.asarray(pose)[None], agent_grid(tpath, w)])
    eng.set_agents(poses, frames)
    eng.step(frames, log_afam=True)
    log = eng.log(0, frames, afam=True)
    st = eng.state()
    n_tied_steps = n_exact = n_vals = 0
    for b, p in enumerate(poses):
        ag = ow.new_agent(*p)
        r = ow.run(ag, frames, log_afam=True)
        n = r["completed"] + (1 if r["status"] in (1, -1) else 0)   # steps that moved the agent
        assert st["status"][b] == r["status"], (name, b)
        assert st["completed"][b] == r["completed"]
        assert np.array_equal(log["best_idx"][:n, b], r["best_idx"][:n]), (name, b)
        assert np.all(log["best_idx"][n:, b] == -1)
        assert np.allclose(log["afam"][:n, b], r["afam"][:n], rtol=FAM_RTOL, atol=0), (name, b)
        assert np.allclose(log["step_fam"][:n, b], r["afam"][:n].max(axis=1), rtol=FAM_RTOL, atol=0), (name, b)
        n_exact += int(np.sum(log["afam"][:n, b] == r["afam"][:n]))
        n_vals += n * r["afam"].shape[1]
        assert np.allclose(log["poses"][:n, b], r["pos"][:n], rtol=0, atol=POS_TOL)
        assert st["nav_frames"][b] == ag.navigated_for_frames
        assert st["err_n"][b] == ag.n_nav_err
        assert np.isclose(st["err_sum"][b], ag.nav_err, rtol=1e-12, atol=0)
        assert np.array_equal(st["coverage"][b], ag._cov)
        top = r["afam"][:n].max(axis=1, keepdims=True)
        n_tied_steps += int(np.sum(np.sum(r["afam"][:n] >= top - 1e-9, axis=1) > 1))
    if name in ("ties", "chem1"):
        assert n_tied_steps > 0       # the tie resolver really was exercised
    assert n_exact > 0.9 * n_vals      # the overwhelming majority is bit-exact


def test_out_of_bounds_and_budget(mods):
    navsim, util, O = mods
    L, w, tpath, pose, frames = build_case("c1_small")
    eng = navsim.NavEngine(L, **w)
    ow = O.World(L, **w)
    eng.train_from_path(tpath)
    ow.train_from_path(tpath)
    # one agent starts outside the bounds test, one has a short frame budget
    poses = np.array([[10.0, 300.0, 0.3], list(pose), list(pose)])
    eng.set_agents(poses, [50, 7, 50])
    eng.step(20)
    st = eng.state()
    assert st["status"][0] == O.OUT_OF_BOUNDS and st["completed"][0] == 0
    assert st["status"][1] == 0 and st["completed"][1] == 7
    ref = ow.run_batch(poses[1:2], 7)
    assert np.allclose(st["poses"][1], ref["poses"][0], rtol=0, atol=POS_TOL)
    assert st["completed"][2] == 20


@pytest.mark.parametrize("name", ["c1_small", "gif", "chem"])
def test_dropin_class_matches_oracle(mods, name):
    """NavBySceneFamiliarity (reference API) replaying the device log, driven the
    way scripts/run_experiment.py:235-258 drives it."""
    navsim, util, O = mods
    L, w, tpath, pose, frames = build_case(name)
    frames = min(frames, 120)
    kw = dict(w)
    cw = kw.pop("chem_weight", 0.0)
    nsf = navsim.NavBySceneFamiliarity(L, familiarity_model=navsim.sads_familiarity(cw), **kw)
    nsf.train_from_path(tpath)
    with pytest.raises(ValueError):
        nsf.train_from_path(tpath)
    nsf.position = (pose[0], pose[1])
    nsf.angle = pose[2]
    ow = O.World(L, **w)
    ow.train_from_path(tpath)
    assert np.array_equal(nsf.familiar_scenes, ow.scenes)
    ag = ow.new_agent(*pose)
    r = ow.run(ag, frames, log_afam=True)
    status, done = 0, 0
    try:
        for f in range(frames):
            nsf.step_forward()
            assert np.allclose(nsf.angle_familiarity, r["afam"][f], rtol=FAM_RTOL, atol=0)
            assert np.isclose(nsf.step_familiarity, r["afam"][f].max(), rtol=FAM_RTOL, atol=0)
            done += 1
    except navsim.StopNavigationException as e:
        status = e.get_code()
    assert (status, done) == (r["status"], r["completed"])
    assert nsf.navigated_for_frames == ag.navigated_for_frames
    assert nsf._n_navigation_error == ag.n_nav_err
    assert np.isclose(nsf._navigation_error, ag.nav_err, rtol=1e-12)
    assert np.array_equal(nsf._coverage_array, ag._cov.astype(bool))
    assert np.allclose(nsf.position, (ag.x, ag.y), rtol=0, atol=POS_TOL)
    # scene_familiarity (plotting only) is evaluated on demand for the last step
    sf = nsf.scene_familiarity
    assert sf.shape == (len(tpath),) and np.all(np.isfinite(sf))
    # get_sensor_mat keeps the reference's exceptions
    with pytest.raises(navsim.OutOfLandscapeBoundsException):
        nsf.get_sensor_mat((1.0, 1.0), 0.0)


@pytest.mark.parametrize("buffers", ["pinned", "pageable", "fresh"])
@pytest.mark.parametrize("name", ["c1_small", "ties"])
def test_host_driven_step_io_matches_oracle(mods, name, buffers):
    """nvb_agents_step_io: the host hands in the poses of every step and reads the results back
    (bench.py's end-to-end form), with pinned buffers, pageable buffers and buffers that
    change on every call.  All must walk the oracle's trajectories."""
    import torch
    navsim, util, O = mods
    L, w, tpath, pose, frames = build_case(name)
    K = 14
    eng = navsim.NavEngine(L, **w)
    ow = O.World(L, **w)
    assert eng.train_from_path(tpath) == (0, -1)
    assert ow.train_from_path(tpath) == (0, -1)
    poses = np.vstack([np.asarray(pose)[None], agent_grid(tpath, w)])[:12]
    B = len(poses)
    eng.set_agents(poses, K + 5)

    def alloc():
        if buffers == "pinned":
            t = (torch.empty((B, 3), dtype=torch.float64).pin_memory(), torch.empty((B, 3), dtype=torch.float64).pin_memory(),
                 torch.empty((B,), dtype=torch.int16).pin_memory(), torch.empty((B,), dtype=torch.float64).pin_memory())
            return t, [x.numpy() for x in t]
        a = [np.empty((B, 3)), np.empty((B, 3)), np.empty(B, np.int16), np.empty(B)]
        return a, a
    keep, (h_in, h_pose, h_best, h_fam) = alloc()
    h_in[:] = poses
    best, pos, fam = [], [], []
    for i in range(K):
        if buffers == "fresh":   # new arrays on every call: the pointers never repeat
            cur = h_in.copy()
            keep, (h_in, h_pose, h_best, h_fam) = alloc()
            h_in[:] = cur
        eng.step_io(h_in, 1, h_best, h_pose, h_fam)
        best.append(h_best.copy()); pos.append(h_pose.copy()); fam.append(h_fam.copy())
        h_in[:] = h_pose
    best, pos, fam = np.array(best), np.array(pos), np.array(fam)
    for b, p in enumerate(poses):
        ag = ow.new_agent(*p)
        r = ow.run(ag, K, log_afam=True)
        n = r["completed"] + (1 if r["status"] in (1, -1) else 0)
        assert n > 0
        assert np.array_equal(best[:n, b], r["best_idx"][:n]), (name, buffers, b)
        assert np.all(best[n:, b] == -1)
        assert np.allclose(pos[:n, b], r["pos"][:n], rtol=0, atol=POS_TOL)
        assert np.allclose(fam[:n, b], r["afam"][:n].max(axis=1), rtol=FAM_RTOL, atol=0)
    st = eng.state()
    assert eng.steps_done == K


FAM_RTOL_CHEM = 1e-6   # angle_familiarity with chem_weight > 0, see below (north_star: FP32-level)


@pytest.mark.parametrize("cw", [0.123, 0.3, 0.77])
def test_non_dyadic_chem_weight(mods, cw):
    """chem_weight > 0: the distance kernel ranks views by floor(4096 * f) and a heading that
    is NOT tied with another heading takes its exact FP64 value from the lowest-index view at
    that minimum; another view in the same 1/4096 bucket can be lower by up to 1/(4096 * 255),
    so angle_familiarity of such headings is exact only to ~1e-8 relative (asserted: 1e-6;
    most values are bit-exact).  Headings tied within the band are resolved over every view in
    exact FP64, so heading sequences, positions, stops and coverage are exact all the same."""
    navsim, util, O = mods
    L, w, tpath, pose, frames = build_case("chem")
    w = dict(w)
    w["chem_weight"] = cw
    frames = min(frames, 100)
    eng = navsim.NavEngine(L, **w)
    ow = O.World(L, **w)
    assert eng.train_from_path(tpath) == (0, -1) and ow.train_from_path(tpath) == (0, -1)
    poses = np.vstack([np.asarray(pose)[None], agent_grid(tpath, w)])
    eng.set_agents(poses, frames)
    eng.step(frames, log_afam=True)
    log = eng.log(0, frames, afam=True)
    st = eng.state()
    n_exact = n_vals = 0
    for b, p in enumerate(poses):
        ag = ow.new_agent(*p)
        r = ow.run(ag, frames, log_afam=True)
        n = r["completed"] + (1 if r["status"] in (1, -1) else 0)
        assert st["status"][b] == r["status"] and st["completed"][b] == r["completed"]
        assert np.array_equal(log["best_idx"][:n, b], r["best_idx"][:n]), (cw, b)
        assert np.array_equal(log["poses"][:n, b], r["pos"][:n])
        assert np.allclose(log["afam"][:n, b], r["afam"][:n], rtol=FAM_RTOL_CHEM, atol=0)
        assert np.allclose(log["step_fam"][:n, b], r["afam"][:n].max(axis=1), rtol=FAM_RTOL_CHEM, atol=0)
        assert np.array_equal(st["coverage"][b], ag._cov)
        n_exact += int(np.sum(log["afam"][:n, b] == r["afam"][:n]))
        n_vals += n * r["afam"].shape[1]
    assert n_exact > 0.9 * n_vals
