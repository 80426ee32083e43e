"""Pins the oracle (oracle/navsim_oracle.c) to the reference's own code where the
compiled reference (oracle/_ref, built in this container by oracle/build_ref.py)
is available: randomised inputs beyond the committed golden vectors.  CPU only."""
import warnings

import numpy as np
import pytest

from cases import CASES, build_case
from oracle import oracle as O, ref_loader

ref = ref_loader.load_reference()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (needs /root/reference)")


def test_fill_sensor_random():
    from navsim import synthetic
    L = synthetic.make_landscape(21, 300, n_chemicals=3)
    rng = np.random.default_rng(5)
    for land in (L, L[::-1], L[:, ::-1]):
        for _ in range(60):
            Hpx, Wpx = int(rng.choice([2, 4, 8, 16, 64])), int(rng.choice([2, 8, 16, 40, 80]))
            x, y, ang = rng.uniform(60, 240), rng.uniform(60, 240), rng.uniform(-7, 7)
            a = np.zeros((Hpx, Wpx, 3), np.uint8)
            b = np.zeros_like(a)
            ref.util.fill_sensor_from(a, x, y, ang, land)
            assert O.fill_sensor(b, x, y, ang, land) == 0
            assert np.array_equal(a, b)
    with pytest.raises(IndexError):
        ref.util.fill_sensor_from(np.zeros((64, 64, 3), np.uint8), 267., 267., 0.7, L)
    assert O.fill_sensor(np.zeros((64, 64, 3), np.uint8), 267., 267., 0.7, L) == O.INDEX_ERROR


def test_downscale_random():
    rng = np.random.default_rng(6)
    for t in range(80):
        fr, fc = int(rng.choice([1, 2, 3, 4])), int(rng.choice([1, 2, 4, 5]))
        img = rng.integers(0, 256, (fr * int(rng.integers(1, 6)), fc * int(rng.integers(1, 9)), 3), dtype=np.uint8)
        if t % 2:
            img[..., 0] = rng.integers(0, 3, img.shape[:2]) * 85
        assert np.array_equal(ref.util.downscale_chem(img, fr, fc), O.downscale_chem(img, fr, fc))


@pytest.mark.parametrize("name", ["c1_small", "chem", "ties", "chem1"])
def test_full_trajectory(name):
    warnings.filterwarnings("ignore")
    L, w, tpath, pose, frames = build_case(name)
    frames = min(frames, 120)
    kw = dict(w)
    cw = kw.pop("chem_weight", 0.0)
    nsf = ref.NavBySceneFamiliarity(L, familiarity_model=ref.util.sads_familiarity(cw), **kw)
    nsf.train_from_path(tpath)
    ow = O.World(L, **w)
    assert ow.train_from_path(tpath) == (0, -1)
    assert np.array_equal(ow.scenes, nsf.familiar_scenes)
    nsf.position = (pose[0], pose[1])
    nsf.angle = pose[2]
    ag = ow.new_agent(*pose)
    r = ow.run(ag, frames, log_afam=True)
    status, done = 0, 0
    try:
        for f in range(frames):
            nsf.step_forward()
            assert np.array_equal(nsf.angle_familiarity, r["afam"][f])
            assert tuple(nsf.position) + (nsf.angle,) == tuple(r["pos"][f])
            done += 1
    except ref.StopNavigationException as e:
        status = e.get_code()
    assert (status, done) == (r["status"], r["completed"])
    assert nsf._navigation_error == ag.nav_err and nsf._n_navigation_error == ag.n_nav_err
    assert np.array_equal(nsf._coverage_array, ag._cov.astype(bool))
