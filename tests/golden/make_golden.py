"""Generates the golden fixtures in this directory from the REFERENCE's own code.

The reference ships no tests or golden vectors, so the
pins are outputs of its hot path run in the build container: oracle/_ref/ is
navsim/util.pyx (3-token NumPy-2 shim) and navsim/NavBySceneFamiliarity.py
(byte-identical) compiled from /root/reference by oracle/build_ref.py.

    python tests/golden/make_golden.py        # needs /root/reference (or a built oracle/_ref)

Inputs are stored next to the outputs (small landscapes, paths, poses) so the
fixtures are self-contained: the tests never re-generate a landscape.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import build_ref, ref_loader  # noqa: E402

TRAJ_CASES = {
    # name: (landscape kwargs, world kwargs, curve, start offset, chem_weight, max frames)
    "c1": (dict(seed=3001, side=320, sigma=6.0),
           dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=5.0, n_test_angles=10,
                n_sensor_levels=5, max_distance_to_training_path=450), 0.0, (0.1, 5.0), 0.0, 200),
    "chem": (dict(seed=3002, side=320, sigma=6.0, n_chemicals=3),
             dict(sensor_dimensions=(20, 4), sensor_pixel_dimensions=(2, 2), step_size=4.0, n_test_angles=12,
                  n_sensor_levels=8, mask_middle_n=2, max_distance_to_training_path=450), 0.5, (-0.3, -10.0), 0.3, 200),
    "ties": (dict(seed=3003, side=300, sigma=6.0, n_chemicals=2),
             dict(sensor_dimensions=(8, 2), sensor_pixel_dimensions=(2, 2), step_size=2.0, n_test_angles=10,
                  n_sensor_levels=5, max_distance_to_training_path=40), 0.0, (0.5, 20.0), 0.0, 200),
}


def make_primitives(ref, out):
    from navsim import synthetic
    rng = np.random.default_rng(42)
    L = synthetic.make_landscape(4001, 200, n_chemicals=3, sigma=4.0)
    g = dict(landscape=L)
    # A1 fill_sensor_from: in-range poses, a wrap-around pose, flipped landscape
    poses, outs = [], []
    for t in range(24):
        x, y, ang = rng.uniform(50, 150), rng.uniform(50, 150), rng.uniform(-7, 7)
        buf = np.zeros((8, 80, 3), np.uint8)
        ref.util.fill_sensor_from(buf, x, y, ang, L)
        poses.append((x, y, ang))
        outs.append(buf)
    g["fill_poses"], g["fill_out"] = np.array(poses), np.array(outs)
    buf = np.zeros((64, 64, 3), np.uint8)
    ref.util.fill_sensor_from(buf, 33.0, 33.0, 0.7, L)
    g["fill_wrap_pose"], g["fill_wrap_out"] = np.array([33.0, 33.0, 0.7]), buf
    buf = np.zeros((8, 80, 3), np.uint8)
    ref.util.fill_sensor_from(buf, 90.25, 110.5, 2.5, L[::-1, ::-1])
    g["fill_flip_pose"], g["fill_flip_out"] = np.array([90.25, 110.5, 2.5]), buf
    # A2 downscale_chem incl. the integer-division / uint8-wrap quirk (util.pyx:131)
    imgs = rng.integers(0, 256, (12, 8, 16, 3), dtype=np.uint8)
    imgs[::2, :, :, 0] = rng.integers(0, 3, (6, 8, 16)) * 85
    imgs[::3, :, :, 1] = 127
    g["down_in"] = imgs
    g["down_out_2x4"] = np.array([ref.util.downscale_chem(im, 2, 4) for im in imgs])
    g["down_out_4x2"] = np.array([ref.util.downscale_chem(im, 4, 2) for im in imgs])
    g["down_out_3x5"] = np.array([ref.util.downscale_chem(im, 3, 5) for im in imgs])
    # A5 sads_hsv_metric for three chem weights
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    scenes = levels[rng.integers(0, 5, (97, 2, 40, 3))]
    scenes[..., 0] = rng.integers(0, 2, (97, 2, 40)) * 127
    q = scenes[rng.integers(0, 97, 5)].copy()
    q[:, 0, 3] = 17
    g["sads_scenes"], g["sads_queries"] = scenes, q
    for cw in (0.0, 0.3, 1.0):
        func = ref.util.sads_familiarity(cw)(scenes)
        fam = np.empty((len(q), len(scenes)))
        for i in range(len(q)):
            func(q[i], fam[i])
        g["sads_fam_cw%02d" % int(cw * 10)] = fam
    # A3 quantisation tables through the reference's get_sensor_mat arithmetic
    luts = []
    for n in (2, 3, 4, 5, 8, 16, 100, 256):
        rb = np.arange(256, dtype=np.uint8).astype(np.float32)
        rb /= 255
        rb *= (n - 1)
        np.rint(rb, out=rb)
        rb /= (n - 1)
        rb *= 255
        o = np.empty(256, np.uint8)
        o[:] = rb
        luts.append(o)
    g["lut_levels"], g["lut_tables"] = np.array([2, 3, 4, 5, 8, 16, 100, 256]), np.array(luts)
    np.savez_compressed(out, **g)


def make_trajectory(ref, name, out):
    from navsim import synthetic
    lk, wk, curve, off, cw, max_frames = TRAJ_CASES[name]
    L = synthetic.make_landscape(kind="stitch", **lk)
    tpath = synthetic.training_path_for(L.shape, wk["step_size"], wk["n_test_angles"], curve)
    nsf = ref.NavBySceneFamiliarity(L, familiarity_model=ref.util.sads_familiarity(cw), **wk)
    nsf.train_from_path(tpath)
    spw = wk["sensor_dimensions"][0] * wk["sensor_pixel_dimensions"][0]
    pose = synthetic.start_pose(tpath, off, spw)
    nsf.position = (pose[0], pose[1])
    nsf.angle = pose[2]
    frames = min(int(3.0 * nsf.training_path_length / nsf.step_size), max_frames)
    best, afam, pos = [], [], []
    status, done = 0, 0
    try:
        for _ in range(frames):
            nsf.step_forward()
            done += 1
            best.append(int(np.argmax(nsf.angle_familiarity)))
            afam.append(nsf.angle_familiarity.copy())
            pos.append((nsf.position[0], nsf.position[1], nsf.angle))
    except ref.StopNavigationException as e:
        status = e.get_code()
        if status in (1, -1):       # the step moved the agent before raising
            best.append(int(np.argmax(nsf.angle_familiarity)))
            afam.append(nsf.angle_familiarity.copy())
            pos.append((nsf.position[0], nsf.position[1], nsf.angle))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rmsd = nsf.navigation_error
    np.savez_compressed(
        out, landscape=L, tpath=tpath, pose=np.array(pose), frames=frames, chem_weight=cw,
        world_keys=np.array(sorted(wk)), world_vals=np.array([repr(wk[k]) for k in sorted(wk)]),
        familiar_scenes=nsf.familiar_scenes, best_idx=np.array(best, np.int32), afam=np.array(afam),
        pos=np.array(pos), status=status, completed=done,
        navigated_for_frames=nsf.navigated_for_frames, nav_err=nsf._navigation_error,
        n_nav_err=nsf._n_navigation_error, coverage=nsf._coverage_array.astype(np.uint8),
        rmsd=rmsd, path_coverage=nsf.percent_recapitulated,
        percent_forgiving=nsf.percent_recapitulated_forgiving(0.05), n_captures=nsf.n_captures(0.05))
    print("  %s: N=%d frames=%d completed=%d status=%d coverage=%.3f" %
          (name, len(tpath), frames, done, status, nsf.percent_recapitulated))


def main():
    warnings.filterwarnings("ignore")
    if not build_ref.build():
        raise SystemExit("reference sources not found and oracle/_ref not built")
    ref = ref_loader.load_reference()
    make_primitives(ref, os.path.join(HERE, "primitives.npz"))
    for name in TRAJ_CASES:
        make_trajectory(ref, name, os.path.join(HERE, "traj_%s.npz" % name))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
