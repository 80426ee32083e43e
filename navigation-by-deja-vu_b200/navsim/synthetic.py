"""Seeded synthetic inputs for tests and benchmarks.

The reference ships no landscapes, trial files or paths, so parity tests and
bench.py build theirs here.  Landscapes mimic what the reference's offline
tools produce; the training-path construction and the start pose restate
scripts/run_experiment.py:95-124 and :203-229 so that the same (landscape
shape, step size, heading count, curve, start offset) gives the same path and
pose the reference driver would use.  NumPy/SciPy only, no device code.
"""
import numpy as np

LANDSCAPE_THRESHOLD = 200  # scripts/run_experiment.py:30


def make_landscape(seed, side, kind="stitch", n_chemicals=1, grain=16, sigma=3.0):
    """(side, side, 3) uint8 HSV landscape.

    kind="stitch": blurred noise rank-equalised to a flat histogram (what
    autocontrast + equalize in scripts/autostitch.py:126-128 leave behind).
    kind="diffuse": a two-valued image drawn from a blurred random-squares
    probability field (scripts/generate_landscapes.py:66-72 in spirit).
    Chemistry follows scripts/run_experiment.py:126-142,185-193: hue is
    chem_index * (255 // n_chemicals) per grain-sized block, S = 127, and S = 0
    where V is below the labelling threshold.
    """
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(seed)
    if kind == "stitch":
        f = gaussian_filter(rng.random((side, side)), sigma=sigma)
        order = np.argsort(f, axis=None, kind="stable")
        rank = np.empty(side * side, np.float64)
        rank[order] = np.arange(side * side) / float(side * side - 1)
        V = np.uint8(255 * rank.reshape(side, side))
    elif kind == "diffuse":
        field = np.zeros((side, side))
        n_sq = max(4, side // 20)
        for _ in range(n_sq):
            s = int(rng.integers(side // 40 + 2, side // 8 + 3))
            y, x = rng.integers(0, side - s, size=2)
            field[y:y + s, x:x + s] = 1.0
        prob = gaussian_filter(field, sigma=8.0)
        prob = prob / max(prob.max(), 1e-12)
        V = np.where(rng.random((side, side)) < 0.15 + 0.7 * prob, 255, 0).astype(np.uint8)
    else:
        raise ValueError(kind)
    out = np.zeros((side, side, 3), np.uint8)
    out[:, :, 2] = V
    if n_chemicals >= 1:
        crng = np.random.default_rng(seed + 500)
        nb = (side + grain - 1) // grain
        chem = crng.integers(n_chemicals, size=(nb, nb), dtype=np.uint8) * np.uint8(255 // n_chemicals)
        hue = np.kron(chem, np.ones((grain, grain), np.uint8))[:side, :side]
        out[:, :, 0] = hue
        out[:, :, 1] = np.where(V >= LANDSCAPE_THRESHOLD, 127, 0).astype(np.uint8)
        out[:, :, 0][V < LANDSCAPE_THRESHOLD] = 0
    return out


def sin_training_path(curveness, start_x, length, arclen=2.0):
    """Points at ~arclen spacing along y = x - 0.5*l*c*sin(...)
    (scripts/run_experiment.py:95-105)."""
    n_fine = 4 * int(np.floor(length / arclen))
    xs = np.linspace(start_x, start_x + length, n_fine)
    ys = xs - 0.5 * length * curveness * np.sin((xs - 0.5 * length - start_x) * np.pi / (0.5 * length))
    seg = np.sqrt((xs[1:] - xs[:-1]) ** 2 + (ys[1:] - ys[:-1]) ** 2)
    wanted = arclen * np.arange(np.floor(np.sum(seg) / arclen))
    pick = np.searchsorted(np.cumsum(seg), wanted)
    return np.vstack((xs[pick], ys[pick])).T


def chop_path_to_len(path, length):
    """Drop points alternately from the front and the back until the polyline is
    no longer than `length` (scripts/run_experiment.py:107-124)."""
    seg = np.linalg.norm(path[1:] - path[:-1], axis=1)
    assert np.sum(seg) >= length
    lo, hi = 0, len(path)
    for _ in range(len(path)):
        if np.sum(seg[lo:hi]) <= length:
            break
        lo += 1
        if np.sum(seg[lo:hi]) <= length:
            break
        hi -= 1
    assert np.sum(seg[lo:hi]) <= length
    return path[lo:hi]


def training_path_for(landscape_shape, step_size, n_test_angles, curve):
    """The path make_nsf builds (scripts/run_experiment.py:203-213)."""
    side = np.min(landscape_shape[:2])
    margin = 0.2 * side
    path = sin_training_path(curve, margin, side - 2 * margin, arclen=step_size / n_test_angles)
    margin = 0.25 * side
    return np.ascontiguousarray(chop_path_to_len(path, np.sqrt(2 * (side - 2 * margin) ** 2)))


def start_pose(tpath, start_offset, sensor_pixel_width):
    """(x, y, angle) as make_nsf sets them (scripts/run_experiment.py:223-229)."""
    d = tpath[2] - tpath[1]
    angle = np.arctan2(d[1], d[0]) % (2 * np.pi)
    angle = angle + np.deg2rad(start_offset[1])
    off = np.array([np.cos(angle + 0.5 * np.pi) * start_offset[0] * sensor_pixel_width,
                    np.sin(angle + 0.5 * np.pi) * start_offset[0] * sensor_pixel_width])
    pos = tpath[1] + off
    return float(pos[0]), float(pos[1]), float(angle)


def start_pose_grid(tpath, sensor_pixel_width, n_lat=32, n_deg=32, lat=0.5, deg=20.0):
    """C2's agent batch: lateral offsets x heading offsets at tpath[1]."""
    poses = [start_pose(tpath, (la, de), sensor_pixel_width)
             for la in np.linspace(-lat, lat, n_lat) for de in np.linspace(-deg, deg, n_deg)]
    return np.asarray(poses, dtype=np.float64)


def default_frames(tpath, step_size, frame_factor=3.0):
    """run_experiment's frame budget (scripts/run_experiment.py:21,238)."""
    plen = np.sum(np.linalg.norm(tpath[1:] - tpath[:-1], axis=1))
    return int(frame_factor * plen / step_size)
