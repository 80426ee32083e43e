"""The tensor-core distance kernel (csrc/distance_tc.cuh: tcgen05 int8, exact thermometer form
of the sum of absolute differences) against the oracle's integer sums and against the byte-SIMD
kernel: minimum and LOWEST view index, exact, for every level count it accepts, odd shapes,
ragged tile edges, duplicate views, masked columns; and the fall-back when a host-supplied
library holds values the sensor's quantisation cannot produce."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _levels(n):
    from navsim import _cabi
    return np.unique(np.concatenate([[0], _cabi.quant_lut(n)])).astype(np.uint8)


@pytest.mark.parametrize("n_levels", [2, 3, 4, 5, 8, 9])
@pytest.mark.parametrize("W,H,G,N", [(40, 2, 128, 256), (40, 2, 300, 1414), (20, 4, 129, 257), (8, 2, 1000, 255),
                                     (50, 5, 200, 700), (64, 64, 130, 300)])
def test_tc_equals_simd_equals_oracle(gpu, n_levels, W, H, G, N):
    import navsim
    from oracle import oracle as O
    if W * H > 1000 and n_levels not in (5, 9):
        pytest.skip("large sensor: two level counts are enough")
    rng = np.random.default_rng(n_levels * 100003 + W * 1009 + G)
    L = np.zeros((64, 64, 3), np.uint8)
    eng = navsim.NavEngine(L, (W, H), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2),
                           n_sensor_levels=n_levels)
    lv = _levels(n_levels)
    assert eng.tc_planes > 0
    scenes = np.zeros((N, H, W, 3), np.uint8)
    scenes[..., 2] = lv[rng.integers(0, len(lv), (N, H, W))]
    scenes[N // 2] = scenes[3]           # duplicate views: the lower index must win
    scenes[N - 1] = scenes[3]
    scenes[:, :, W // 2 - 1:W // 2 + 1, :] = 0   # a masked centre column pair (value 0 on both sides)
    eng.set_library(scenes)
    q = np.zeros((G, H, W, 3), np.uint8)
    src = rng.integers(0, N, G)
    q[..., 2] = scenes[src][..., 2]
    flip = rng.random((G, H, W)) < 0.2
    q[..., 2][flip] = lv[rng.integers(0, len(lv), int(flip.sum()))]
    q[0] = scenes[3]
    md_tc, vi_tc = eng.familiarity_min(q)
    eng.set_distance_kernel(simd_only=True)
    md_simd, vi_simd = eng.familiarity_min(q)
    eng.set_distance_kernel(simd_only=False)
    assert np.array_equal(md_tc, md_simd) and np.array_equal(vi_tc, vi_simd)
    for g in rng.choice(G, size=min(G, 24), replace=False):
        _, vt = O.sad_int(scenes, q[g])
        assert md_tc[g] == vt.min() and vi_tc[g] == int(np.argmin(vt))
    assert md_tc[0] == 0 and vi_tc[0] == 3


def test_values_outside_the_levels_fall_back(gpu):
    """A library or a query uploaded from the host may hold any bytes: the planes cannot
    express them, the byte-SIMD kernel answers instead (same API, exact)."""
    import navsim
    from oracle import oracle as O
    rng = np.random.default_rng(9)
    W, H, N, G = 40, 2, 500, 200
    L = np.zeros((64, 64, 3), np.uint8)
    eng = navsim.NavEngine(L, (W, H), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2), n_sensor_levels=5)
    scenes = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)     # not level values
    eng.set_library(scenes)
    q = rng.integers(0, 256, (G, H, W, 3), dtype=np.uint8)
    md, vi = eng.familiarity_min(q)
    for g in range(0, G, 17):
        _, vt = O.sad_int(scenes, q[g])
        assert md[g] == vt.min() and vi[g] == int(np.argmin(vt))
    # level-valued library, one query with a stray value
    lv = _levels(5)
    scenes2 = np.zeros((N, H, W, 3), np.uint8)
    scenes2[..., 2] = lv[rng.integers(0, 5, (N, H, W))]
    eng.set_library(scenes2)
    q2 = scenes2[rng.integers(0, N, G)].copy()
    q2[5, 0, 7, 2] = 17
    md, vi = eng.familiarity_min(q2)
    for g in (0, 5, 6, G - 1):
        _, vt = O.sad_int(scenes2, q2[g])
        assert md[g] == vt.min() and vi[g] == int(np.argmin(vt))


@pytest.mark.parametrize("name", ["c1_small", "gif", "square"])
def test_trajectories_identical_with_either_kernel(gpu, name):
    """The stepping loop with the tensor-core kernel (the sampler writes the operand planes)
    and with the byte-SIMD kernel: same log, bit for bit."""
    import navsim
    from cases import agent_grid, build_case
    L, w, tpath, pose, frames = build_case(name)
    frames = min(frames, 60)
    poses = np.vstack([np.asarray(pose)[None], agent_grid(tpath, w, 4, 4)])
    logs = []
    for simd_only in (False, True):
        eng = navsim.NavEngine(L, **w)
        assert eng.train_from_path(tpath) == (0, -1)
        eng.set_distance_kernel(simd_only=simd_only)
        eng.set_agents(poses, frames)
        assert eng.distance_kernel == ("k2_sad_v" if simd_only else "k2_tc")
        eng.step(frames, log_afam=True)
        logs.append((eng.log(0, frames, afam=True), eng.state()))
    (la, sa), (lb, sb) = logs
    assert np.array_equal(la["best_idx"], lb["best_idx"])
    assert np.array_equal(la["poses"], lb["poses"])
    assert np.array_equal(la["afam"], lb["afam"], equal_nan=True)
    for k in ("status", "completed", "coverage", "err_n", "err_sum"):
        assert np.array_equal(sa[k], sb[k])


@pytest.mark.parametrize("cw", [0.3, 1.0, 0.123])
@pytest.mark.parametrize("W,H,G,N", [(40, 2, 300, 1414), (20, 4, 17, 130), (8, 2, 1000, 255), (50, 5, 64, 700), (64, 64, 40, 150)])
def test_chemistry_kernel_min_argmin(gpu, cw, W, H, G, N):
    """chem_weight > 0 (tiled k2_sad_hsv_t for 16 glimpses and more): the fixed-point score
    floor(4096 * (cw * 0.5 * X + (1 - cw) * V)) of the oracle's integer sums, minimum and lowest
    view index, exact."""
    import navsim
    from oracle import oracle as O
    rng = np.random.default_rng(int(cw * 1000) + W * 7 + G)
    L = np.zeros((64, 64, 3), np.uint8)
    eng = navsim.NavEngine(L, (W, H), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2), chem_weight=cw)
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    scenes = np.zeros((N, H, W, 3), np.uint8)
    scenes[..., 0] = rng.integers(0, 3, (N, H, W)) * 85
    scenes[..., 1] = rng.integers(0, 2, (N, H, W)) * 127
    scenes[..., 2] = levels[rng.integers(0, 5, (N, H, W))]
    scenes[N // 2] = scenes[3]
    scenes[N - 1] = scenes[3]
    eng.set_library(scenes)
    q = scenes[rng.integers(0, N, G)].copy()
    flip = rng.random((G, H, W)) < 0.3
    q[..., 2][flip] = levels[rng.integers(0, 5, int(flip.sum()))]
    q[..., 0][flip] = rng.integers(0, 3, int(flip.sum())) * 85
    q[0] = scenes[3]
    md, vi = eng.familiarity_min(q)
    for g in rng.choice(G, size=min(G, 16), replace=False):
        xt, vt = O.sad_int(scenes, q[g])
        score = np.floor(4096.0 * ((xt.astype(np.float64) * 0.5) * cw + (1.0 - cw) * vt.astype(np.float64)))
        assert md[g] == score.min() / 4096.0, (g, md[g], score.min() / 4096.0)
        assert vi[g] == int(np.argmin(score))
    assert md[0] == 0 and vi[0] == 3


@pytest.mark.parametrize("W,H,G,N", [(40, 2, 10, 20000), (40, 2, 1, 16384), (40, 2, 16, 70001), (24, 2, 7, 33333),
                                     (8, 2, 13, 100003), (56, 2, 10, 16385)])
def test_streaming_kernel_few_glimpses_large_library(gpu, W, H, G, N):
    """k2_stream (csrc/distance.cuh: a handful of glimpses, every warp streams its own range of the
    library): minimum and LOWEST view index against NumPy, ragged ranges, duplicate views in
    different warps' ranges, and the same keys as the tiled kernel."""
    import navsim
    rng = np.random.default_rng(W * 7919 + G * 31 + N)
    L = np.zeros((64, 64, 3), np.uint8)
    eng = navsim.NavEngine(L, (W, H), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2), n_sensor_levels=256)
    scenes = np.zeros((N, H, W, 3), np.uint8)
    scenes[..., 2] = rng.integers(0, 256, (N, H, W))
    q = np.zeros((G, H, W, 3), np.uint8)
    src = rng.integers(0, N, G)
    q[..., 2] = scenes[src][..., 2]
    noise = rng.random((G, H, W)) < 0.3
    q[..., 2][noise] = rng.integers(0, 256, int(noise.sum()))
    # exact duplicates of glimpse 0's best view far apart: the lowest index must win
    for v in (17, N // 3, N - 1):
        scenes[v] = q[0]
    eng.set_library(scenes)
    md, vi = eng.familiarity_min(q)
    lib = scenes[..., 2].reshape(N, -1).astype(np.int32)
    for g in range(G):
        d = np.abs(lib - q[g, ..., 2].reshape(1, -1).astype(np.int32)).sum(axis=1)
        assert md[g] == d.min() and vi[g] == int(np.argmin(d)), (g, md[g], vi[g], d.min(), int(np.argmin(d)))
    assert md[0] == 0 and vi[0] == 17
