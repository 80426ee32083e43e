"""Landscape preparation on the device (csrc/landscape.cuh) against the host path the driver
uses (scripts/run_experiment.py:160-199 through navsim.compat's scipy stand-ins for the three
scikit-image calls, and navsim.util.set_HS_where_equal): threshold, modal filter, 8-connected
grain labels in raster order, grain areas, chemistry painting and flips -- all exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _host_labels(V, threshold, min_d):
    from navsim import compat
    m = (V >= threshold).astype(np.uint8)
    w = min_d // 2
    if not w == 0:
        if w % 2 == 0:
            w -= 1
        w = int(w)
        m = compat.modal_filter(m, np.ones((w, w), np.uint8))
    labels = compat.label_image(m)
    return labels, np.array([p.area for p in compat.region_props(labels)], np.int64)


@pytest.mark.parametrize("side,sigma,min_d", [(300, 6.0, 2), (300, 6.0, 8), (517, 2.0, 6), (1000, 3.0, 12), (64, 0.7, 4)])
def test_grain_labels_and_areas(gpu, side, sigma, min_d):
    import navsim
    from navsim import synthetic
    L = synthetic.make_landscape(9000 + side, side, sigma=sigma)
    eng = navsim.NavEngine(L, (8, 2), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2))
    areas = eng.label_grains(200, min_d)
    want_labels, want_areas = _host_labels(L[:, :, 2], 200, min_d)
    assert np.array_equal(areas, want_areas)
    assert np.array_equal(eng.grain_labels(), want_labels)


def test_worst_case_shapes(gpu):
    """Spirals, checkerboards (one 8-connected component), single pixels, everything / nothing set."""
    import navsim
    from navsim import compat
    side = 96
    imgs = []
    yy, xx = np.mgrid[0:side, 0:side]
    imgs.append(((yy + xx) % 2 == 0))                      # checkerboard: diagonally connected
    imgs.append(np.zeros((side, side), bool))
    imgs.append(np.ones((side, side), bool))
    sp = np.zeros((side, side), bool)                      # comb: long thin components meeting at the bottom
    sp[:, ::4] = True
    sp[-1, :] = True
    imgs.append(sp)
    rng = np.random.default_rng(3)
    imgs.append(rng.random((side, side)) < 0.45)           # percolation-like noise
    for im in imgs:
        L = np.zeros((side, side, 3), np.uint8)
        L[:, :, 2] = np.where(im, 255, 0)
        eng = navsim.NavEngine(L, (8, 2), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2))
        areas = eng.label_grains(200, 2)
        want = compat.label_image(im.astype(np.uint8))
        assert np.array_equal(eng.grain_labels(), want)
        assert len(areas) == want.max()


def test_chemistry_and_flips_match_host(gpu):
    import navsim
    from navsim import compat, synthetic, util
    L = synthetic.make_landscape(9100, 400, sigma=5.0)
    eng = navsim.NavEngine(L, (40, 2), 5.0, n_test_angles=10, sensor_pixel_dimensions=(2, 4))
    areas = eng.label_grains(200, 6)
    chems, sats = eng.add_chemistry(areas, n_chemicals=3, min_grain_diameter=6, rng=np.random.default_rng(5))
    # host: the same tables through set_HS_where_equal
    labels, _ = _host_labels(L[:, :, 2], 200, 6)
    want = L.copy()
    util.set_HS_where_equal(labels, want, chems, sats)
    assert np.array_equal(eng.download_landscape(), want)
    assert len(np.unique(want[:, :, 0])) > 1 and np.any(sats == 0) and np.any(sats == 127)
    eng.flip_landscape(vertical=True, horizontal=False)
    assert np.array_equal(eng.download_landscape(), want[::-1])
    eng.flip_landscape(vertical=True, horizontal=True)
    assert np.array_equal(eng.download_landscape(), want[:, ::-1])
    # the prepared device landscape steps like the host-prepared one
    tpath = synthetic.training_path_for(L.shape, 5.0, 10, 0.0)
    ref = navsim.NavEngine(np.ascontiguousarray(want[:, ::-1]), (40, 2), 5.0, n_test_angles=10,
                           sensor_pixel_dimensions=(2, 4), chem_weight=0.3)
    eng.set_world((40, 2), 5.0, n_test_angles=10, sensor_pixel_dimensions=(2, 4), chem_weight=0.3)
    assert eng.train_from_path(tpath) == ref.train_from_path(tpath) == (0, -1)
    assert np.array_equal(eng.familiar_scenes, ref.familiar_scenes)


@pytest.mark.parametrize("side,nstep", [(16, 5), (101, 37), (256, 450), (7, 1), (64, 0)])
def test_diffuse_bit_identical(gpu, side, nstep):
    """navsim.util.diffuse (util.pyx:186-235, offline landscape generation) on the
    device: bit-identical to the compiled reference's Cython loop (where oracle/_ref exists) and
    to the NumPy restatement; same short circuit for nstep == 0."""
    from navsim import util
    rng = np.random.default_rng(side * 1000 + nstep)
    m = rng.random((side, side))
    if side == 101:
        m = (rng.random((side, side)) < 0.3).astype(np.uint8)   # the generator's 0/1 images (generate_landscapes.py)
    got = util.diffuse(m, nstep)
    if nstep == 0:
        assert got is m
        return
    assert got.dtype == np.float64 and got.shape == m.shape
    assert np.array_equal(got, util.diffuse_host(m, nstep))
    try:
        from oracle import ref_loader
        ref = ref_loader.load_reference()
    except Exception:
        ref = None
    if ref is not None:
        assert np.array_equal(got, ref.util.diffuse(m, nstep))
