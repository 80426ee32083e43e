"""The device build of csrc/glibc_trig.cuh against the host libm on 10^7 arguments."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_sincos_equals_host_libm(gpu):
    import navsim
    eng = navsim.NavEngine(np.zeros((64, 64, 3), np.uint8), (8, 2), 1.0, n_test_angles=4, sensor_pixel_dimensions=(2, 2))
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-math.pi, 2 * math.pi, 10_000_000), rng.uniform(-0.13, 0.13, 200_000),
                        rng.uniform(0.85, 0.86, 200_000), rng.uniform(2.42, 2.43, 200_000),
                        rng.uniform(-1e5, 1e5, 400_000),
                        [0.0, -0.0, 1e-300, 2.0 ** -26, 2.0 ** -27, 0.126, math.pi / 2, math.pi, 2 * math.pi]])
    s, c = eng.device_sincos(x)
    assert np.array_equal(s, np.sin(x)) and np.array_equal(c, np.cos(x))
