"""Synthetic landscape primitives, importable as navsim.generate_landscapes because the
reference's driver imports image_from_prob_mat from there (scripts/run_experiment.py:86).
Offline data generation, off the hot path (component 14): plain
NumPy / SciPy on the host, same call signatures as navsim/generate_landscapes.py:9-72."""
import random

import numpy as np

from .util import diffuse  # noqa: F401  (navsim/generate_landscapes.py:5)


def random_squares(shape, s, n, value=1):
    """n squares of side s at random places (navsim/generate_landscapes.py:9-19)."""
    assert s % 2 == 0, "Side length must be even"
    mat = np.zeros(shape=shape, dtype=int)
    h = s // 2
    for _ in range(n):
        x, y = random.randrange(0, shape[0]), random.randrange(0, shape[1])
        mat[x - h:x + h, y - h:y + h] = value
    return mat


def random_squares_rot(shape, s, n):
    """As random_squares with randomly rotated squares (navsim/generate_landscapes.py:21-40)."""
    from scipy import ndimage
    square = np.ones((s, s))
    squares = [(ndimage.rotate(square, ang, reshape=True, order=1) > 0.5).astype(int)
               for ang in np.linspace(0., 360., 120)]
    mat = np.zeros(shape=tuple(e + 4 * s for e in shape), dtype=int)
    for _ in range(n):
        x, y = random.randrange(2 * s, shape[0] + 2 * s), random.randrange(2 * s, shape[1] + 2 * s)
        sq = random.choice(squares)
        a, b = sq.shape[0] // 2, sq.shape[1] // 2
        mat[x - a:x + (sq.shape[0] - a), y - b:y + (sq.shape[1] - b)] += sq
    mat[mat >= 1] = 1
    out = mat[2 * s:-(2 * s), 2 * s:-(2 * s)]
    assert out.shape == tuple(shape)
    return out


def random_matrix_bw_balance(shape, proportion=0.5, threshold=0.06, max_iter=100, func=random_squares, **kwargs):
    """Draw from `func` until the share of set pixels is within `threshold` of `proportion`
    (navsim/generate_landscapes.py:42-58)."""
    assert 0 < threshold < 1 and 0 < proportion < 1
    total = shape[0] * shape[1]
    for _ in range(max_iter):
        mat = func(shape, **kwargs)
        share = np.sum(mat) / total
        if proportion - threshold < share < proportion + threshold:
            return mat
    raise RuntimeError("Couldn't generate a matrix within the desired range")


def checkerboard(shape, checkersize):
    out = np.zeros(shape=(shape, shape))
    for i in range(checkersize):
        for j in range(checkersize):
            out[i::checkersize * 2, j::checkersize * 2] = 1.0
            out[(i + checkersize)::checkersize * 2, (j + checkersize)::checkersize * 2] = 1.0
    return out


def image_from_prob_mat(prob_mat):
    """A 0/1 image drawn pixel by pixel from a probability matrix (navsim/generate_landscapes.py:66-72)."""
    rand = np.random.random(size=prob_mat.shape)
    out = np.zeros(shape=prob_mat.shape)
    out[rand < prob_mat] = 1
    return out
