"""Loading of the golden fixtures written by tests/golden/make_golden.py."""
import ast
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAJ = ["c1", "chem", "ties"]


def primitives():
    return np.load(os.path.join(GOLDEN, "primitives.npz"))


def trajectory(name):
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    world = {str(k): ast.literal_eval(str(v)) for k, v in zip(g["world_keys"], g["world_vals"])}
    return g, world
