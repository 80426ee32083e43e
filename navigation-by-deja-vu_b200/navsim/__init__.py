"""navsim -- drop-in package for the reference's `navsim` (navsim/__init__.py:1):
`from navsim import NavBySceneFamiliarity, StopNavigationException,
sads_familiarity` (scripts/run_experiment.py:84) resolves here, with the hot
path running on a B200 behind the C ABI in include/navsim_b200.h.
"""
from .NavBySceneFamiliarity import *  # noqa: F401,F403
from .engine import NavEngine  # noqa: F401
