"""Load the compiled reference (oracle/_ref/, see build_ref.py) without
touching the product's own `navsim` package.

TEST INFRASTRUCTURE ONLY.

    ref = load_reference()         # None if oracle/_ref has not been built
    ref.util.fill_sensor_from(...)
    nsf = ref.NavBySceneFamiliarity(landscape, ...)

The reference module imports matplotlib, mpl_toolkits and skimage at the top
(NavBySceneFamiliarity.py:3-15; plotting only, never on the hot path) and uses
NumPy aliases removed in NumPy 1.24 (np.float :134, np.bool :206-207,
np.product :98).  Both are satisfied here, the way SURVEY.md 8(c) verified:
throw-away stub modules for the duration of the import, and four attribute
aliases on numpy.
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "navsim")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.ticker",
    "matplotlib.patches", "matplotlib.gridspec", "matplotlib.font_manager", "matplotlib.colors",
    "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.anchored_artists",
    "mpl_toolkits.axes_grid1.inset_locator", "skimage", "skimage.transform",
]


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        raise RuntimeError("plotting stub called: %s" % self.__name__)


def numpy_aliases():
    for alias, target in (("float", float), ("int", int), ("bool", bool)):
        if alias not in np.__dict__:
            setattr(np, alias, target)
    if "product" not in np.__dict__:
        np.product = np.prod


def _find(stem):
    hits = glob.glob(os.path.join(REF_DIR, stem + ".*.so"))
    return hits[0] if hits else None


def available():
    return _find("util") is not None and _find("NavBySceneFamiliarity") is not None


_cached = None


def load_reference():
    """Returns a namespace with .util (compiled util.pyx) and everything
    NavBySceneFamiliarity.py defines, or None when oracle/_ref is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    numpy_aliases()
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k == "navsim" or k.startswith("navsim.")}
    for k in saved:
        del sys.modules[k]
    stubbed = []
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
            stubbed.append(name)
    try:
        pkg = types.ModuleType("navsim")
        pkg.__path__ = [REF_DIR]
        sys.modules["navsim"] = pkg
        mods = {}
        for stem in ("util", "NavBySceneFamiliarity"):
            full = "navsim." + stem
            loader = importlib.machinery.ExtensionFileLoader(full, _find(stem))
            spec = importlib.util.spec_from_loader(full, loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            loader.exec_module(mod)
            setattr(pkg, stem, mod)
            mods[stem] = mod
    finally:
        for k in [k for k in sys.modules if k == "navsim" or k.startswith("navsim.")]:
            del sys.modules[k]
        for k, v in saved.items():
            sys.modules[k] = v
        for name in stubbed:
            sys.modules.pop(name, None)
    ns = types.SimpleNamespace(util=mods["util"], module=mods["NavBySceneFamiliarity"])
    for k, v in vars(mods["NavBySceneFamiliarity"]).items():
        if not k.startswith("_"):
            setattr(ns, k, v)
    _cached = ns
    return ns
