"""Environment shims that let the reference's UNMODIFIED drivers run on top of this package
(SURVEY.md 7.3 H7, 8(f) N1): scripts/run_experiment.py and scripts/load_experiments.py are
byte-for-byte the reference's files; what they need and a current environment lacks is
supplied here, before they are imported:

* NumPy aliases removed in NumPy 1.24: np.float / np.int / np.bool / np.product
  (navsim/NavBySceneFamiliarity.py:98,134,206-207; scripts/load_experiments.py:11-20);
* np.array_split on the list of trial tuples (scripts/run_experiment.py:327): the tuples hold
  arrays of different lengths, which NumPy >= 1.24 refuses to turn into one array -- the
  shim splits the list itself, same contiguous parts;
* mpi4py.MPI.COMM_WORLD (scripts/run_experiment.py:91,294-347: rank, size, one bcast of a
  directory name, one barrier) when mpi4py is not installed: rank / size from the launcher's
  environment (torchrun, mpirun, srun), bcast and barrier through files in a rendezvous dir;
* the three scikit-image calls of make_nsf (scripts/run_experiment.py:172-180: rank.modal on a
  0/1 image, measure.label, measure.regionprops(...).equivalent_diameter) on top of
  scipy.ndimage when scikit-image is not installed, and empty matplotlib modules (imported
  by the reference's modules, only used for plotting).

    python -m navsim.run_reference /path/to/scripts/run_experiment.py trials.json landscapes/

runs the unmodified driver, every trial going through this package's NavBySceneFamiliarity.
Nothing here is on the hot path.
"""
import glob
import importlib
import os
import pickle
import sys
import tempfile
import time
import types

import numpy as np


# ---------------------------------------------------------------- NumPy
def numpy_aliases():
    for alias, target in (("float", float), ("int", int), ("bool", bool)):
        if alias not in np.__dict__:
            setattr(np, alias, target)
    if "product" not in np.__dict__:
        np.product = np.prod


def ragged_array_split():
    """np.array_split that also takes a list of ragged tuples (returned as lists of the same
    items, split like np.array_split splits: the first len % n parts get one more)."""
    if getattr(np.array_split, "_navsim_ragged", False):
        return
    orig = np.array_split

    def array_split(ary, indices_or_sections, axis=0):
        if isinstance(ary, (list, tuple)) and isinstance(indices_or_sections, (int, np.integer)) and axis == 0:
            try:
                return orig(ary, indices_or_sections, axis)
            except ValueError:
                n, parts = len(ary), int(indices_or_sections)
                base, rem = divmod(n, parts)
                out, at = [], 0
                for r in range(parts):
                    cnt = base + (1 if r < rem else 0)
                    out.append(list(ary[at:at + cnt]))
                    at += cnt
                return out
        return orig(ary, indices_or_sections, axis)

    array_split._navsim_ragged = True
    np.array_split = array_split


# ---------------------------------------------------------------- MPI
def _env_int(names, default):
    for n in names:
        if n in os.environ:
            return int(os.environ[n])
    return default


class FileComm(object):
    """COMM_WORLD stand-in for the two collectives the reference driver uses."""

    def __init__(self):
        self.rank = _env_int(("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"), 0)
        self.size = _env_int(("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"), 1)
        job = os.environ.get("NAVSIM_MPI_JOB") or os.environ.get("MASTER_PORT") or os.environ.get("SLURM_JOB_ID") or "solo"
        self._dir = os.environ.get("NAVSIM_MPI_DIR") or os.path.join(tempfile.gettempdir(), "navsim_mpi_%s_%d" % (job, os.getppid()))
        self._seq = 0
        if self.size > 1:
            os.makedirs(self._dir, exist_ok=True)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def _wait(self, path, timeout=3600.0):
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > timeout:
                raise RuntimeError("rank %d timed out waiting for %s" % (self.rank, path))
            time.sleep(0.01)

    def bcast(self, obj, root=0):
        self._seq += 1
        if self.size == 1:
            return obj
        path = os.path.join(self._dir, "bcast_%d.pkl" % self._seq)
        if self.rank == root:
            with open(path + ".tmp", "wb") as f:
                pickle.dump(obj, f)
            os.replace(path + ".tmp", path)
            return obj
        self._wait(path)
        with open(path, "rb") as f:
            return pickle.load(f)

    def barrier(self):
        self._seq += 1
        if self.size == 1:
            return
        open(os.path.join(self._dir, "bar_%d_%d" % (self._seq, self.rank)), "w").close()
        for r in range(self.size):
            self._wait(os.path.join(self._dir, "bar_%d_%d" % (self._seq, r)))

    Barrier = barrier


def mpi_stub():
    try:
        importlib.import_module("mpi4py")
        return False
    except ImportError:
        pass
    pkg = types.ModuleType("mpi4py")
    mpi = types.ModuleType("mpi4py.MPI")
    mpi.COMM_WORLD = FileComm()
    pkg.MPI = mpi
    sys.modules["mpi4py"] = pkg
    sys.modules["mpi4py.MPI"] = mpi
    return True


# ---------------------------------------------------------------- scikit-image
class _Region(object):
    __slots__ = ("label", "area")

    def __init__(self, label, area):
        self.label, self.area = int(label), int(area)

    @property
    def equivalent_diameter(self):
        return float(np.sqrt(4.0 * self.area / np.pi))

    equivalent_diameter_area = equivalent_diameter


def label_image(image):
    """skimage.measure.label for a 2-D image: 8-connected components of the non-zero pixels,
    numbered in raster order of their first pixel."""
    from scipy import ndimage
    return ndimage.label(np.asarray(image) != 0, structure=np.ones((3, 3), int))[0].astype(np.int64)   # skimage returns intp


def region_props(labels):
    counts = np.bincount(np.asarray(labels).ravel())
    return [_Region(k, counts[k]) for k in range(1, len(counts)) if counts[k] > 0]


def modal_filter(image, footprint):
    """skimage.filters.rank.modal on a 0/1 uint8 image: the more frequent value inside the
    footprint (clipped at the border); 0 on a tie (the lower histogram bin wins)."""
    from scipy import ndimage
    img = np.asarray(image)
    if img.max(initial=0) > 1:
        raise NotImplementedError("modal filter stand-in handles 0/1 images only (make_nsf thresholds first)")
    fp = np.asarray(footprint, dtype=np.float64)
    ones = ndimage.correlate((img != 0).astype(np.float64), fp, mode="constant", cval=0.0)
    total = ndimage.correlate(np.ones(img.shape, np.float64), fp, mode="constant", cval=0.0)
    return (ones > total - ones).astype(img.dtype)


def skimage_stub():
    try:
        importlib.import_module("skimage.measure")
        return False
    except ImportError:
        pass
    sk = types.ModuleType("skimage")
    sk.__path__ = []
    measure = types.ModuleType("skimage.measure")
    measure.label = label_image
    measure.regionprops = region_props
    filters = types.ModuleType("skimage.filters")
    filters.__path__ = []
    rank = types.ModuleType("skimage.filters.rank")
    rank.modal = modal_filter
    filters.rank = rank
    transform = types.ModuleType("skimage.transform")

    def rotate(image, angle, resize=False, **kw):
        from scipy import ndimage
        return ndimage.rotate(np.asarray(image, dtype=np.float64), angle, reshape=resize, order=1)
    transform.rotate = rotate
    sk.measure, sk.filters, sk.transform = measure, filters, transform
    for name, mod in (("skimage", sk), ("skimage.measure", measure), ("skimage.filters", filters),
                      ("skimage.filters.rank", rank), ("skimage.transform", transform)):
        sys.modules[name] = mod
    return True


class _Empty(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Empty(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        raise RuntimeError("plotting is not available in this environment (%s)" % self.__name__)


def matplotlib_stub():
    try:
        importlib.import_module("matplotlib")
        return False
    except ImportError:
        pass
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.ticker", "matplotlib.patches",
                 "matplotlib.gridspec", "matplotlib.font_manager", "matplotlib.colors", "mpl_toolkits",
                 "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.anchored_artists",
                 "mpl_toolkits.axes_grid1.inset_locator"):
        sys.modules.setdefault(name, _Empty(name))
    return True


def install():
    """All of the above; returns what had to be stubbed."""
    numpy_aliases()
    ragged_array_split()
    return {"mpi4py": mpi_stub(), "skimage": skimage_stub(), "matplotlib": matplotlib_stub()}
