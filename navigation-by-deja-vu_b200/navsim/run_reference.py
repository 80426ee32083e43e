"""Runs one of the reference's UNMODIFIED driver scripts on top of this package:

    python -m navsim.run_reference /path/to/scripts/run_experiment.py trials.json landscape_dir/
    torchrun --nproc-per-node 8 ... -m navsim.run_reference /path/to/scripts/run_experiment.py ...

The script's own `from navsim import NavBySceneFamiliarity, StopNavigationException,
sads_familiarity` (scripts/run_experiment.py:84) resolves to this package (hot path on the
GPU behind the C ABI); what the script needs and the environment lacks (mpi4py, scikit-image,
matplotlib, pre-1.24 NumPy aliases, a ragged-safe np.array_split) comes from navsim.compat.
Under torchrun / mpirun every process takes its rank's slice of the trials exactly as the
script's np.array_split does, and its GPU (LOCAL_RANK)."""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    pkg_parent = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if pkg_parent not in sys.path:
        sys.path.insert(0, pkg_parent)
    from navsim import compat
    compat.install()
    script = argv[0]
    sys.argv = argv
    sys.path.insert(1, os.path.dirname(os.path.abspath(script)))
    if script.endswith(".so"):
        run_compiled_as_main(script)
    else:
        runpy.run_path(script, run_name="__main__")


def run_compiled_as_main(path):
    """The same for a driver compiled to an extension module (how the unmodified script travels
    to a machine without the reference's sources): its module body runs with
    __name__ == "__main__", so the script's own main block executes."""
    import importlib.machinery
    import importlib.util
    stem = os.path.basename(path).split(".")[0]
    loader = importlib.machinery.ExtensionFileLoader(stem, path)
    spec = importlib.util.spec_from_loader(stem, loader)
    mod = loader.create_module(spec)
    mod.__name__ = "__main__"
    loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    main()
