"""Build the reference's own hot-path code into oracle/_ref/ (binary only).

TEST INFRASTRUCTURE ONLY.  Compiles, from the sources where they lie under
/root/reference, with Cython + gcc:

  navsim/util.pyx                  -> oracle/_ref/navsim/util.<abi>.so
  navsim/NavBySceneFamiliarity.py  -> oracle/_ref/navsim/NavBySceneFamiliarity.<abi>.so
  scripts/run_experiment.py        -> oracle/_ref/scripts/run_experiment.<abi>.so    (byte-identical)
  scripts/load_experiments.py      -> oracle/_ref/scripts/load_experiments.<abi>.so  (byte-identical)

The two driver scripts are compiled so that the reference's UNMODIFIED entry points can be
executed on the GPU box (where /root/reference does not exist) on top of the product package:
tests/test_gpu_reference_driver.py loads run_experiment's module body as __main__ through
navsim.run_reference.

No reference source enters the repository: the two files are copied to a
scratch directory under /tmp, built there, and only the shared objects are
kept (oracle/_ref/ is git-ignored, but travels to the GPU box).  The single
edit is the three-token NumPy-2 shim SURVEY.md section 8(c) documents
(`np.int_t` -> `long` at util.pyx:77,82,97; that typedef no longer exists in
NumPy 2's .pxd).  NavBySceneFamiliarity.py is compiled byte-identical; its
matplotlib / skimage imports and pre-1.24 NumPy aliases are satisfied at load
time by oracle/ref_loader.py.

Run:  python -m oracle.build_ref     (needs /root/reference; a no-op elsewhere)
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("NAVSIM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "navsim")
OUT_SCRIPTS = os.path.join(HERE, "_ref", "scripts")

_SETUP = """
from setuptools import setup
from Cython.Build import cythonize
import numpy as np
setup(ext_modules=cythonize(["navsim/util.pyx", "navsim/NavBySceneFamiliarity.py",
                             "scripts/run_experiment.py", "scripts/load_experiments.py"],
                            language_level=3),
      include_dirs=[np.get_include()], script_args=["build_ext", "--inplace"])
"""


def have_ref():
    return (len(glob.glob(os.path.join(OUT, "util.*.so"))) > 0
            and len(glob.glob(os.path.join(OUT, "NavBySceneFamiliarity.*.so"))) > 0
            and len(glob.glob(os.path.join(OUT_SCRIPTS, "run_experiment.*.so"))) > 0
            and len(glob.glob(os.path.join(OUT_SCRIPTS, "load_experiments.*.so"))) > 0)


def build(force=False):
    if have_ref() and not force:
        return True
    src = os.path.join(REF_ROOT, "navsim")
    if not os.path.isfile(os.path.join(src, "util.pyx")):
        return False
    tmp = tempfile.mkdtemp(prefix="navsim_ref_build_")
    try:
        pkg = os.path.join(tmp, "navsim")
        os.makedirs(pkg)
        with open(os.path.join(src, "util.pyx")) as f:
            pyx = f.read()
        assert pyx.count("np.int_t") == 3
        with open(os.path.join(pkg, "util.pyx"), "w") as f:
            f.write(pyx.replace("np.int_t", "long"))
        shutil.copy(os.path.join(src, "NavBySceneFamiliarity.py"), pkg)
        open(os.path.join(pkg, "__init__.py"), "w").close()
        scr = os.path.join(tmp, "scripts")
        os.makedirs(scr)
        for name in ("run_experiment.py", "load_experiments.py"):
            shutil.copy(os.path.join(REF_ROOT, "scripts", name), scr)
        with open(os.path.join(tmp, "setup_ref.py"), "w") as f:
            f.write(_SETUP)
        subprocess.check_call([sys.executable, "setup_ref.py"], cwd=tmp,
                              stdout=subprocess.DEVNULL)
        os.makedirs(OUT, exist_ok=True)
        for so in glob.glob(os.path.join(pkg, "*.so")):
            shutil.copy(so, OUT)
        os.makedirs(OUT_SCRIPTS, exist_ok=True)
        # (scripts/ is no package: build_ext --inplace leaves these two next to setup_ref.py)
        for so in glob.glob(os.path.join(scr, "*.so")) + glob.glob(os.path.join(tmp, "*.so")):
            shutil.copy(so, OUT_SCRIPTS)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return have_ref()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "reference sources not found; skipped")
