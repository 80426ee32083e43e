"""Golden run of the reference's UNMODIFIED experiment driver (scripts/run_experiment.py,
compiled byte-identical into oracle/_ref/scripts by oracle/build_ref.py) on the reference's
own navsim code (oracle/_ref/navsim): writes tests/golden/driver/{landscapes/, trials.json,
task-0.csv}.  tests/test_gpu_reference_driver.py runs the same compiled driver on the product
package and expects the same CSV.

    python tests/golden/make_driver_golden.py        # CPU; needs oracle/_ref (built from /root/reference)

mpi4py, scikit-image and matplotlib are absent here: the driver gets them from
navsim.compat (COMM_WORLD stand-in, scipy.ndimage-based label / regionprops / modal) in both
runs, so the comparison covers everything below those three calls.
"""
import glob
import json
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "navigation-by-deja-vu_b200")
OUT = os.path.join(HERE, "driver")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

TRIALS = {
    "_comment": "driver golden: 2 landscapes x 2 curves x 2 start offsets x 2 chem weights x 2 grain diameters",
    "landscape_class": ["synthetic"],
    "landscape_name": ["a.png", "b.png"],
    "training_path_curve": [0.0, 0.5],
    "sensor_dimensions": [[40, 2, 2, 4]],
    "n_sensor_levels": [5],
    "step_size": [5.0],
    "n_chemicals": [1],
    "chem_weight": [0.0, 0.3],
    "min_chem_grain_diameter": [2, 8],
    "start_offset": [[0.0, 0.0], [0.1, 5.0]],
}


def write_inputs():
    from PIL import Image
    from navsim import synthetic
    os.makedirs(os.path.join(OUT, "landscapes", "synthetic"), exist_ok=True)
    for name, seed, side in (("a.png", 8101, 320), ("b.png", 8102, 300)):
        V = synthetic.make_landscape(seed, side, sigma=6.0)[:, :, 2]
        Image.fromarray(V, mode="L").save(os.path.join(OUT, "landscapes", "synthetic", name))
    with open(os.path.join(OUT, "trials.json"), "w") as f:
        json.dump(TRIALS, f, indent=1)


def reference_navsim_package():
    """`navsim` = the compiled reference (util.pyx + NavBySceneFamiliarity.py) for this process."""
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    pkg = types.ModuleType("navsim")
    pkg.__path__ = []
    for k, v in vars(ref.module).items():
        if not k.startswith("_"):
            setattr(pkg, k, v)
    pkg.sads_familiarity = ref.util.sads_familiarity
    gl = types.ModuleType("navsim.generate_landscapes")
    gl.image_from_prob_mat = lambda prob_mat: (np.random.random(size=prob_mat.shape) < prob_mat).astype(float)
    # NumPy >= 2 drift (NEP 50): `uint8 array * (255 // np.int64(n))` at scripts/run_experiment.py:132
    # now promotes to int64, which util.pyx's typed uint8 arguments (:77-79) refuse; NumPy 1.x kept
    # uint8.  The values fit a byte; cast them back, nothing else changes.
    util = types.ModuleType("navsim.util")
    for k in dir(ref.util):
        if not k.startswith("__"):
            setattr(util, k, getattr(ref.util, k))
    util.set_HS_where_equal = lambda labels, image, H, S: ref.util.set_HS_where_equal(
        labels, image, np.asarray(H, np.uint8), np.asarray(S, np.uint8))
    pkg.util, pkg.generate_landscapes, pkg.NavBySceneFamiliarity_module = util, gl, ref.module
    sys.modules["navsim"] = pkg
    sys.modules["navsim.util"] = util
    sys.modules["navsim.generate_landscapes"] = gl
    sys.modules["navsim.NavBySceneFamiliarity"] = ref.module


def main():
    import warnings
    warnings.filterwarnings("ignore")
    write_inputs()
    from navsim import compat, run_reference          # the product's shims (imported before `navsim` is swapped)
    compat.install()
    reference_navsim_package()
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "scripts", "run_experiment.*.so"))
    if not so:
        raise SystemExit("oracle/_ref/scripts not built: python -m oracle.build_ref")
    cwd = os.getcwd()
    work = os.path.join(OUT, "_work")
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    os.chdir(work)
    try:
        sys.argv = [so[0], os.path.join(OUT, "trials.json"), os.path.join(OUT, "landscapes")]
        run_reference.run_compiled_as_main(so[0])
    finally:
        os.chdir(cwd)
    csv = glob.glob(os.path.join(work, "output-*", "task-0.csv"))
    assert len(csv) == 1, csv
    shutil.copy(csv[0], os.path.join(OUT, "task-0.csv"))
    shutil.rmtree(work, ignore_errors=True)
    print(open(os.path.join(OUT, "task-0.csv")).read())


if __name__ == "__main__":
    main()
