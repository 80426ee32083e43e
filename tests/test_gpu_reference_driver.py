"""The reference's UNMODIFIED entry points on top of the product package (north_star: "the
scripts/run_experiment.py / load_experiments.py entry points stay unchanged"):

* scripts/run_experiment.py, compiled byte-identical into oracle/_ref/scripts (the sources do
  not exist on the GPU box), executed as __main__ through navsim.run_reference with the
  product's navsim package: its task-0.csv must equal, character for character, the CSV the
  same driver wrote on the reference's own navsim code (tests/golden/driver/task-0.csv,
  tests/golden/make_driver_golden.py);
* scripts/load_experiments.load_runs reads that directory;
* the batched driver (python -m navsim.experiments, same command line and schema) writes the
  same rows.
"""
import glob
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "navigation-by-deja-vu_b200")
GOLD = os.path.join(ROOT, "tests", "golden", "driver")


def _compiled(stem):
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "scripts", stem + ".*.so"))
    if not hits:
        pytest.skip("oracle/_ref/scripts not built (reference sources absent when build() ran)")
    return hits[0]


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = PKG + os.pathsep + env.get("PYTHONPATH", "")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    return env


def test_unmodified_driver_on_product_matches_reference_csv(gpu, tmp_path):
    so = _compiled("run_experiment")
    r = subprocess.run([sys.executable, "-m", "navsim.run_reference", so, os.path.join(GOLD, "trials.json"),
                        os.path.join(GOLD, "landscapes")], cwd=str(tmp_path), env=_env(), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = glob.glob(os.path.join(str(tmp_path), "output-trials-*-run0", "task-0.csv"))
    assert len(out) == 1
    got, want = open(out[0]).read(), open(os.path.join(GOLD, "task-0.csv")).read()
    assert got == want
    # the reference's loader (compiled, unmodified) reads the run directory
    sys.path.insert(0, PKG)
    from navsim import compat, run_reference
    compat.install()
    import importlib.machinery
    import importlib.util
    lso = _compiled("load_experiments")
    loader = importlib.machinery.ExtensionFileLoader("load_experiments", lso)
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("load_experiments", loader))
    loader.exec_module(mod)
    data = mod.load_runs([os.path.dirname(out[0])])
    assert len(data["stop_status"]) == 32 and data["sensor_dimensions"].shape == (32, 4)
    assert set(np.unique(data["stop_status"])) <= {1, 0, -1, -2}


def test_batched_driver_cli_writes_the_same_rows(gpu, tmp_path):
    out = str(tmp_path / "run")
    r = subprocess.run([sys.executable, "-m", "navsim.experiments", os.path.join(GOLD, "trials.json"),
                        os.path.join(GOLD, "landscapes"), "--outdir", out, "--workers", "3"],
                       cwd=str(tmp_path), env=_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = open(os.path.join(out, "task-0.csv")).read().splitlines()
    want = open(os.path.join(GOLD, "task-0.csv")).read().splitlines()
    assert got[0] == want[0]
    assert sorted(got[1:]) == sorted(want[1:])
    assert got == want            # one rank: same order as the reference's trial order
