"""The alternative step pipelines and staging paths give the same trajectories as the default.

The engine reads its tuning knobs (NAVSIM_B200_STEP_FORM, _NO_PDL, _NO_TMA, ...) once per
process, so every variant runs the trajectory parity tests of test_gpu_parity.py -- oracle
comparison on the small, chemistry and tie-heavy worlds -- in a child process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"NAVSIM_B200_STEP_FORM": "1"},          # decide + ties + move + sample in one launch
    {"NAVSIM_B200_STEP_FORM": "2"},          # cooperative tie queue
    {"NAVSIM_B200_STEP_FORM": "4"},          # tie pass folded into move + sample
    {"NAVSIM_B200_STEP_FORM": "3"},          # K2 | decide | grid-wide tie pass | move + sample (the default is 5: K2 | k3_step_tm)
    {"NAVSIM_B200_NO_TC": "1"},              # byte-SIMD distance kernel everywhere
    {"NAVSIM_B200_NO_PDL": "1"},             # plain stream order, no programmatic dependent launch
    {"NAVSIM_B200_NO_TMA": "1"},             # landscape window gathered from global memory
]


@pytest.mark.gpu
@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
def test_variant_matches_oracle(env):
    child_env = dict(os.environ)
    child_env.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                        "-m", "gpu", "-k", "trajectories or out_of_bounds", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=child_env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
