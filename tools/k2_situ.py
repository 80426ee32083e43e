"""Per-item stamps of the distance kernel (k2_tc_bs) INSIDE the replayed step graph on the bench
workload: where the time between "dependency met" and "done" goes.  Builds a second library with
-DNVB_TC_SITU_STAMPS (tools/k2_situ.py --build, here on the CPU box; it travels with the snapshot)
and prints percentiles over the 148 CTAs.   (tuning aid, not a bench arm)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "navigation-by-deja-vu_b200")
SITU = os.path.join(PKG, "lib", "libnavsim_b200_situ.so")
if "--build" in sys.argv:
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-DNVB_TC_SITU_STAMPS", "-DNVB_DEV_MIN", "-Xcompiler", "-fPIC", "-shared", "-o", SITU,
                           os.path.join(PKG, "csrc", "engine.cu")])
    sys.exit(0)
for p in (ROOT, PKG):
    sys.path.insert(0, p)
import numpy as np

from navsim import _cabi
_cabi.LIB_PATH = SITU
import bench
from navsim import NavEngine

L, tpath, poses, kw = bench.build_world_inputs(bench.WORKLOAD)
eng = NavEngine(L, **kw)
assert eng.train_from_path(tpath) == (0, -1)
eng.set_agents(poses)
eng.step(24)
eng.sync()
raw = eng.timeline(16, raw=True).astype(np.float64)
flat = raw.reshape(-1)
k2 = raw[0, :148]
t0 = k2[:, 0][k2[:, 0] > 0].min()
st = flat[6144:6144 + 148 * 32].reshape(148, 4, 8)


def line(name, v):
    v = v[v > 0]
    if len(v) == 0:
        print("%-44s -" % name)
        return
    v = (v - t0) / 1e3
    print("%-44s n=%3d  min %6.2f  p10 %6.2f  p50 %6.2f  p90 %6.2f  max %6.2f us"
          % (name, len(v), v.min(), *np.percentile(v, [10, 50, 90]), v.max()))


line("CTA resident", k2[:, 0])
line("dependency met (producer)", st[:, 0, 3])
for it in range(4):
    line("item %d: inputs landed (MMA warp)" % it, st[:, it, 0])
    line("item %d: MMAs issued" % it, st[:, it, 1])
    line("item %d: commit issued" % it, st[:, it, 2])
    line("item %d: epilogue woke (accumulator done)" % it, st[:, it, 4])
    line("item %d: accumulator released" % it, st[:, it, 5])
    line("item %d: folded" % it, st[:, it, 6])
line("CTA done", k2[:, 2])
tm = raw[3, :len(poses)]
line("k3_step_tm dependency met", tm[:, 1])
