"""Per-CTA checkpoints of the single-launch step kernel (k3_step_tm) on the bench workload:
percentiles over the 1024 agent CTAs of the time from the step-batch's first stamp to
  resident | dependency met | state + tile minima in | pose known | update_error done |
  all warps at the gather | window landed | done.   (tuning aid, not a bench arm)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200")):
    sys.path.insert(0, p)
import numpy as np

import bench
from navsim import NavEngine

L, tpath, poses, kw = bench.build_world_inputs(bench.WORKLOAD)
eng = NavEngine(L, **kw)
assert eng.train_from_path(tpath) == (0, -1)
eng.set_agents(poses)
eng.step(20)
eng.sync()
raw = eng.timeline(8, raw=True).astype(np.float64)
B = len(poses)
t0 = raw[0][raw[0] > 0].min()
names = [("k2 resident", 0, 0), ("k2 dep", 0, 1), ("k2 done", 0, 2), ("tm resident", 3, 0), ("tm dep met", 3, 1),
         ("tm inputs in", 4, 0), ("tm pose known", 4, 1), ("tm update_error done", 4, 2), ("tm at gather", 5, 0),
         ("tm window landed", 5, 1), ("tm done", 3, 2)]
for name, k, w in names:
    v = raw[k, :B if k else 148, w]
    v = v[v > 0]
    if len(v) == 0:
        print("%-24s -" % name)
        continue
    v = (v - t0) / 1e3
    print("%-24s n=%4d  min %6.2f  p10 %6.2f  p50 %6.2f  p90 %6.2f  p99 %6.2f  max %6.2f us"
          % (name, len(v), v.min(), *np.percentile(v, [10, 50, 90, 99]), v.max()))
d = (raw[3, :B, 2] - raw[3, :B, 1]) / 1e3
print("per CTA, dependency met -> done: p50 %.2f  p90 %.2f  max %.2f us" % (np.percentile(d, 50), np.percentile(d, 90), d.max()))

# the late agents: pose known more than 2 us after the median (agents with headings tied at the
# step's integer minimum: they compare every view at the minimum in FP64 first)
pose = (raw[4, :B, 1] - t0) / 1e3
late = np.nonzero(pose > np.median(pose) + 2.0)[0]
print("late agents: %d of %d; pose known at %s us" % (len(late), B, ", ".join("%.1f" % pose[bb] for bb in late[:16])))
