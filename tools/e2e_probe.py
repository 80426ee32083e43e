"""Per-call time of the host-driven step (nvb_agents_step_io, zero-copy graph) on the bench
workload, for whatever NAVSIM_B200_* knobs the environment sets.  (tuning aid)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import bench
from navsim import NavEngine

L, tpath, poses, kw = bench.build_world_inputs(bench.WORKLOAD)
eng = NavEngine(L, **kw)
assert eng.train_from_path(tpath) == (0, -1)
B = len(poses)
h_in = torch.empty((B, 3), dtype=torch.float64).pin_memory()
h_pose = torch.empty((B, 3), dtype=torch.float64).pin_memory()
h_best = torch.empty((B,), dtype=torch.int16).pin_memory()
h_fam = torch.empty((B,), dtype=torch.float64).pin_memory()
a_in, a_pose = h_in.numpy(), h_pose.numpy()
step_io = eng.bind_step_io(a_in, h_best.numpy(), a_pose, h_fam.numpy())
eng.set_agents(poses)
a_in[:] = poses
for _ in range(10):
    step_io()
    a_in[:] = a_pose
n0 = eng.launch_count
t0 = time.perf_counter()
for _ in range(80):
    step_io()
    a_in[:] = a_pose
dt = time.perf_counter() - t0
print("step_io: %.1f us per call, %.1f launches per call, kernel %s, env %s"
      % (dt / 80 * 1e6, (eng.launch_count - n0) / 80, eng.distance_kernel,
         {k: v for k, v in os.environ.items() if k.startswith("NAVSIM_B200")}))
