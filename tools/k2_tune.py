"""Times the distance kernel alone on the bench workload (tuning aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "navigation-by-deja-vu_b200"))
import numpy as np
import bench
from navsim import NavEngine
L, tpath, poses, kw = bench.build_world_inputs(bench.WORKLOAD)
eng = NavEngine(L, **kw)
assert eng.train_from_path(tpath) == (0, -1)
eng.set_agents(poses)
eng.step(3); eng.sync()
ms = eng.time_distance_kernel(50)
peak = eng.probe_sad_peak(8192)
G, N, P = len(poses) * 10, eng.n_views, 80
print("K2 %.2f us  frac %.3f  (env MG=%s)" % (ms * 1e3, (G * N * P / (ms * 1e-3)) / peak, os.environ.get("NAVSIM_B200_K2_MG")))
