"""chem_weight > 0 on the C2 world (1024 agents, 10 headings, 1414 views, P = 80, three hues):
step time and the distance kernel's (k2_sad_hsv*) time and integer-ALU fraction.
Algorithmic ops per pixel pair for the HSV metric (util.pyx:48-72): hue compare, select,
|dS| or S+S, |dV|, two accumulates = 6.  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200")):
    sys.path.insert(0, p)

import numpy as np

from navsim import NavEngine, synthetic


def main():
    cw = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
    L = synthetic.make_landscape(2001, 2000, "stitch", sigma=6.0, n_chemicals=3)
    kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=10.0, n_test_angles=10,
              n_sensor_levels=5, max_distance_to_training_path=450.0, chem_weight=cw)
    tpath = synthetic.training_path_for(L.shape, 10.0, 10, 0.0)
    eng = NavEngine(L, **kw)
    assert eng.train_from_path(tpath) == (0, -1)
    poses = synthetic.start_pose_grid(tpath, 80, n_lat=32, n_deg=32)
    eng.set_agents(poses)
    eng.step(10)
    eng.sync()
    eng.rewind()
    eng.sync()
    t0 = time.perf_counter()
    eng.step(100)
    eng.sync()
    dt = (time.perf_counter() - t0) / 100
    k2_ms = eng.time_distance_kernel(10)
    peak = eng.probe_sad_peak(4096)
    B, A, N, P = len(poses), 10, eng.n_views, 80
    ops = 6.0 * B * A * N * P
    print(json.dumps({"workload": "C2 world with chemistry (3 hues), chem_weight %.2f" % cw, "distance_kernel": eng.distance_kernel,
                      "us_per_step": dt * 1e6, "comparisons_per_sec": B * A * N / dt, "k2_us": k2_ms * 1e3,
                      "k2_int_TOPs_at_6_ops_per_pixel": ops / (k2_ms * 1e-3) / 1e12,
                      "k2_frac_of_vabsdiff4_peak": ops / (k2_ms * 1e-3) / (2.0 * peak),
                      "int_alu_peak_TOPs": 2.0 * peak / 1e12}))


if __name__ == "__main__":
    main()
