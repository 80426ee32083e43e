"""Step-batch time of the BASELINE.json configs that are not the bench headline (one GPU):
C1 (one agent, reference defaults), C3 (64x64 sensor, 360 headings, 8192 views; 1 and 64
agents) and C4 (10^6 views; 1 agent).  One JSON line per case.  (evidence, not a bench arm)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200")):
    sys.path.insert(0, p)

import numpy as np

from navsim import NavEngine, synthetic


def timed(eng, poses, steps, fake=False):
    eng.set_agents(poses, steps * 4)
    eng.step(6, fake=fake)
    eng.sync()
    eng.rewind()
    eng.sync()
    t0 = time.perf_counter()
    eng.step(steps, fake=fake)
    eng.sync()
    dt = time.perf_counter() - t0
    k2_ms = eng.time_distance_kernel(10)
    return dt / steps, k2_ms * 1e-3


def report(name, eng, poses, A, P, steps, k2_peak, fake=False):
    per_step, k2_s = timed(eng, poses, steps, fake)
    B, N = len(poses), eng.n_views
    ops = 2.0 * B * A * N * P
    print(json.dumps({"workload": name, "agents": B, "headings": A, "views": N, "sensor_pixels": P,
                      "us_per_step": per_step * 1e6, "comparisons_per_sec": B * A * N / per_step,
                      "agent_steps_per_sec": B / per_step, "k2_us": k2_s * 1e6,
                      "k2_frac_of_vabsdiff4_peak": ops / k2_s / (2.0 * k2_peak),
                      "k2_library_gbs": N * P / k2_s / 1e9}), flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c3", "c4"]
    L = synthetic.make_landscape(3001, 2000, sigma=6.0)
    if "c1" in which:
        kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=10.0, n_test_angles=10,
                  n_sensor_levels=5, max_distance_to_training_path=450)
        tpath = synthetic.training_path_for(L.shape, 10.0, 10, 0.0)
        eng = NavEngine(L, **kw)
        assert eng.train_from_path(tpath) == (0, -1)
        peak = eng.probe_sad_peak(4096)
        report("C1: reference defaults, one agent", eng, np.array([synthetic.start_pose(tpath, (0.05, 3.0), 80)]),
               10, 80, 100, peak)
        eng.close()
    if "c3" in which:
        kw = dict(sensor_dimensions=(64, 64), sensor_pixel_dimensions=(1, 1), step_size=10.0, n_test_angles=360,
                  n_sensor_levels=5, saccade_degrees=180., max_distance_to_training_path=450)
        base = synthetic.training_path_for(L.shape, 10.0, 360, 0.0)
        # 8192 views: the genuine path resampled to that many points
        s = np.linspace(0, len(base) - 1, 8192)
        tpath = np.stack([np.interp(s, np.arange(len(base)), base[:, 0]), np.interp(s, np.arange(len(base)), base[:, 1])], 1)
        eng = NavEngine(L, **kw)
        t0 = time.perf_counter()
        assert eng.train_from_path(tpath) == (0, -1)
        build_s = time.perf_counter() - t0
        peak = eng.probe_sad_peak(4096)
        for B in (1, 64):
            poses = np.array([synthetic.start_pose(base, (0.02 * (b % 8) - 0.07, 2.0 * (b // 8) - 7.0), 64) for b in range(B)])
            report("C3: 64x64 sensor, 360 headings, %d agent(s); library build %.3f s" % (B, build_s), eng, poses,
                   360, 4096, 20, peak)
        eng.close()
    if "c4" in which:
        kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=10.0, n_test_angles=10,
                  n_sensor_levels=5, max_distance_to_training_path=450)
        tpath = synthetic.training_path_for(L.shape, 10.0, 10, 0.0)
        eng = NavEngine(L, **kw)
        assert eng.train_from_path(tpath) == (0, -1)
        genuine = eng.familiar_scenes
        n_total = 1000000
        levels = np.array([0, 63, 127, 191, 255], np.uint8)
        scenes = np.zeros((n_total, 2, 40, 3), np.uint8)
        scenes[..., 2] = levels[np.random.default_rng(3700).integers(0, 5, (n_total, 2, 40))]
        scenes[:len(genuine)] = genuine
        path = np.vstack([tpath, np.repeat(tpath[-1:], n_total - len(tpath), axis=0)])
        eng.set_library(scenes, path)
        peak = eng.probe_sad_peak(4096)
        report("C4: 10^6 views, one agent (one GPU holds the whole library)", eng,
               np.array([synthetic.start_pose(tpath, (0.05, 3.0), 80)]), 10, 80, 50, peak)
        eng.close()


if __name__ == "__main__":
    main()
