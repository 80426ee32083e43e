"""Step-batch time of the BASELINE.json configs that are not the bench headline (one GPU):
C1 (one agent, reference defaults: the engine, AND the drop-in NavBySceneFamiliarity class
stepped the way scripts/run_experiment.py does, next to the compiled reference on the host),
C3 (64x64 sensor, 360 headings, 8192 views; 1 and 64 agents), C4 (10^6 views; 1, 64 and 1024
agents) and the device-side build of a 10^6-view library from a 10^6-point path (N3).
One JSON line per case.  (evidence, not a bench arm)"""
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


PEAKS = {}


def report(name, eng, poses, A, P, steps, k2_peak, fake=False):
    per_step, k2_s = timed(eng, poses, steps, fake)
    B, N = len(poses), eng.n_views
    kern = eng.distance_kernel
    rec = {"workload": name, "agents": B, "headings": A, "views": N, "sensor_pixels": P,
           "us_per_step": per_step * 1e6, "comparisons_per_sec": B * A * N / per_step,
           "agent_steps_per_sec": B / per_step, "distance_kernel": kern, "k2_us": k2_s * 1e6,
           "k2_library_gbs": N * P / k2_s / 1e9}
    if kern == "k2_tc":
        if "mma" not in PEAKS:
            PEAKS["mma"] = eng.probe_mma_peak(4096)
        ops = 2.0 * B * A * N * eng.tc_planes * P
        rec["k2_tensor_TOPs"] = ops / k2_s / 1e12
        rec["k2_frac_of_int8_mma_probe"] = ops / k2_s / PEAKS["mma"]
        rec["k2_frac_of_2x_measured_bf16"] = ops / k2_s / 1e12 / (2.0 * MEASURED.get("bf16_tflops", 1590.0))
    else:
        rec["k2_frac_of_vabsdiff4_peak"] = 2.0 * B * A * N * P / k2_s / (2.0 * k2_peak)
        rec["k2_frac_of_measured_hbm"] = N * P / k2_s / 1e9 / MEASURED.get("hbm_gbs", 6650.0)
    print(json.dumps(rec), flush=True)


try:
    MEASURED = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    MEASURED = {}


def c1_dropin(L, kw, tpath):
    """configs[0] through the real API: nsf.step_forward() for a whole trajectory, the product
    class (device run-ahead + replay) and the compiled reference (host), same start pose."""
    import warnings
    import navsim
    from oracle import ref_loader
    pose = synthetic.start_pose(tpath, (0.05, 3.0), 80)
    frames = synthetic.default_frames(tpath, kw["step_size"])

    def drive(mod, nsf):
        nsf.train_from_path(tpath)
        nsf.position = (pose[0], pose[1])
        nsf.angle = pose[2]
        done, status = 0, 0
        t0 = time.perf_counter()
        try:
            for _ in range(frames):
                nsf.step_forward()
                done += 1
        except mod.StopNavigationException as e:
            status = e.get_code()
        return done, status, time.perf_counter() - t0, nsf.position

    drive(navsim, navsim.NavBySceneFamiliarity(L, familiarity_model=navsim.sads_familiarity(0.0), **kw))   # warm-up
    d1, s1, t1, p1 = drive(navsim, navsim.NavBySceneFamiliarity(L, familiarity_model=navsim.sads_familiarity(0.0), **kw))
    rec = {"workload": "C1 through the drop-in class: NavBySceneFamiliarity.step_forward() x %d frames "
                       "(run-ahead of 64 steps on the device, replayed per call)" % frames,
           "frames_completed": d1, "stop_status": s1, "us_per_step_forward": t1 / max(d1, 1) * 1e6,
           "agent_steps_per_sec": d1 / t1}
    ref = ref_loader.load_reference()
    if ref is not None:
        warnings.filterwarnings("ignore")
        d2, s2, t2, p2 = drive(ref, ref.NavBySceneFamiliarity(L, familiarity_model=ref.util.sads_familiarity(0.0), **kw))
        rec.update({"reference_us_per_step_forward": t2 / max(d2, 1) * 1e6, "reference_frames_completed": d2,
                    "same_trajectory": bool(d1 == d2 and s1 == s2 and tuple(p1) == tuple(p2)),
                    "speedup_vs_compiled_reference_one_core": (t2 / max(d2, 1)) / (t1 / max(d1, 1))})
    print(json.dumps(rec), flush=True)


def library_build_at_scale(L, kw):
    """N3: train_from_path on a 10^6-point path, on the device (sampler + planar layout +
    thermometer planes for the tensor-core kernel)."""
    tpath = synthetic.training_path_for(L.shape, 0.01414, 10, 0.0)[:1000000]
    eng = NavEngine(L, **kw)
    t0 = time.perf_counter()
    rc, bad = eng.train_from_path(tpath)
    eng.sync()
    t_build = time.perf_counter() - t0
    assert rc == 0, (rc, bad)
    poses = synthetic.start_pose_grid(tpath[::700], 80, n_lat=8, n_deg=8)
    eng.set_agents(poses, 4)
    t0 = time.perf_counter()
    eng.step(1)                      # first step: encodes the library planes for k2_tc
    eng.sync()
    t_first = time.perf_counter() - t0
    print(json.dumps({"workload": "N3: library build from a %d-point training path on the device" % len(tpath),
                      "views": len(tpath), "seconds_train_from_path": t_build, "views_per_sec": len(tpath) / t_build,
                      "seconds_first_step_incl_plane_encoding": t_first, "distance_kernel": eng.distance_kernel,
                      "library_bytes_u8_planes": int(3 * len(tpath) * 80),
                      "library_bytes_int8_thermometer_planes": int(len(tpath) * 384)}), flush=True)
    eng.close()


def main():
    which = sys.argv[1:] or ["c1", "c3", "c4", "n3"]
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
        c1_dropin(L, kw, tpath)
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
        for n in (8, 32):
            report("C4: 10^6 views, %d agents (one GPU holds the whole library)" % (n * n), eng,
                   synthetic.start_pose_grid(tpath, 80, n_lat=n, n_deg=n), 10, 80, 10, peak)
        eng.close()
    if "n3" in which:
        kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=10.0, n_test_angles=10,
                  n_sensor_levels=5, max_distance_to_training_path=450)
        library_build_at_scale(L, kw)


if __name__ == "__main__":
    main()
