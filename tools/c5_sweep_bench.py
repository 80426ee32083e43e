"""BASELINE.json configs[4] (C5): a parameter sweep of 10^4 independent experiments
(8 landscapes 1000^2 x 5 path curves x 5 sensors x 5 step sizes x 10 start
offsets) through the batched driver navsim/experiments.py, on 1..8 GPUs.

  python tools/c5_sweep_bench.py [--gpus N] [--workers W] [--trials-per-axis ...] [--check K] [--cpu K]

Trials that share a world (everything but the start offset) are the agents of one engine
batch; one engine per host thread is reused for all its worlds; worlds are split over the
GPUs with no collective.  --check K compares K trials (spread over the grid) with the oracle
run one trial at a time, formatted result record for record; --cpu K times K oracle trials
on all host cores (the reference's deployment: one process per core over independent trials,
scripts/run_experiment.py:327) for the experiments/s baseline.  One JSON line on stdout.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

SENSORS = [[20, 4, 2, 2], [40, 2, 2, 4], [8, 2, 2, 2], [16, 16, 1, 1], [40, 4, 2, 2]]
STEPS = [2.0, 4.0, 6.0, 8.0, 10.0]
CURVES = [0.0, 0.25, 0.5, 0.75, 1.0]
OFFSETS = [[la, de] for la, de in zip(np.linspace(-0.4, 0.4, 10).round(3).tolist(), np.linspace(-20, 20, 10).round(2).tolist())]
LEVELS = [3, 5, 8]


def grid(n_land=8, side=1000):
    return {"landscape_class": ["synthetic"], "landscape_name": ["land%d" % i for i in range(n_land)],
            "training_path_curve": CURVES, "sensor_dimensions": SENSORS, "step_size": STEPS, "start_offset": OFFSETS,
            "n_sensor_levels": [5], "n_test_angles": [10]}


def landscapes(n_land=8, side=1000):
    from navsim import synthetic
    return {"land%d" % i: synthetic.make_landscape(5000 + i, side, sigma=6.0) for i in range(n_land)}


def _rank(argv):
    rank, world, workers, n_land, side = argv
    import torch
    from navsim import experiments as X
    torch.cuda.set_device(rank)
    lands = landscapes(n_land, side)
    variables, trials = X.expand_trials(grid(n_land, side))
    worlds = X.split_worlds(X.group_worlds(trials), world, rank)
    # warm-up: one small world (library, graph capture, lazy module load)
    X.run_trials(trials, lands, device=rank, workers=1, worlds=worlds[:1])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = X.run_trials(trials, lands, device=rank, workers=workers, worlds=worlds)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mine = sorted(i for _, ix in worlds for i in ix)
    return rank, dt, {i: res[i] for i in mine}


def _oracle_trial(argv):
    tr, side = argv
    from navsim import experiments as X, synthetic
    from navsim.engine import n_captures, percent_recapitulated_forgiving
    from oracle import oracle as O
    land = synthetic.make_landscape(5000 + int(tr["landscape_name"][4:]), side, sigma=6.0)
    sd = tr["sensor_dimensions"]
    w = O.World(land, sd[0:2], tr["step_size"], n_test_angles=tr["n_test_angles"], sensor_pixel_dimensions=sd[2:4],
                max_distance_to_training_path=450, n_sensor_levels=tr["n_sensor_levels"])
    tpath = synthetic.training_path_for(land.shape, tr["step_size"], tr["n_test_angles"], tr["training_path_curve"])
    t0 = time.perf_counter()
    assert w.train_from_path(tpath) == (0, -1)
    pose = synthetic.start_pose(tpath, tr["start_offset"], sd[0] * sd[2])
    frames = synthetic.default_frames(tpath, tr["step_size"])
    ag = w.new_agent(*pose)
    r = w.run(ag, frames)
    dt = time.perf_counter() - t0
    cov = ag._cov.astype(bool)
    with np.errstate(invalid="ignore", divide="ignore"):
        rmsd = float(np.sqrt(ag.nav_err / ag.n_nav_err)) if ag.n_nav_err else float("nan")
    return dict(path_coverage=cov.sum() / len(cov), rmsd_error=rmsd, completed_frames=int(r["completed"]),
                stop_status=int(r["status"]), n_captures=n_captures(cov, 0.05),
                percent_forgiving=percent_recapitulated_forgiving(cov, 0.05)), dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workers", type=int, default=6)
    ap.add_argument("--landscapes", type=int, default=8)
    ap.add_argument("--side", type=int, default=1000)
    ap.add_argument("--check", type=int, default=200, help="trials compared with the oracle")
    ap.add_argument("--cpu", type=int, default=64, help="oracle trials timed on all host cores")
    args = ap.parse_args()
    from navsim import experiments as X
    variables, trials = X.expand_trials(grid(args.landscapes, args.side))
    jobs = [(r, args.gpus, args.workers, args.landscapes, args.side) for r in range(args.gpus)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    if args.gpus == 1:
        done = [_rank(jobs[0])]
    else:
        with ctx.Pool(args.gpus) as pool:
            done = pool.map(_rank, jobs)
    wall = time.perf_counter() - t0
    results = {}
    for _, _, res in done:
        results.update(res)
    assert len(results) == len(trials)
    t_gpu = max(d[1] for d in done)
    out = {"workload": "C5: %d experiments = %d landscapes %d^2 x %d curves x %d sensors x %d step sizes x %d start offsets"
                       % (len(trials), args.landscapes, args.side, len(CURVES), len(SENSORS), len(STEPS), len(OFFSETS)),
           "n_gpus": args.gpus, "host_threads_per_gpu": args.workers, "worlds": len(X.group_worlds(trials)),
           "seconds": t_gpu, "experiments_per_sec": len(trials) / t_gpu, "seconds_per_rank": [round(d[1], 3) for d in done],
           "wall_seconds_incl_process_start_and_landscape_generation": wall,
           "agent_steps": int(sum(r["completed_frames"] for r in results.values())),
           "stop_status_histogram": {str(k): int(v) for k, v in zip(*np.unique([r["stop_status"] for r in results.values()], return_counts=True))}}
    fmt = X.RESULT_FORMATS
    ncores = len(os.sched_getaffinity(0))
    pick = np.unique(np.linspace(0, len(trials) - 1, max(args.check, args.cpu)).astype(int))
    if len(pick):
        t1 = time.perf_counter()
        with ctx.Pool(ncores) as pool:
            ref = pool.map(_oracle_trial, [(trials[i], args.side) for i in pick])
        t_cpu = time.perf_counter() - t1
        bad = []
        for i, (want, _) in zip(pick[:args.check] if args.check < len(pick) else pick, ref):
            got = results[int(i)]
            for k in fmt:
                if fmt[k].format(got[k]) != fmt[k].format(want[k]):
                    bad.append((int(i), k, got[k], want[k]))
        out["oracle_check"] = {"trials": int(min(args.check, len(pick))), "mismatching_fields": len(bad), "first": bad[:3]}
        out["cpu_baseline"] = {"kind": "port", "cores": ncores, "trials": len(pick), "seconds": t_cpu,
                               "experiments_per_sec": len(pick) / t_cpu,
                               "sum_of_trial_seconds": float(sum(d for _, d in ref)),
                               "sample": "%d trials spread evenly over the grid, oracle (C restatement) one trial per process "
                                         "slot on all cores" % len(pick)}
    print(json.dumps(out))
    if out.get("oracle_check", {}).get("mismatching_fields"):
        sys.exit(1)


if __name__ == "__main__":
    main()
