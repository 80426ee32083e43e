"""Batched experiment driver (navsim/experiments.py) against the reference driver's
conventions (scripts/run_experiment.py) and, on the GPU, against the oracle run
trial by trial."""
import io
import os

import numpy as np
import pytest

GRID = {
    "_comment": "small C5-style sweep",
    "landscape_class": ["synthetic"],
    "landscape_name": ["a", "b"],
    "training_path_curve": [0.0, 0.5],
    "sensor_dimensions": [[40, 2, 2, 4], [8, 2, 2, 2]],
    "n_sensor_levels": [5],
    "step_size": [5.0],
    "n_test_angles": [10],
    "start_offset": [[0.0, 0.0], [0.1, 5.0], [-0.3, -10.0]],
    "landscape_flip_vertical": [False, True],
}


def test_trial_grid_and_csv_format():
    from navsim import experiments as X
    variables, trials = X.expand_trials(GRID)
    assert variables == sorted(k for k in GRID if not k.startswith("_comment"))
    assert len(trials) == 2 * 2 * 2 * 3 * 2
    assert trials[0]["landscape_name"] == "a" and trials[-1]["landscape_name"] == "b"
    with pytest.raises(ValueError):
        X.expand_trials({"step_size": 5.0})
    parts = [X.split_trials(trials, 5, r) for r in range(5)]
    assert [len(p) for p in parts] == [len(x) for x in np.array_split(np.arange(len(trials)), 5)]
    assert sum(parts, []) == trials
    res = dict(path_coverage=0.5, rmsd_error=1.25, completed_frames=17, stop_status=-2, n_captures=3,
               percent_forgiving=0.75)
    row = X.format_row(variables, trials[1], res)
    # reference formats: "{:4f}", "{0[0]:d};{0[1]:d};...", bools as ints, results sorted by name
    assert row.startswith("synthetic, 0, a, 5, 10, 40;2;2;4, 0.000000;0.000000, 5.000000, 0.500000, ")
    assert row.endswith("17, 3, 0.500000, 0.750000, 1.250000, -2")


@pytest.mark.parametrize("world_size", [1, 2, 3, 4, 8, 16])
def test_worlds_split_over_ranks_is_a_partition(world_size):
    """Multi-GPU sweeps: every world -- the trials that become the agents of one
    engine batch -- goes to exactly one rank, in order, no collective; shares are balanced by trial
    count to within one world; every trial of the grid is run by exactly one rank."""
    from navsim import experiments as X
    variables, trials = X.expand_trials(GRID)
    worlds = X.group_worlds(trials)
    assert sorted(i for _, ix in worlds for i in ix) == list(range(len(trials)))
    for _, ix in worlds:      # a world = same everything but the start offset
        assert len({tuple(sorted((k, str(v)) for k, v in trials[i].items() if k != "start_offset")) for i in ix}) == 1
    shares = [X.split_worlds(worlds, world_size, r) for r in range(world_size)]
    assert sum(shares, []) == worlds                               # contiguous, ordered, complete, disjoint
    counts = [sum(len(ix) for _, ix in sh) for sh in shares]
    biggest = max(len(ix) for _, ix in worlds)
    assert max(counts) - min(counts) <= 2 * biggest or world_size > len(worlds)
    assert sum(counts) == len(trials)


@pytest.mark.gpu
def test_sweep_matches_oracle_trial_by_trial(gpu, tmp_path):
    from navsim import experiments as X, synthetic
    from oracle import oracle as O
    lands = {"a": synthetic.make_landscape(6001, 420, sigma=6.0), "b": synthetic.make_landscape(6002, 400, sigma=5.0)}
    variables, trials = X.expand_trials(GRID)
    results = X.run_trials(trials, lands)
    fmt = X.RESULT_FORMATS
    n_end = 0
    for tr, res in zip(trials, results):
        land = lands[tr["landscape_name"]][::(-1 if tr["landscape_flip_vertical"] else 1)]
        sd = tr["sensor_dimensions"]
        w = O.World(land, sd[0:2], tr["step_size"], n_test_angles=tr["n_test_angles"], sensor_pixel_dimensions=sd[2:4],
                    max_distance_to_training_path=450, n_sensor_levels=tr["n_sensor_levels"])
        tpath = synthetic.training_path_for(land.shape, tr["step_size"], tr["n_test_angles"], tr["training_path_curve"])
        assert w.train_from_path(tpath) == (0, -1)
        pose = synthetic.start_pose(tpath, tr["start_offset"], sd[0] * sd[2])
        frames = synthetic.default_frames(tpath, tr["step_size"])
        ag = w.new_agent(*pose)
        r = w.run(ag, frames)
        cov = ag._cov.astype(bool)
        from navsim.engine import n_captures, percent_recapitulated_forgiving
        want = dict(path_coverage=cov.sum() / len(cov), rmsd_error=np.sqrt(ag.nav_err / ag.n_nav_err),
                    completed_frames=r["completed"], stop_status=r["status"],
                    n_captures=n_captures(cov, 0.05), percent_forgiving=percent_recapitulated_forgiving(cov, 0.05))
        for k in fmt:
            assert fmt[k].format(res[k]) == fmt[k].format(want[k]), (tr, k)
        n_end += int(r["status"] == 1)
    assert n_end > 0                                   # some runs do reach the end of the path
    out = tmp_path / "task-0.csv"
    X.write_csv(str(out), variables, trials, results)
    lines = out.read_text().splitlines()
    assert len(lines) == len(trials) + 1 and lines[0].endswith("n_captures, path_coverage, percent_forgiving, rmsd_error, stop_status")
    import pandas as pd                                 # the reference's loader reads it like this
    df = pd.read_csv(str(out), sep=",", header=0, skipinitialspace=True,
                     converters={"sensor_dimensions": lambda s: [int(e) for e in s.split(";")],
                                 "start_offset": lambda s: [float(e) for e in s.split(";")]})
    assert list(df["completed_frames"]) == [r["completed_frames"] for r in results]
