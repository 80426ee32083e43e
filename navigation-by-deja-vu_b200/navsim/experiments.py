"""Batched experiment driver: the trial grid, result record and
CSV schema of scripts/run_experiment.py, fed to the batched engine.

The reference runs one trial at a time (make_nsf + run_experiment,
scripts/run_experiment.py:146-258) and shards trials over MPI ranks
(:327-328).  Here trials that share a world (landscape, sensor, heading sweep,
path, metric) and differ only in `start_offset` become the agents of ONE
NavEngine batch; worlds are processed one after another; ranks take a
contiguous slice of the worlds' trials exactly like np.array_split.

Landscapes are passed in as HSV uint8 arrays (name -> array).  Loading PNGs,
grain labelling and chemistry painting (run_experiment.py:160-199, skimage /
set_HS_where_equal) are landscape preparation and stay with the caller.
"""
import itertools

import numpy as np

from . import synthetic
from .engine import NavEngine

FRAME_FACTOR = 3.0            # run_experiment.py:21
N_CONSECUTIVE_SCENES = 0.05   # run_experiment.py:28

DEFAULTS = {'n_test_angles': 10, 'max_distance_to_training_path': 450}   # run_experiment.py:38-42

_float = "{:6f}"
RESULT_FORMATS = {            # run_experiment.py:44-52
    'path_coverage': _float, 'rmsd_error': _float, 'completed_frames': "{:d}", 'stop_status': "{:d}",
    'n_captures': "{:d}", 'percent_forgiving': _float,
}
VARIABLE_FORMATS = {          # run_experiment.py:54-71
    'landscape_class': "{}", 'landscape_name': "{}", 'training_path_curve': "{:4f}",
    'landscape_noise_factor': "{:4f}", 'n_chemicals': "{:d}", 'min_chem_grain_diameter': "{:4f}",
    'chem_weight': "{:4f}", 'sensor_dimensions': "{0[0]:d};{0[1]:d};{0[2]:d};{0[3]:d}",
    'mask_middle_n': "{:d}", 'n_sensor_levels': "{:d}", 'step_size': "{:4f}", 'saccade_degrees': "{:4f}",
    'n_test_angles': "{:d}", 'start_offset': "{0[0]:4f};{0[1]:4f}",
    'landscape_flip_vertical': '{:d}', 'landscape_flip_horizontal': '{:d}',
}


def expand_trials(variable_dict):
    """JSON grid -> (sorted variable names, list of trial dicts): every value is a list,
    `_comment*` keys are dropped, cartesian product over the sorted variables
    (run_experiment.py:269-290)."""
    vd = {}
    for k, v in variable_dict.items():
        if k.startswith("_comment"):
            continue
        if not isinstance(v, list):
            raise ValueError("Variable %s must have a list of values!" % k)
        vd[k] = v
    variables = sorted(vd)
    trials = [dict(zip(variables, vals)) for vals in itertools.product(*[vd[k] for k in variables])]
    return variables, trials


def split_trials(trials, world_size, rank):
    """Static contiguous split like np.array_split(trials, size)[rank] (run_experiment.py:327-328),
    without building a ragged array (which NumPy >= 1.24 refuses)."""
    from .sharded import shard_bounds
    off, cnt = shard_bounds(len(trials), world_size, rank)
    return trials[off:off + cnt]


def _world_key(trial):
    t = dict(DEFAULTS)
    t.update(trial)
    sd = tuple(int(v) for v in t['sensor_dimensions'])
    return (t.get('landscape_class', ''), t['landscape_name'], bool(t.get('landscape_flip_vertical', False)),
            bool(t.get('landscape_flip_horizontal', False)), sd, int(t.get('n_sensor_levels', 5)),
            int(t.get('mask_middle_n', 0)), float(t['step_size']), int(t['n_test_angles']),
            float(t.get('saccade_degrees', 180.)), float(t['training_path_curve']),
            float(t.get('chem_weight', 0.)), float(t['max_distance_to_training_path']))


def run_trials(trials, landscapes, device=None):
    """Runs every trial; returns one result dict per trial, in order, with the keys of
    run_experiment()'s record (run_experiment.py:251-258)."""
    groups = {}
    for i, tr in enumerate(trials):
        groups.setdefault(_world_key(tr), []).append(i)
    results = [None] * len(trials)
    for key, idxs in groups.items():
        (_cls, name, flip_v, flip_h, sd, levels, mask, step, A, saccade, curve, cw, max_dist) = key
        land = landscapes[name]
        land = land[::(-1 if flip_v else 1), ::(-1 if flip_h else 1)]            # run_experiment.py:196-199
        eng = NavEngine(land, sd[0:2], step, n_test_angles=A, sensor_pixel_dimensions=sd[2:4],
                        max_distance_to_training_path=max_dist, n_sensor_levels=levels,
                        mask_middle_n=mask, saccade_degrees=saccade, chem_weight=cw, device=device)
        tpath = synthetic.training_path_for(land.shape, step, A, curve)          # :203-213
        rc, bad = eng.train_from_path(tpath)
        if rc != 0:
            raise RuntimeError("training path leaves the landscape at point %d (status %d)" % (bad, rc))
        spw = sd[0] * sd[2]                                                      # sensor_pixel_width, :155
        poses = [synthetic.start_pose(tpath, trials[i]['start_offset'], spw) for i in idxs]   # :223-229
        frames = int(FRAME_FACTOR * eng.training_path_length / step)             # :238
        res = eng.run(np.asarray(poses), frames)
        out = eng.results(N_CONSECUTIVE_SCENES)
        for j, i in enumerate(idxs):
            results[i] = {
                'path_coverage': float(out['path_coverage'][j]), 'rmsd_error': float(out['rmsd_error'][j]),
                'completed_frames': int(out['completed_frames'][j]), 'stop_status': int(out['stop_status'][j]),
                'percent_forgiving': float(out['percent_forgiving'][j]), 'n_captures': int(out['n_captures'][j]),
            }
        eng.close()
    return results


def format_row(variables, trial, result):
    """One CSV line in the reference's format (run_experiment.py:341-344)."""
    left = ", ".join(VARIABLE_FORMATS[v].format(trial[v]) for v in variables)
    right = ", ".join(RESULT_FORMATS[v].format(result[v]) for v in sorted(RESULT_FORMATS))
    return left + ", " + right


def write_csv(path, variables, trials, results):
    """task-<rank>.csv as scripts/run_experiment.py writes it (:318-344), loadable by
    scripts/load_experiments.load_runs."""
    with open(path, "w") as f:
        print(", ".join(list(variables) + sorted(RESULT_FORMATS)), file=f)
        for tr, res in zip(trials, results):
            print(format_row(variables, tr, res), file=f)
