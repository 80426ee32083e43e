"""Batched experiment driver: the trial grid, result record and
CSV schema of scripts/run_experiment.py, fed to the batched engine.

The reference runs one trial at a time (make_nsf + run_experiment,
scripts/run_experiment.py:146-258) and shards trials over MPI ranks
(:327-328).  Here trials that share a world (landscape, sensor, heading sweep,
path, metric) and differ only in `start_offset` become the agents of ONE
NavEngine batch; worlds are processed one after another; ranks take a
contiguous slice of the worlds' trials exactly like np.array_split.

Landscapes come either as HSV uint8 arrays (name -> array; used as they are) or from
<landscape_dir>/<class>/<name> image files, prepared on the host the way make_nsf does
(run_experiment.py:160-199: PIL -> HSV, V >= 200 threshold, modal filter, grain labelling,
chemistry painted with set_HS_where_equal, flips).

    python -m navsim.experiments trials.json landscape_dir/ [--gpus N] [--workers W] [--seed S]

keeps the reference driver's command line, output directory naming and CSV schema
(loadable by scripts/load_experiments.load_runs).  One engine per host thread is reused for
every world it runs (set_landscape / set_world / train_from_path: no engine, stream or graph
re-creation beyond what the new shapes need); several threads per GPU keep small worlds from
leaving the device idle; worlds are split over GPUs without any collective.
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


LANDSCAPE_THRESHOLD = 200     # run_experiment.py:30


def _world_key(trial, index=None):
    """Trials with the same key are agents of one world (same landscape pixels, sensor,
    heading sweep, training path and metric); they differ in start_offset only.  With two or
    more chemicals the reference paints fresh random hues for EVERY trial
    (run_experiment.py:126-142,186-193), so such a trial is a world of its own."""
    t = dict(DEFAULTS)
    t.update(trial)
    sd = tuple(int(v) for v in t['sensor_dimensions'])
    n_chem = int(t.get('n_chemicals', 1))
    return (t.get('landscape_class', ''), t['landscape_name'], bool(t.get('landscape_flip_vertical', False)),
            bool(t.get('landscape_flip_horizontal', False)), sd, int(t.get('n_sensor_levels', 5)),
            int(t.get('mask_middle_n', 0)), float(t['step_size']), int(t['n_test_angles']),
            float(t.get('saccade_degrees', 180.)), float(t['training_path_curve']),
            float(t.get('chem_weight', 0.)), float(t['max_distance_to_training_path']),
            n_chem, float(t.get('min_chem_grain_diameter', 2)), index if n_chem >= 2 else None)


# ---- landscape preparation (host; scripts/run_experiment.py:160-199) ---------------------
class LandscapeStore(object):
    """name -> HSV uint8 array, either handed in as arrays or loaded from
    <landscape_dir>/<class>/<name> the way make_nsf does: PIL -> HSV, grains = 8-connected
    components of (V >= 200) after a modal filter, memoised per (class, name)."""

    def __init__(self, landscapes=None, landscape_dir=None, seed=None, trials=None):
        self.arrays = dict(landscapes or {})
        self.dir = landscape_dir
        self.rng = np.random.default_rng(seed)     # the reference's RNG is unseeded (run_experiment.py:2)
        self._grains = {}
        # make_nsf memoises (landscape, grain labels, grain properties) per (class, name) ONLY
        # (run_experiment.py:163-182): the grains of a landscape are labelled with the
        # min_chem_grain_diameter of the FIRST trial that uses it and reused by every later
        # trial, whatever its own value.  Kept: the drop-in must give the reference's rows.
        self.first_min_d = {}
        for tr in (trials or []):
            k = (tr.get('landscape_class', ''), tr['landscape_name'])
            self.first_min_d.setdefault(k, tr.get('min_chem_grain_diameter', 2))

    def base(self, cls, name):
        if name in self.arrays:
            return self.arrays[name]
        key = (cls, name)
        if key not in self.arrays:
            from PIL import Image
            import os
            self.arrays[key] = np.asarray(Image.open(os.path.join(self.dir, str(cls), str(name))).convert('HSV'))
        return self.arrays[key]

    def grains(self, cls, name, min_diameter):
        min_diameter = self.first_min_d.get((cls, name), min_diameter)
        key = (cls, name)
        if key not in self._grains:
            from . import compat
            land = self.base(cls, name)
            for_labeling = (land[:, :, 2] >= LANDSCAPE_THRESHOLD).astype(np.uint8)
            w = min_diameter // 2
            if not w == 0:
                if w % 2 == 0:
                    w -= 1
                w = int(w)
                for_labeling = compat.modal_filter(for_labeling, np.ones((w, w), np.uint8))
            labels = compat.label_image(for_labeling)
            self._grains[key] = (labels, compat.region_props(labels))
        return self._grains[key]

    def prepared(self, key):
        """The landscape a world steps on: chemistry painted (n_chemicals >= 1), then flipped."""
        from .util import set_HS_where_equal
        cls, name, flip_v, flip_h = key[0], key[1], key[2], key[3]
        n_chem, min_d = key[13], key[14]
        land = self.base(cls, name)
        if n_chem >= 1 and self.dir is not None:
            labels, props = self.grains(cls, name, min_d)
            land = land.copy()
            chems = self.rng.integers(n_chem, size=len(props), dtype=np.uint8) * (255 // n_chem)
            sats = self.rng.integers(127, 128, size=len(props), dtype=np.uint8)
            for g, pr in enumerate(props):
                if pr.equivalent_diameter < min_d:
                    sats[g] = 0
            set_HS_where_equal(labels, land, chems, sats)
        return land[::(-1 if flip_v else 1), ::(-1 if flip_h else 1)]            # run_experiment.py:196-199


# ---- running worlds ------------------------------------------------------------------------
def _run_world(eng_cache, store, key, idxs, trials, device):
    (_cls, name, flip_v, flip_h, sd, levels, mask, step, A, saccade, curve, cw, max_dist, n_chem, _md, _uniq) = key
    land_key = key[:4] + (n_chem, _md, _uniq)
    world = dict(sensor_dimensions=sd[0:2], step_size=step, n_test_angles=A, sensor_pixel_dimensions=sd[2:4],
                 max_distance_to_training_path=max_dist, n_sensor_levels=levels, mask_middle_n=mask,
                 saccade_degrees=saccade, chem_weight=cw)
    eng = eng_cache.get("engine")
    if eng is None:
        land = store.prepared(key)
        eng = NavEngine(land, device=device, **world)
        eng_cache["engine"], eng_cache["land_key"] = eng, land_key
    else:
        if eng_cache["land_key"] != land_key:          # another landscape: one upload, same engine
            eng.set_landscape(store.prepared(key))
            eng_cache["land_key"] = land_key
        eng.set_world(**world)
    land = eng.landscape
    tpath = synthetic.training_path_for(land.shape, step, A, curve)          # :203-213
    rc, bad = eng.train_from_path(tpath)
    if rc != 0:
        raise RuntimeError("training path leaves the landscape at point %d (status %d)" % (bad, rc))
    spw = sd[0] * sd[2]                                                      # sensor_pixel_width, :155
    poses = [synthetic.start_pose(tpath, trials[i]['start_offset'], spw) for i in idxs]   # :223-229
    frames = int(FRAME_FACTOR * eng.training_path_length / step)             # :238
    eng.run(np.asarray(poses), frames)
    out = eng.results(N_CONSECUTIVE_SCENES)
    res = {}
    for j, i in enumerate(idxs):
        res[i] = {
            'path_coverage': float(out['path_coverage'][j]), 'rmsd_error': float(out['rmsd_error'][j]),
            'completed_frames': int(out['completed_frames'][j]), 'stop_status': int(out['stop_status'][j]),
            'percent_forgiving': float(out['percent_forgiving'][j]), 'n_captures': int(out['n_captures'][j]),
        }
    return res


def group_worlds(trials):
    """[(world key, [trial indices])], worlds in order of first appearance, worlds on the same
    landscape next to each other (one upload serves them all)."""
    groups = {}
    for i, tr in enumerate(trials):
        groups.setdefault(_world_key(tr, i), []).append(i)
    order = sorted(groups.items(), key=lambda kv: (str(kv[0][:4]), kv[1][0]))
    return order


def run_trials(trials, landscapes=None, device=None, landscape_dir=None, workers=1, seed=None, worlds=None):
    """Runs every trial; returns one result dict per trial, in order, with the keys of
    run_experiment()'s record (run_experiment.py:251-258).  `worlds`: only these
    (key, indices) groups (a rank's share); results of the others stay None.
    workers > 1: that many host threads, each with its own engine and stream on `device`,
    take worlds from a shared queue (small worlds leave most of a B200 idle)."""
    store = LandscapeStore(landscapes, landscape_dir, seed, trials)
    worlds = group_worlds(trials) if worlds is None else worlds
    results = [None] * len(trials)
    if workers <= 1:
        cache = {}
        for key, idxs in worlds:
            for i, r in _run_world(cache, store, key, idxs, trials, device).items():
                results[i] = r
        if cache.get("engine") is not None:
            cache["engine"].close()
        return results
    import queue
    import threading
    q = queue.Queue()
    for w in worlds:
        q.put(w)
    errors = []
    lock = threading.Lock()

    def work():
        cache = {}
        try:
            while True:
                try:
                    key, idxs = q.get_nowait()
                except queue.Empty:
                    break
                with lock:                      # landscape preparation touches shared caches and the RNG
                    store.prepared(key) if False else None
                res = _run_world(cache, _LockedStore(store, lock), key, idxs, trials, device)
                for i, r in res.items():
                    results[i] = r
        except Exception as e:   # noqa: BLE001 -- reported to the caller below
            errors.append(e)
        finally:
            if cache.get("engine") is not None:
                cache["engine"].close()

    threads = [threading.Thread(target=work) for _ in range(int(workers))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


class _LockedStore(object):
    def __init__(self, store, lock):
        self._store, self._lock = store, lock

    def prepared(self, key):
        with self._lock:
            return self._store.prepared(key)


def split_worlds(worlds, world_size, rank):
    """Contiguous share of the world list for one rank, balanced by trial count (the reference
    splits the raw trial list, run_experiment.py:327-328; keeping a world on one rank builds
    its library once)."""
    total = sum(len(ix) for _, ix in worlds)
    bounds = [total * r / float(world_size) for r in range(world_size + 1)]
    out, acc = [], 0
    for w in worlds:
        mid = acc + 0.5 * len(w[1])
        if bounds[rank] <= mid < bounds[rank + 1]:
            out.append(w)
        acc += len(w[1])
    return out


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


# ---- command line: the reference driver's (scripts/run_experiment.py:262-352) --------------
def _make_outdir(mode):
    import os
    import time
    run_number = 0
    datestring = time.strftime("%Y-%m-%d")
    while True:                                                   # :303-311
        outdir = "output-%s-%s-run%i" % (mode, datestring, run_number)
        if not os.path.exists(outdir):
            os.makedirs(outdir)
            return outdir
        run_number += 1


def run_rank(trial_file, landscape_dir, outdir, rank, world_size, device, workers=4, seed=None, landscapes=None):
    """One rank's share: writes <outdir>/task-<rank>.csv; returns (trials run, seconds)."""
    import json
    import os
    import time
    with open(trial_file) as f:
        variables, trials = expand_trials(json.load(f))
    for tr in trials:
        for k in ('sensor_dimensions', 'start_offset'):
            if k in tr:
                tr[k] = np.asarray(tr[k])
    worlds = split_worlds(group_worlds(trials), world_size, rank)
    t0 = time.time()
    results = run_trials(trials, landscapes=landscapes, device=device, landscape_dir=landscape_dir,
                         workers=workers, seed=seed, worlds=worlds)
    mine = [i for _, ix in worlds for i in ix]
    mine.sort()
    write_csv(os.path.join(outdir, "task-%i.csv" % rank), variables, [trials[i] for i in mine],
              [results[i] for i in mine])
    return len(mine), time.time() - t0


def _rank_entry(argv):
    trial_file, landscape_dir, outdir, rank, world_size, device, workers, seed = argv
    return run_rank(trial_file, landscape_dir, outdir, rank, world_size, device, workers, seed)


def main(argv=None):
    """python -m navsim.experiments trials.json landscape_dir/ [--gpus N] [--workers W] [--seed S]
    Same positional arguments, output directory naming, CSV schema and formats as
    scripts/run_experiment.py; trials that share a world are stepped as one agent batch, worlds
    are spread over --gpus devices (one process each) or, under torchrun / mpirun, over the ranks."""
    import argparse
    import logging
    import os
    import time
    ap = argparse.ArgumentParser(prog="python -m navsim.experiments")
    ap.add_argument("trial_file")
    ap.add_argument("landscape_dir")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workers", type=int, default=4, help="host threads (engines) per GPU")
    ap.add_argument("--seed", type=int, default=None, help="chemistry RNG seed (the reference's is unseeded)")
    ap.add_argument("--outdir", default=None)
    args = ap.parse_args(argv)
    logging.basicConfig(format='%(asctime)s %(levelname)-8s %(message)s', level=logging.INFO, datefmt='%Y-%m-%d %H:%M:%S')
    log = logging.getLogger('experiments')
    mode = os.path.basename(args.trial_file)
    if mode.endswith('.json'):
        mode = mode[:-5]
    landscape_dir = os.path.abspath(args.landscape_dir)
    from .compat import FileComm
    comm = FileComm()
    t0 = time.time()
    if comm.size > 1:                                  # launched by torchrun / mpirun: one rank per GPU
        outdir = comm.bcast(args.outdir or (_make_outdir(mode) if comm.rank == 0 else None), root=0)
        os.makedirs(outdir, exist_ok=True)
        device = int(os.environ.get("LOCAL_RANK", comm.rank))
        n, dt = run_rank(args.trial_file, landscape_dir, outdir, comm.rank, comm.size, device, args.workers, args.seed)
        comm.barrier()
        if comm.rank == 0:
            log.info("Done! rank 0 ran %i trials in %.1f s" % (n, dt))
        return outdir
    outdir = args.outdir or _make_outdir(mode)
    os.makedirs(outdir, exist_ok=True)
    jobs = [(args.trial_file, landscape_dir, outdir, r, args.gpus, r, args.workers, args.seed) for r in range(args.gpus)]
    if args.gpus <= 1:
        done = [_rank_entry(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(args.gpus) as pool:
            done = pool.map(_rank_entry, jobs)
    n = sum(d[0] for d in done)
    log.info("Done! Finished %i trials" % n)
    log.info("It took about %.1f s; each trial added about %.4f s of wall-clock time" % (time.time() - t0, (time.time() - t0) / max(n, 1)))
    return outdir


if __name__ == "__main__":
    main()
