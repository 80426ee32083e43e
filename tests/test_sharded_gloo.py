"""View-sharded stepping over torch.distributed (gloo, world_size 2, CPU).

Exercises the host-side logic of navsim/sharded.py -- contiguous view split,
packed (difference, global view index) keys, the two int64 MIN all-reduces and
the identical decision on every rank -- with the oracle standing in for the CUDA
engine, and checks the result against the unsharded oracle trajectory."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

EXACT_NONE = np.array([np.inf]).view(np.int64)[0]


class OracleShardEngine(object):
    """Engine-like stand-in: phases 1-3 of the step on the CPU via the oracle."""

    def __init__(self, L, world_kw, tpath, scenes, rank, world, poses):
        import torch
        from navsim.sharded import shard_bounds
        from oracle import oracle as O
        self.O = O
        self.w = O.World(L, **world_kw)
        self.off, self.cnt = shard_bounds(len(scenes), world, rank)
        self.scenes = np.ascontiguousarray(scenes[self.off:self.off + self.cnt])
        self.path = tpath
        self.poses = np.array(poses, dtype=np.float64)
        self.A = self.w.A
        B = len(self.poses)
        self.keys = torch.zeros(B * self.A, dtype=torch.int64)
        self.exact = torch.zeros(B * self.A, dtype=torch.int64)
        self.best_log = []

    def keys_tensor(self):
        return self.keys

    def exact_tensor(self):
        return self.exact

    def phase(self, which, fake=False, log_afam=False):
        from navsim.sharded import pack_key, unpack_key
        O, w = self.O, self.w
        B, A = len(self.poses), self.A
        if which == 1:
            self.glimpses = []
            keys = np.empty(B * A, np.int64)
            for b, (x, y, ang) in enumerate(self.poses):
                for k in range(A):
                    a = (ang + w.offsets[k]) % (2 * np.pi)
                    rc, g = w.get_sensor_mat((x, y), a)
                    assert rc == 0
                    self.glimpses.append(g)
                    _, vt = O.sad_int(self.scenes, g)
                    j = int(np.argmin(vt))
                    keys[b * A + k] = pack_key(int(vt[j]), self.off + j)
            self.keys.copy_(__import__("torch").from_numpy(keys))
        elif which == 2:
            keys = self.keys.numpy()
            score, view = unpack_key(keys)
            ex = np.full(B * A, EXACT_NONE, np.int64)
            for b in range(B):
                sl = slice(b * A, (b + 1) * A)
                gmin = score[sl].min()
                tied = np.flatnonzero(score[sl] == gmin)
                for k in range(A):
                    g = b * A + k
                    cand = []
                    if len(tied) > 1 and k in tied:
                        _, vt = O.sad_int(self.scenes, self.glimpses[g])
                        cand = list(np.flatnonzero(vt == gmin))
                    elif self.off <= view[g] < self.off + self.cnt:
                        cand = [int(view[g]) - self.off]
                    if cand:
                        # the exact difference itself, as the kernels keep it
                        d = min(self._exact_diff(self.scenes[c], self.glimpses[g]) for c in cand)
                        ex[g] = np.array([d]).view(np.int64)[0]
            self.exact.copy_(__import__("torch").from_numpy(ex))
        else:
            ex = self.exact.numpy().view(np.float64)
            P = w.W * w.H
            for b in range(B):
                fam = P - ex[b * A:(b + 1) * A]
                best = int(np.argmax(fam))
                self.best_log.append(best)
                ang = (self.poses[b, 2] + w.offsets[best]) % (2 * np.pi)
                self.poses[b] = (self.poses[b, 0] + w.step_size * np.cos(ang),
                                 self.poses[b, 1] + w.step_size * np.sin(ang), ang)

    @staticmethod
    def _exact_diff(view, glimpse):
        d = 0.0
        for a, b in zip(view[..., 2].ravel().astype(int), glimpse[..., 2].ravel().astype(int)):
            d += abs(int(a) - int(b)) / 255.
        return d


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from cases import build_case
    from navsim.sharded import ShardedStepper
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, w, tpath, pose, frames = build_case("ties")        # tiny sensor: ties on nearly every step
    ow = O.World(L, **w)
    assert ow.train_from_path(tpath) == (0, -1)
    poses = [pose, (pose[0] + 3.0, pose[1] - 2.0, pose[2] + 0.2)]
    eng = OracleShardEngine(L, w, tpath, ow.scenes, rank, world, poses)
    st = ShardedStepper(eng, keys=eng.keys_tensor(), exact=eng.exact_tensor())
    steps = 6
    st.step(steps)
    # unsharded oracle trajectories
    ref_best = []
    agents = [ow.new_agent(*p) for p in poses]
    for s in range(steps):
        for a in agents:
            rc, best, _, _ = ow.step_forward(a, fake=True)
            assert rc == 0
            ref_best.append(best)
    ok = (eng.best_log == ref_best) and all(
        abs(eng.poses[i, 0] - agents[i].x) < 1e-12 and abs(eng.poses[i, 1] - agents[i].y) < 1e-12
        for i in range(len(poses)))
    q.put((rank, ok, eng.best_log, ref_best))
    dist.destroy_process_group()


def test_shard_bounds_and_keys():
    from navsim.sharded import pack_key, shard_bounds, unpack_key
    for n in (1, 7, 8, 1414, 10 ** 6):
        for world in (1, 2, 3, 8):
            parts = [shard_bounds(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert [c for _, c in parts] == [len(x) for x in np.array_split(np.arange(n), world)]
    k = pack_key([5, 5, 4], [10, 3, 999999])
    assert np.argmin(k) == 2 and np.argmin(k[:2]) == 1      # lower difference, then lower view index
    assert [list(x) for x in unpack_key(k)] == [[5, 5, 4], [10, 3, 999999]]
    assert np.all(k > 0) and EXACT_NONE > 0                # int64 MIN treats "none" as +infinity


def test_sharded_stepping_matches_unsharded_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, got, want in res:
        assert ok, (rank, got, want)
    assert res[0][2] == res[1][2]                            # identical decisions on every rank
