"""View-sharded library across GPUs (BASELINE.json configs[3]).

Not in the reference (it keeps the whole library in one process,
navsim/NavBySceneFamiliarity.py:122); this is the one place on the hot path
with a real exchange step.  Every rank holds the landscape, ALL agents and a
contiguous slice of the training views.  Per step-batch:

  phase 1  every rank samples the same glimpses and scores them against its slice:
           one packed 64-bit key (difference << idx_bits | GLOBAL view index) per
           (agent, heading)                                   -> all-reduce MIN
  phase 2  every rank evaluates the exact FP64 difference of the views it owns
           among the winners and the tied candidates (+inf elsewhere)
                                                              -> all-reduce MIN
  phase 3  every rank applies the identical decision and moves its copy of the
           agents: no broadcast back, the states stay bit-identical.

Keys and exact differences are non-negative when read as int64, so a plain
MIN reduction on int64 works (torch.distributed / NCCL ncclMin); the lowest
global view index wins ties deterministically.

Two ways to run the exchange: `ShardedStepper` below (host-driven phases, a
torch.distributed MIN all-reduce between them), or `NavEngine.p2p_attach()` after
which `NavEngine.step()` runs the whole sharded sequence on the device with the
exchange done by a kernel over NVLink peer memory (csrc/step.cuh, k_p2p_min).

`ShardedStepper` drives any engine-like object exposing phase(k),
keys_tensor() and exact_tensor(); NavEngine provides them on the GPU, the CPU
tests plug in a stand-in to exercise this logic over gloo.
"""
import numpy as np


def shard_bounds(n_total, world_size, rank):
    """Contiguous split like np.array_split (the reference's own static split,
    scripts/run_experiment.py:327): the first n_total % world_size shards get one
    more view.  Returns (offset, count)."""
    base, rem = divmod(int(n_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def pack_key(score, view_index, idx_bits=32):
    return (np.asarray(score, dtype=np.int64) << idx_bits) | np.asarray(view_index, dtype=np.int64)


def unpack_key(key, idx_bits=32):
    key = np.asarray(key, dtype=np.int64)
    return key >> idx_bits, key & ((1 << idx_bits) - 1)


class _DevicePtr(object):
    """Minimal __cuda_array_interface__ holder so torch can wrap an engine buffer."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2}


def engine_reduction_tensors(engine):
    """(keys, exact) int64 CUDA tensors aliasing the engine's two reduction buffers."""
    import torch
    from . import _cabi
    n = engine.n_agents * engine.n_test_angles
    dev = "cuda:%d" % engine.device
    keys = torch.as_tensor(_DevicePtr(engine.device_ptr(_cabi.PTR_KEYS), n), device=dev)
    exact = torch.as_tensor(_DevicePtr(engine.device_ptr(_cabi.PTR_TIE), n), device=dev)
    return keys, exact


class ShardedStepper(object):
    """Steps one engine per rank in lock step, MIN-reducing the two buffers."""

    def __init__(self, engine, group=None, keys=None, exact=None):
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.dist = dist
        if keys is None:
            keys, exact = engine_reduction_tensors(engine)
        self.keys, self.exact = keys, exact
        # phase() only ENQUEUES kernels on the engine's stream (by default a non-blocking stream
        # of its own) while the all-reduce runs on torch's current stream: the two are ordered
        # explicitly around every reduction.  Engines without a CUDA stream (the CPU stand-in
        # of the gloo tests) need nothing.
        self._engine_stream = None
        if keys.is_cuda:
            import torch
            handle = getattr(engine, "stream_handle", None)
            if handle is None:
                raise ValueError("a CUDA engine must expose stream_handle for the reduction ordering")
            self._torch = torch
            self._engine_stream = torch.cuda.ExternalStream(handle, device=keys.device)

    def _all_reduce_min(self, buf):
        dist = self.dist
        es = self._engine_stream
        if es is None:
            dist.all_reduce(buf, op=dist.ReduceOp.MIN, group=self.group)
            return
        cur = self._torch.cuda.current_stream(buf.device)
        same = cur.cuda_stream == es.cuda_stream
        if not same:
            cur.wait_stream(es)          # the kernels that wrote `buf` are done before NCCL reads it
        dist.all_reduce(buf, op=dist.ReduceOp.MIN, group=self.group)
        if not same:
            es.wait_stream(cur)          # the reduced values have landed before the next phase reads them

    def step(self, nsteps=1, fake=False, log_afam=False):
        for _ in range(int(nsteps)):
            self.engine.phase(1, fake=fake, log_afam=log_afam)
            self._all_reduce_min(self.keys)
            self.engine.phase(2, fake=fake, log_afam=log_afam)
            self._all_reduce_min(self.exact)
            self.engine.phase(3, fake=fake, log_afam=log_afam)


def make_sharded_engine(engine_cls, scenes, path, rank, world_size, *args, **kwargs):
    """Builds an engine holding this rank's slice of `scenes` (N, H, W, 3) while
    keeping the whole training `path` (update_error needs every point)."""
    off, cnt = shard_bounds(len(scenes), world_size, rank)
    eng = engine_cls(*args, **kwargs)
    eng.set_library_shard(scenes[off:off + cnt], off, len(scenes), path)
    return eng
