"""View-sharded stepping on real kernels: two engines on one B200, each holding half
of the library, reduced with an element-wise minimum (what the NCCL MIN all-reduce
does across GPUs).  Must reproduce the unsharded engine exactly."""
import numpy as np
import pytest

from cases import agent_grid, build_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1_small", "ties"])
def test_two_shards_one_gpu(gpu, name):
    import torch
    import navsim
    from navsim.sharded import engine_reduction_tensors, shard_bounds
    L, w, tpath, pose, frames = build_case(name)
    frames = 40
    full = navsim.NavEngine(L, **w)
    assert full.train_from_path(tpath) == (0, -1)
    scenes = full.familiar_scenes
    poses = np.vstack([np.asarray(pose)[None], agent_grid(tpath, w, 2, 2)])
    full.set_agents(poses, frames)
    full.step(frames, log_afam=True)
    want = full.log(0, frames, afam=True)
    want_state = full.state()

    shards = []
    for r in range(2):
        off, cnt = shard_bounds(len(scenes), 2, r)
        e = navsim.NavEngine(L, **w)
        e.set_library_shard(scenes[off:off + cnt], off, len(scenes), tpath)
        e.set_agents(poses, frames)
        shards.append(e)
    bufs = [engine_reduction_tensors(e) for e in shards]
    for _ in range(frames):
        for which, idx in ((1, 0), (2, 1)):
            for e in shards:
                e.phase(which, log_afam=True)
                e.sync()
            m = torch.minimum(bufs[0][idx], bufs[1][idx])
            bufs[0][idx].copy_(m)
            bufs[1][idx].copy_(m)
            torch.cuda.synchronize()
        for e in shards:
            e.phase(3, log_afam=True)
            e.sync()
    for e in shards:
        got = e.log(0, frames, afam=True)
        st = e.state()
        assert np.array_equal(got["best_idx"], want["best_idx"])
        assert np.array_equal(got["poses"], want["poses"])          # same kernels, same bits
        assert np.allclose(got["afam"], want["afam"], rtol=1e-12, atol=0, equal_nan=True)
        assert np.array_equal(st["status"], want_state["status"])
        assert np.array_equal(st["coverage"], want_state["coverage"])
