"""View-sharded stepping over real NCCL (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29513 tools/sharded_nccl_check.py

Every rank holds a contiguous slice of the library.  Three exchanges are checked against
the unsharded engine, bit for bit: the two MIN all-reduces of navsim/sharded.py over NCCL
with the engine on torch's stream, the same with a DEFAULT-constructed engine (its own
non-blocking stream: ShardedStepper orders the streams itself), and the device-resident
exchange over NVLink peer memory (csrc/step.cuh, nvb_p2p_min_agent) replayed as a CUDA
graph.  Last, one rank queues a step its peers do not: the exchange must time out into
p2p_error() == 1 instead of hanging.  tests/test_gpu_multi.py runs this file when the box
has at least two GPUs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

from cases import agent_grid, build_case
from navsim import NavEngine
from navsim.sharded import ShardedStepper, shard_bounds


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ok = True
    for name, frames in (("c1_small", 60), ("ties", 60)):
        L, w, tpath, pose, _ = build_case(name)
        full = NavEngine(L, device=local, **w)
        assert full.train_from_path(tpath) == (0, -1)
        scenes = full.familiar_scenes
        poses = np.vstack([np.asarray(pose)[None], agent_grid(tpath, w, 3, 3)])
        full.set_agents(poses, frames)
        full.step(frames)
        want = full.log(0, frames)
        off, cnt = shard_bounds(len(scenes), world, rank)
        eng = NavEngine(L, device=local, stream=stream.cuda_stream, **w)
        eng.set_library_shard(scenes[off:off + cnt], off, len(scenes), tpath)
        eng.set_agents(poses, frames)
        st = ShardedStepper(eng)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.step(frames)
        eng.sync()
        dt = time.perf_counter() - t0
        got = eng.log(0, frames)
        same = np.array_equal(got["best_idx"], want["best_idx"]) and np.array_equal(got["poses"], want["poses"])
        ok = ok and same
        print("rank %d %-8s views [%d, %d) of %d: NCCL all-reduce  %s  (%.1f us/step, host-driven phases)" %
              (rank, name, off, off + cnt, len(scenes), "identical to unsharded" if same else "MISMATCH",
               dt / frames * 1e6), flush=True)
        # a default-constructed engine: its own non-blocking stream, NCCL on torch's stream
        eng_d = NavEngine(L, device=local, **w)
        eng_d.set_library_shard(scenes[off:off + cnt], off, len(scenes), tpath)
        eng_d.set_agents(poses, frames)
        ShardedStepper(eng_d).step(frames)
        eng_d.sync()
        got_d = eng_d.log(0, frames)
        same_d = np.array_equal(got_d["best_idx"], want["best_idx"]) and np.array_equal(got_d["poses"], want["poses"])
        ok = ok and same_d
        print("rank %d %-8s NCCL all-reduce, engine on its own stream: %s" %
              (rank, name, "identical to unsharded" if same_d else "MISMATCH"), flush=True)
        # the same shards, exchanged over NVLink peer memory inside the step sequence
        eng2 = NavEngine(L, device=local, stream=stream.cuda_stream, **w)
        eng2.set_library_shard(scenes[off:off + cnt], off, len(scenes), tpath)
        eng2.set_agents(poses, frames)
        eng2.p2p_attach(rank, world)
        eng2.step(3)                      # warm-up: plain launches + graph capture
        eng2.sync()
        eng2.rewind()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng2.step(frames)
        eng2.sync()
        dt2 = time.perf_counter() - t0
        got2 = eng2.log(0, frames)
        same2 = (np.array_equal(got2["best_idx"], want["best_idx"]) and np.array_equal(got2["poses"], want["poses"])
                 and eng2.p2p_error() == 0)
        ok = ok and same2
        print("rank %d %-8s views [%d, %d) of %d: NVLink P2P exchange %s  (%.1f us/step, device-resident)" %
              (rank, name, off, off + cnt, len(scenes), "identical to unsharded" if same2 else "MISMATCH",
               dt2 / frames * 1e6), flush=True)
    # a peer that does not show up: rank 0 queues one step more than the others
    os.environ["NAVSIM_B200_P2P_SPIN"] = "200000000"      # ~0.1 s of SM clock instead of ~2 s
    L, w, tpath, pose, _ = build_case("c1_small")
    full = NavEngine(L, device=local, **w)
    assert full.train_from_path(tpath) == (0, -1)
    scenes = full.familiar_scenes
    off, cnt = shard_bounds(len(scenes), world, rank)
    eng3 = NavEngine(L, device=local, **w)
    eng3.set_library_shard(scenes[off:off + cnt], off, len(scenes), tpath)
    eng3.set_agents(np.asarray(pose)[None], 50)
    eng3.p2p_attach(rank, world)
    eng3.step(4)
    eng3.sync()
    err_before = eng3.p2p_error()
    dist.barrier()
    t0 = time.perf_counter()
    if rank == 0:
        eng3.step(1)
        eng3.sync()
    waited = time.perf_counter() - t0
    err_after = eng3.p2p_error()
    timeout_ok = err_before == 0 and (err_after == 1 if rank == 0 else err_after == 0) and waited < 30.0
    ok = ok and timeout_ok
    print("rank %d missing peer: error flag %d -> %d after %.2f s  %s" %
          (rank, err_before, err_after, waited, "ok" if timeout_ok else "WRONG"), flush=True)
    dist.barrier()
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
