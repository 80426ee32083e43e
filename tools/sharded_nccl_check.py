"""View-sharded stepping over real NCCL (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29513 tools/sharded_nccl_check.py

Every rank holds a contiguous slice of the library; the two MIN all-reduces of
navsim/sharded.py run over NCCL.  The result must equal the unsharded engine's."""
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
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
