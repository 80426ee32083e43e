"""BASELINE.json configs[3] (C4): a 10^6-view library, B = 1 agent x 10 headings.

  python tools/c4_sharded_bench.py                       # one GPU, whole library
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29514 tools/c4_sharded_bench.py      # views sharded over N GPUs, NVLink P2P exchange

Views: the genuine training-path views first, then synthetic views drawn from the level
alphabet.  Reports comparisons/s, the distance kernel's time and the HBM
rate it implies (algorithmic bytes = local views x P bytes per step)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "navigation-by-deja-vu_b200")):
    sys.path.insert(0, p)

import numpy as np
import torch

from navsim import NavEngine, synthetic
from navsim.sharded import shard_bounds


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    n_total = int(os.environ.get("C4_VIEWS", 1000000))
    B = int(os.environ.get("C4_AGENTS", 1))
    steps = 50
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = synthetic.make_landscape(4001, 2000, sigma=6.0)
    kw = dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=10.0, n_test_angles=10,
              n_sensor_levels=5, max_distance_to_training_path=450)
    tpath = synthetic.training_path_for(L.shape, 10.0, 10, 0.0)
    eng = NavEngine(L, device=local, **kw)
    assert eng.train_from_path(tpath) == (0, -1)
    genuine = eng.familiar_scenes
    rng = np.random.default_rng(4700)
    levels = np.array([0, 63, 127, 191, 255], np.uint8)
    off, cnt = shard_bounds(n_total, world, rank)
    scenes = np.zeros((cnt, 2, 40, 3), np.uint8)
    scenes[..., 2] = levels[np.random.default_rng(4700 + rank).integers(0, 5, (cnt, 2, 40))]
    if off < len(genuine):
        k = min(len(genuine) - off, cnt)
        scenes[:k] = genuine[off:off + k]
    path = np.vstack([tpath, np.repeat(tpath[-1:], n_total - len(tpath), axis=0)])
    eng.set_library_shard(scenes, off, n_total, path) if world > 1 else eng.set_library(scenes, path)
    spw = 80
    poses = synthetic.start_pose_grid(tpath, spw, n_lat=int(np.sqrt(B)), n_deg=B // int(np.sqrt(B))) if B > 1 \
        else np.array([synthetic.start_pose(tpath, (0.05, 3.0), spw)])
    eng.set_agents(poses)
    if world > 1:
        eng.p2p_attach(rank, world)
    eng.step(5)
    eng.sync()
    eng.rewind()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.step(steps)
    eng.sync()
    dt = time.perf_counter() - t0
    k2_ms = eng.time_distance_kernel(20)
    st = eng.state(coverage=False)
    eng.rewind()
    tl = eng.timeline(8)   # (every rank queues the same steps: the exchanges need all of them)
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        A, P = 10, 80
        print(json.dumps({
            "workload": "C4: %d views, %d agent(s) x %d headings, P=%d" % (n_total, len(poses), A, P),
            "n_gpus": world, "views_per_gpu": cnt, "us_per_step": dt / steps * 1e6,
            "comparisons_per_sec": len(poses) * A * n_total * steps / dt,
            "k2_us": k2_ms * 1e3, "k2_hbm_gbs": cnt * P / (k2_ms * 1e-3) / 1e9,
            "exchange": "NVLink P2P (k_p2p_min)" if world > 1 else "none",
            "p2p_error": eng.p2p_error() if world > 1 else 0,
            "agents_still_running": int((st["status"] == 0).sum()),
            "step_timeline_us": {k: [round(x, 2) for x in v] if isinstance(v, tuple) else round(v, 2) for k, v in tl.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
