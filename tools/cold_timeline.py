import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "navigation-by-deja-vu_b200"))
import numpy as np, torch
import bench
from navsim import NavEngine
L, tpath, poses, kw = bench.build_world_inputs(bench.WORKLOAD)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
eng = NavEngine(L, device=0, stream=stream.cuda_stream, **kw)
assert eng.train_from_path(tpath) == (0, -1)
eng.set_agents(poses)
eng.step(30); eng.sync()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def show(tag, tl):
    print(tag, {k: (tuple(round(x, 2) for x in v) if isinstance(v, tuple) else round(v, 2)) for k, v in tl.items()})
for rep in range(2):
    show("warm eager", eng.timeline(1))
    flush.fill_(rep); torch.cuda.synchronize()
    show("cold eager", eng.timeline(1))
