"""Do two independent half-batches on two streams overlap (distance kernel of one with the
latency-bound step kernels of the other)?  Compares one 1024-agent engine with two
512-agent engines stepping concurrently.  (tuning aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "navigation-by-deja-vu_b200"))
import numpy as np, torch
import bench
from navsim import NavEngine
L, tpath, poses, kw = bench.build_world_inputs(bench.WORKLOAD)
K = 100
def make(p, stream):
    e = NavEngine(L, device=0, stream=stream.cuda_stream, **kw)
    assert e.train_from_path(tpath) == (0, -1)
    e.set_agents(p)
    return e
def run(engines, streams, label):
    for e in engines: e.rewind(); e.step(K)
    torch.cuda.synchronize()
    for e in engines: e.rewind()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(K // 10):
        for e in engines: e.step(10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-40s %.2f us per 1024-agent step-batch" % (label, dt / K * 1e6))
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
parts = int(os.environ.get("PARTS", "2"))
streams = [torch.cuda.Stream() for _ in range(parts)]
whole = make(poses, s0)
run([whole], [s0], "one engine, 1024 agents")
n = len(poses) // parts
engs = [make(poses[i * n:(i + 1) * n], streams[i]) for i in range(parts)]
run(engs, streams, "%d engines x %d agents, %d streams" % (parts, n, parts))
run(engs[:1], streams[:1], "one engine of %d agents alone (x%d)" % (n, parts))
