"""Where does one step-batch go?  Wall-clock (globaltimer) span of every step kernel inside the
replayed graph: first CTA resident -> dependency met -> last CTA done.  (tuning aid)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "navigation-by-deja-vu_b200"))
import numpy as np
import bench
from navsim import NavEngine, _cabi
L, tpath, poses, kw = bench.build_world_inputs(bench.WORKLOAD)
eng = NavEngine(L, **kw)
assert eng.train_from_path(tpath) == (0, -1)
eng.set_agents(poses)
eng.step(20); eng.sync()
out = np.zeros((4, 2048, 3), np.int64)
_cabi.check(eng._lib.nvb_debug_timeline(eng._h, 12, _cabi.ptr(out)))
names = ["k2", "k3_decide", "k3_ties", "k3_move_sample"]
t0 = None
rows = []
for k, name in enumerate(names):
    d = out[k]
    ok = d[:, 0] > 0
    if not ok.any():
        continue
    done = d[ok][:, 2]
    done = done[done > 0]
    rows.append((name, int(ok.sum()), d[ok][:, 0].min(), d[ok][:, 0].max(), d[ok][:, 1].min(), d[ok][:, 1].max(),
                 done.min() if len(done) else 0, done.max() if len(done) else 0))
t0 = min(r[2] for r in rows)
print("%-16s %5s  %9s %9s  %9s %9s  %9s %9s   (us since the first stamp of the step-batch)" %
      ("kernel", "CTAs", "res.first", "res.last", "dep.first", "dep.last", "done.first", "done.last"))
for r in rows:
    print("%-16s %5d  " % (r[0], r[1]) + "  ".join("%9.2f %9.2f" % ((r[i] - t0) / 1e3, (r[i + 1] - t0) / 1e3) for i in (2, 4, 6)))
d = out[0]
ok = d[:, 0] > 0
done = (d[ok][:, 2] - t0) / 1e3
idx = np.nonzero(ok)[0]
print("k2 done time by CTA index (us):")
for lo in range(0, len(done), 37):
    seg = done[lo:lo + 37]
    print("  CTA %3d-%3d  min %6.2f mean %6.2f max %6.2f" % (idx[lo], idx[min(lo + 36, len(done) - 1)], seg.min(), seg.mean(), seg.max()))
hist, edges = np.histogram(done, bins=12)
print("k2 done histogram:", [(round(float(e), 1), int(h)) for h, e in zip(hist, edges[:-1])])
for k in (1, 3):
    d = out[k]; ok = d[:, 2] > 0
    dn = (d[ok][:, 2] - t0) / 1e3
    hist, edges = np.histogram(dn, bins=10)
    print(names[k], "done histogram:", [(round(float(e), 1), int(h)) for h, e in zip(hist, edges[:-1])])
