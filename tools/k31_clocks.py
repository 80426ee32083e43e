"""Where does the fused step+sample kernel spend its time? (tuning aid)"""
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
eng.step(5); eng.sync()
out2 = np.zeros((2, len(poses), 8), np.int64)
out = out2[0]
_cabi.check(eng._lib.nvb_debug_step_clocks(eng._h, _cabi.ptr(out2)))
ok = out[:, 6] > 0
d = out[ok]
names = (["start->active", "active->decided", "decided->moved", "moved->window requested", "requested->landed", "landed->sampled"]
         if os.environ.get("NAVSIM_B200_STEP_FORM") == "1" else
         ["resident->dependency met", "dependency met->pose", "pose->moved (scan, bookkeeping, rotations)", "moved->gather entered", "entered->window landed", "landed->sampled"])
for i, n in enumerate(names):
    x = (d[:, i + 1] - d[:, i]) / 1.965e3
    print("%-26s mean %6.2f us  p10 %6.2f  p90 %6.2f" % (n, x.mean(), np.percentile(x, 10), np.percentile(x, 90)))
if os.environ.get("NAVSIM_B200_STEP_FORM") != "1":
    f = out2[1][ok]
    base = d[:, 2]   # pose published
    for i, n in enumerate(["window requested", "warp 1 hook done", "warp 0 scanned", "warp 2 scanned", "reduced", "bookkeeping done"]):
        got = f[:, i] > 0
        if not got.any():
            continue   # stamp not on the path taken (e.g. the full scan when the block prefilter runs)
        x = (f[got, i] - base[got]) / 1.965e3
        print("   pose -> %-20s mean %6.2f us  p10 %6.2f  p90 %6.2f" % (n, x.mean(), np.percentile(x, 10), np.percentile(x, 90)))
tot = (d[:, 6] - d[:, 0]) / 1.965e3
print("total per CTA  mean %.2f us  max %.2f us   (n=%d)" % (tot.mean(), tot.max(), len(d)))
order = np.argsort(-tot)[:8]
print("slowest CTAs (us per phase):")
for i in order:
    print("  ", np.round((d[i, 1:7] - d[i, 0:6]) / 1.965e3, 2), "total %.1f" % tot[i], "tie-scan exact evaluations:", d[i, 7])
# kernel-level: spread of start and end clocks is per-SM, so only durations are comparable
hist, edges = np.histogram(tot, bins=[0, 10, 15, 20, 25, 30, 40, 60])
print("total-duration histogram:", dict(zip(["<10", "<15", "<20", "<25", "<30", "<40", "<60"], hist)))

print("careful-path blocks per agent (of %d): mean %.2f max %d" % (800, d[:, 7].mean(), d[:, 7].max()))
print("agents with tie scans:", int((d[:, 7] > 0).sum()), " exact evaluations: mean %.1f max %d" % (d[d[:, 7] > 0, 7].mean(), d[:, 7].max()))
