"""Turns the raw files of tools/evidence_run.sh (gpurun_out/<tag>_*) into the tracked
summaries under profiles/:  python tools/summarize_profiles.py r01
Needs `ncu` on PATH to read the .ncu-rep files (no GPU needed)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(tag):
    rows = [r for r in csv.reader(open(os.path.join(OUT, tag + "_launches.csv"))) if len(r) > 14 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r[4], []).append(float(r[14]) / 1e3)
    return per


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    per = launches(tag)
    shutil.copy(os.path.join(OUT, tag + "_launches.csv"), os.path.join(PROF, tag + "_launches.csv"))
    bench = json.load(open(os.path.join(OUT, tag + "_bench_n1.json")))
    with open(os.path.join(PROF, tag + "_launch_summary.txt"), "w") as f:
        f.write("# per-kernel device time from profiles/%s_launches.csv\n" % tag)
        f.write("# (ncu --metrics gpu__time_duration.sum --clock-control none on `bench.py --steps 20 --warmup 3 --quick --no-cpu`;\n")
        f.write("#  cold-cache, serialised, no programmatic-dependent-launch overlap: compare SHARES, not absolutes)\n\n")
        step = {}
        for k, v in per.items():
            f.write("%-72s n=%4d mean=%8.2f us\n" % (k[:72], len(v), sum(v) / len(v)))
            for key in ("k2_tc", "k2_sad", "k3_decide", "k3_ties", "k3_move_sample", "k3_step_tm"):
                if key in k and len(v) > 8:   # (the kernels of the steady-state step-batch, not the one-off first step)
                    step[key] = sum(v) / len(v)
        tot = sum(step.values())
        f.write("\nkernels launched >= 8 times (steady-state step-batch and the roofline legs) = " + " + ".join(step) + " = %.1f us under ncu\n" % tot)
        f.write("shares: " + ", ".join("%s %.1f %%" % (k, 100 * v / tot) for k, v in step.items()) + "\n")
        r = bench["roofline"]
        f.write("bench.py (CUDA events, same build, no profiler): K2 %.1f us of a %.1f us L2-warm / %.1f us cold-L2 step-batch = %.1f %% / %.1f %%\n"
                % (r["launch_ms"] * 1e3, bench["ms_per_step_l2_warm"] * 1e3, bench["ms_per_step"] * 1e3,
                   100 * r["launch_ms"] / bench["ms_per_step_l2_warm"], 100 * r["launch_ms"] / bench["ms_per_step"]))
        if r.get("launch_ms_in_graph"):
            f.write("inside the replayed graph (global-timer stamps written by the kernels, bench.py step_timeline_us): K2 %.1f us "
                    "of a %.1f us step-batch = %.1f %%; timeline (first CTA resident, first dependency met, last CTA done): %s\n"
                    % (r["launch_ms_in_graph"] * 1e3, bench["ms_per_step_l2_warm"] * 1e3,
                       100 * r["launch_ms_in_graph"] / bench["ms_per_step_l2_warm"], json.dumps(r.get("step_timeline_us"))))
        f.write("(events around K2 in the eager timing pass include its launch latency, which the graph replay hides behind\n"
                " the previous kernel; the k3 kernels gain more from a warm L2 than K2 does: hence the larger live share)\n")
    traffic = None
    traffic_kernel = ""
    with open(os.path.join(PROF, tag + "_ncu_summary.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, C2 workload (bench.py --quick), one launch each\n")
        for rep in (tag + "_k2.ncu-rep", tag + "_k3.ncu-rep"):
            path = os.path.join(OUT, rep)
            if not os.path.exists(path):
                continue
            head, units, rows = raw_page(path)
            # every metric of the capture, one line per metric (first launch): the summaries can be re-derived
            with open(os.path.join(PROF, rep.replace(".ncu-rep", "_raw_metrics.csv")), "w") as g:
                w = csv.writer(g)
                w.writerow(["metric", "unit"] + ["launch_%d" % i for i in range(len(rows))])
                for i, m in enumerate(head):
                    w.writerow([m, units[i]] + [row[i] for row in rows])
            stall = [m for m in head if m.startswith("smsp__average_warps_issue_stalled_") and m.endswith("_per_issue_active.ratio")]
            for row in rows:
                f.write("\n== %s  (%s)\n" % (row[head.index("Kernel Name")], rep))
                for m in METRICS:
                    if m in head:
                        i = head.index(m)
                        f.write("  %-72s %s %s\n" % (m, row[i], units[i]))
                f.write("  warp stall reasons (warps stalled per issued instruction, largest first):\n")
                for m, v in sorted(((m, float(row[head.index(m)] or 0)) for m in stall), key=lambda t: -t[1])[:8]:
                    f.write("    %-40s %.3f\n" % (m[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
                name = row[head.index("Kernel Name")]
                if "k2_tc" in name and traffic is None:
                    rd, wr = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
                    traffic = float(row[rd]) * UNIT_SCALE[units[rd]] + float(row[wr]) * UNIT_SCALE[units[wr]]
                    traffic_kernel = name.split("(")[0]
                if ("k3_step_tm" in name or "k3_move_sample" in name) and "smsp__inst_executed.sum" in head:
                    json.dump({"kernel": "k3_step_tm" if "k3_step_tm" in name else "k3_move_sample",
                               "warp_instructions": float(row[head.index("smsp__inst_executed.sum")]),
                               "source": "profiles/%s_ncu_summary.txt (ncu --set full, smsp__inst_executed.sum, one launch on C2)" % tag},
                              open(os.path.join(PROF, "step_kernel_inst.json"), "w"))
    if traffic is not None:
        json.dump({"dram_bytes_per_launch": traffic,
                   "source": "profiles/%s_ncu_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                             "one launch of %s on C2, L2 flushed before the step)" % (tag, traffic_kernel)},
                  open(os.path.join(PROF, "k2_traffic.json"), "w"))
    for name in ("_bench_n1.json", "_bench_reference_arm.json"):
        shutil.copy(os.path.join(OUT, tag + name), os.path.join(PROF, tag + name))
    tail = open(os.path.join(OUT, tag + "_pytest_gpu.log")).read().strip().splitlines()[-1]
    open(os.path.join(PROF, tag + "_pytest_gpu.txt"), "w").write("python -m pytest tests -m gpu -x -q  (B200, this build)\n" + tail + "\n")
    print("traffic", traffic)


if __name__ == "__main__":
    main()
