#!/bin/bash
# Fast GPU check of a step-path change: trajectory parity, the bench-size C2 parity test,
# then a short bench line (tools/quick_check.sh [extra pytest -k expression])
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_workloads.py tests/test_gpu_golden.py -m gpu -x -q -k "${1:-trajectories or out_of_bounds or c2_bench or golden or dropin or step_io}" > $out/qc_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $out/qc_pytest.log
python bench.py --steps 100 --warmup 5 --no-cpu > $out/qc_bench.json 2> $out/qc_bench_err.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/qc_bench.json"))
r=d["roofline"]
print("cold us/step", d["ms_per_step"]*1e3, "warm", d["ms_per_step_l2_warm"]*1e3, "e2e", d["e2e"]["ms_per_step"]*1e3)
print("k2", r["kernel"][:12], "events us", r["launch_ms"]*1e3, "alone", r["launch_ms_alone"]*1e3, "in graph", r.get("launch_ms_in_graph"))
print("timeline", r["step_timeline_us"])
PY
