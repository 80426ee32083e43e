#!/bin/bash
# One GPU-box pass that regenerates everything under profiles/ for a round:
#   tools/evidence_run.sh r01        (run through gpurun; results land in gpurun_out/)
# Order matters: every ncu pass repeats a command that has already exited 0 without ncu.
# (one launch per full capture: two reports of two launches each exceed gpurun's 64 MiB return limit)
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
if [ -z "$SKIP_PYTEST" ]; then
  python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
  tail -2 $out/${tag}_pytest_gpu.log
fi
python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref_err.log; echo "ref rc=$?"
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_err.log; echo "bench rc=$?"
cat $out/${tag}_bench_n1.json
python bench.py --steps 20 --warmup 3 --quick --no-cpu > $out/${tag}_plain.log 2>&1; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
      --log-file $out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --quick --no-cpu > $out/${tag}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:k2_ -s 40 -c 1 -f \
      -o $out/${tag}_k2 python bench.py --steps 20 --warmup 3 --quick --no-cpu > $out/${tag}_ncu2.log 2>&1
  echo "ncu k2 rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:k3_ -s 40 -c 1 -f \
      -o $out/${tag}_k3 python bench.py --steps 20 --warmup 3 --quick --no-cpu > $out/${tag}_ncu3.log 2>&1
  echo "ncu k3 rc=$?"
fi
