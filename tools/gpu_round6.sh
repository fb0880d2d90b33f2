#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 6 "gpurun_out/$name.log" | cut -c1-600 | tee -a gpurun_out/summary.txt; }
run tests 900 python -m pytest tests -m gpu -q --tb=short
run bench 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
run ncu_launches 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
run ncu_full_n128 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/prof2_conv_n128 python tools/bench_conv.py --reps 1 --only "E 128->128 333 @17"
run ncu_full_n256 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/prof2_conv_n256 python tools/bench_conv.py --reps 1 --only "D 512->512 333 @9x288"
