#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/$name.log" | cut -c1-900 | tee -a gpurun_out/summary.txt; }
# DRAM traffic of every launch of one step (2 metrics, one pass each)
run ncu_traffic 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 1250 -c 800 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
# C3: SD3 variant, 33x512x512, bf16 (2 chunks x 1 tile)
run bench_c3_sd3 900 python tools/bench_sd3.py
run smoke 600 python __graft_entry__.py smoke
run bench 900 python bench.py --steps 5 --warmup 3
