#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 6 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt; }
run tests 900 python -m pytest tests -m gpu -q --tb=short
run trace 600 python tools/trace_conv.py
run bench_conv 900 python tools/bench_conv.py --reps 3
run bench 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
