#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/$name.log" | cut -c1-600 | tee -a gpurun_out/summary.txt; }
run tests_conv 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "conv"
run tests 900 python -m pytest tests -m gpu -q --tb=short
run bench_conv 600 python tools/bench_conv.py --reps 3 --only "128 "
run bench 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
