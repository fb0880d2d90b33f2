#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 8 "gpurun_out/$name.log" | cut -c1-600 | tee -a gpurun_out/summary.txt; }
run tests_new 600 python -m pytest tests -m gpu -q --tb=short -x -k "ragged or golden or image or graph"
run graphs 900 python tools/bench_graphs.py
