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
run bench_conv 900 python tools/bench_conv.py --reps 3
run ncu_full_psw 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_psw -s 1 -c 1 -f -o gpurun_out/prof3_conv_psw python tools/bench_conv.py --reps 1 --only "D 128->128 333 @17"
run ncu_launches 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1250 -c 800 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
run bench_c3_sd3 900 python tools/bench_sd3.py
run bench 900 python bench.py --steps 5 --warmup 3
