#!/bin/bash
# One GPU session under gpurun: parity tests, per-layer conv microbench, smoke, bench.  Every stage runs in its own
# process under `timeout` so that a trapping kernel cannot take the rest of the session with it; logs land in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh [profile]'
# With `profile`: also the ncu launch list (+ DRAM bytes) of one bench step and `--set full` captures of the conv kernels
# (the summaries under profiles/ are produced from those).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/$name.log" | cut -c1-900 | tee -a gpurun_out/summary.txt; }
run tests 1200 python -m pytest tests -m gpu -q --tb=short
run bench_conv 900 python tools/bench_conv.py --reps 3
run smoke 600 python __graft_entry__.py smoke
run bench 900 python bench.py --steps 5 --warmup 3
if [ "${1:-}" = "profile" ]; then
  run ncu_launches 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -s 1250 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
  run ncu_full_psw 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_psw -s 1 -c 1 -f \
      -o gpurun_out/prof_conv_psw python tools/bench_conv.py --reps 1 --only "D 128->128 333 @17"
  run ncu_full_n256 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f \
      -o gpurun_out/prof_conv_n256 python tools/bench_conv.py --reps 1 --only "D 512->512 333 @9x288"
fi
