#!/bin/bash
# One GPU session under gpurun: parity tests, per-layer conv microbench, smoke, bench.  Every stage runs in its own
# process under `timeout` so that a trapping kernel cannot take the rest of the session with it; logs land in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh [profile] [quick]'
# With `profile`: also the ncu launch list (+ DRAM bytes) of one bench step and `--set full` captures of the conv kernels
# (the summaries under profiles/ are produced from those).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  local t0=$(date +%s)
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $? ($(( $(date +%s) - t0 )) s)" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/$name.log" | cut -c1-1500 | tee -a gpurun_out/summary.txt; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import bench; print(bench.source_hash())" > gpurun_out/source_hash.txt 2>/dev/null
run tests 1500 python -m pytest tests -m gpu -q --tb=short -x
run smoke 600 python __graft_entry__.py smoke
run bench 1200 python bench.py --steps 5 --warmup 3
run bench_conv 900 python tools/bench_conv.py --reps 3
for a in "$@"; do
if [ "$a" = "profile" ]; then
  run ncu_launches 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -s 300 -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline
  run ncu_full_psw 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_psw -s 1 -c 1 -f \
      -o gpurun_out/prof_conv_psw python tools/bench_conv.py --reps 1 --only "D 128->128 333 @17"
  run ncu_full_n256 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f \
      -o gpurun_out/prof_conv_n256 python tools/bench_conv.py --reps 1 --only "D 512->512 333 @9x288"
fi
if [ "$a" = "configs" ]; then
  run bench_c3 900 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline
  run bench_c4 900 python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline
  run bench_c5 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline
fi
done
