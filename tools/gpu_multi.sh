#!/bin/bash
# Multi-GPU session:  gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_multi.sh N [full]'
# NCCL parity tests (N >= 2), then bench.py through torchrun exactly as the driver launches it; one JSON line per run in
# gpurun_out/scale_<config>_n<N>.log.  `full`: every N in {1,2,4,8} <= the box size for c2 (weak) / c4 / c5 (strong).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=${1:-2}
: > gpurun_out/summary_multi.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary_multi.txt
  local t0=$(date +%s)
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $? ($(( $(date +%s) - t0 )) s)" | tee -a gpurun_out/summary_multi.txt
  grep -E '^\{|passed|failed|Error' "gpurun_out/$name.log" | tail -n 2 | cut -c1-700 | tee -a gpurun_out/summary_multi.txt; }
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run nccl_tests 900 python -m pytest tests/test_gpu_parallel_nccl.py -q -m gpu --tb=short
if [ "${2:-}" = "full" ]; then
  run scale_c2_n1 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-torch-baseline
  run scale_c4_n1 900 python bench.py --gpus 1 --config c4 --steps 2 --warmup 3 --no-cpu-baseline
  run scale_c5_n1 600 python bench.py --gpus 1 --config c5 --steps 3 --warmup 3 --no-cpu-baseline
  for n in 2 4 8; do
    [ "$n" -le "$NG" ] || continue
    run scale_c2_n$n 600 bash tools/trun.sh $n --steps 10 --warmup 3
    run scale_c4_n$n 900 bash tools/trun.sh $n --config c4 --steps 2 --warmup 3
    run scale_c5_n$n 600 bash tools/trun.sh $n --config c5 --steps 3 --warmup 3
  done
  run scale_c3_n${NG} 600 bash tools/trun.sh ${NG} --config c3 --steps 5 --warmup 3
  run scale_c4unit33_n${NG} 600 bash tools/trun.sh ${NG} --config c4 --frames 33 --shard unit --steps 3 --warmup 3
else
  run scale_c2_n1 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-torch-baseline
  run scale_c2_n${NG} 600 bash tools/trun.sh ${NG} --steps 10 --warmup 3
  run scale_c2unit_n2 600 bash tools/trun.sh 2 --steps 10 --warmup 3 --shard unit
  run scale_c4_n${NG} 900 bash tools/trun.sh ${NG} --config c4 --steps 2 --warmup 3
  run scale_c5_n${NG} 600 bash tools/trun.sh ${NG} --config c5 --steps 3 --warmup 3
  run scale_c3_n${NG} 600 bash tools/trun.sh ${NG} --config c3 --steps 5 --warmup 3
fi
