#!/bin/bash
# One GPU session: per-operator parity, end-to-end parity, smoke, bench.  Every stage runs in its own
# process under `timeout` so that a trapping kernel cannot take the rest of the session with it.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 6 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run t_probe 300 python -m pytest tests/test_gpu_ops.py -q -x -k "library or probe" -s
run t_conv_tc 600 python -m pytest tests/test_gpu_ops.py -q -k "conv_tc" --tb=short
run t_conv_other 600 python -m pytest tests/test_gpu_ops.py -q -k "conv_direct or attention_style" --tb=short
run t_norm_misc 600 python -m pytest tests/test_gpu_ops.py -q -k "groupnorm or layernorm or data_movement" --tb=short
run t_e2e 900 python -m pytest tests/test_gpu_e2e.py -q --tb=short
run smoke 600 python __graft_entry__.py smoke
run bench 900 python bench.py --steps 2 --warmup 3
