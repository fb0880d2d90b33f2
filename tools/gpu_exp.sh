#!/bin/bash
# Ad-hoc experiment session (kernel variants through the environment knobs); logs in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary_exp.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary_exp.txt
  local t0=$(date +%s)
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $? ($(( $(date +%s) - t0 )) s)" | tee -a gpurun_out/summary_exp.txt
  tail -n 3 "gpurun_out/$name.log" | cut -c1-600 | tee -a gpurun_out/summary_exp.txt; }
run tests_pack 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "pack_taps or groupnorm"
run ov_base 300 python tools/exp_overlap.py
CVVAE_CONV_SMEM_RESERVE=8192 run ov_r8 300 python tools/exp_overlap.py
CVVAE_CONV_SMEM_RESERVE=8192 CVVAE_GN_CTAS_PER_SM=1 run ov_r8_g1 300 python tools/exp_overlap.py
CVVAE_CONV_SMEM_RESERVE=8192 CVVAE_GN_CTAS_PER_SM=2 run ov_r8_g2 300 python tools/exp_overlap.py
CVVAE_CONV_SMEM_RESERVE=16384 CVVAE_GN_CTAS_PER_SM=2 run ov_r16_g2 300 python tools/exp_overlap.py
run bench 600 python bench.py --steps 5 --warmup 3 --no-torch-baseline --no-cpu-baseline
