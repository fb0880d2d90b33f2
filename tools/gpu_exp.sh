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
  tail -n 6 "gpurun_out/$name.log" | cut -c1-600 | tee -a gpurun_out/summary_exp.txt; }
run tests 1500 python -m pytest tests -m gpu -q --tb=short
run conv_wide1 300 python tools/bench_conv.py --reps 3 --only "128 "
CVVAE_CONV_WIDE=0 run conv_wide0 300 python tools/bench_conv.py --reps 3 --only "128 "
CVVAE_CONV_PW=16 run conv_pw16 300 python tools/bench_conv.py --reps 3 --only "128 "
CVVAE_CONV_PW=12 run conv_pw12 300 python tools/bench_conv.py --reps 3 --only "128 "
run conv_5x72 300 python tools/bench_conv.py --reps 3 --only "@5x72"
CVVAE_CONV_NACC=1 run conv_5x72_nacc1 300 python tools/bench_conv.py --reps 3 --only "@5x72"
run bench 1200 python bench.py --steps 5 --warmup 3 --no-torch-baseline --no-cpu-baseline
CVVAE_CONV_WIDE=0 run bench_wide0 1200 python bench.py --steps 5 --warmup 3 --no-torch-baseline --no-cpu-baseline
run racecheck 900 compute-sanitizer --tool racecheck --racecheck-detect-level info python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "test_conv_tc_matches_spec and pair_n256_odd"
