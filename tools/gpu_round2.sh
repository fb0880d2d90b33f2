#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt; }
run tests 900 python -m pytest tests -m gpu -q --tb=short
run bench_conv 900 python tools/bench_conv.py --reps 3
run bench 900 python bench.py --steps 3 --warmup 3
# launch list of one bench step (prepack ~130 launches + 3 warm-up steps skipped)
run ncu_launches 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1450 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
run ncu_full_n128 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/prof_conv_n128 python tools/bench_conv.py --reps 1 --only "E 128->128 333 @17"
run ncu_full_n256 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/prof_conv_n256 python tools/bench_conv.py --reps 1 --only "D 512->512 333 @9x288"
run ncu_full_gn 900 ncu --set full --clock-control none -k regex:gn_ -s 2 -c 2 -f -o gpurun_out/prof_gn python -c "
import torch,sys
sys.path.insert(0,'.')
from cvvae_b200.ops import CudaOps
o=CudaOps()
x=torch.randn(1,17,576,576,128,device='cuda').half()
g=torch.ones(128,device='cuda');b=torch.zeros(128,device='cuda')
for _ in range(2): y=o.groupnorm(x,g,b,32,1e-5)
torch.cuda.synchronize()
"
