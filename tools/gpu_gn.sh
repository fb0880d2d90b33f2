#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in "4 0 8" "8 0 8" "4 1 8" "8 1 8" "2 0 8" "4 0 16" "4 0 4" "8 1 16"; do
  set -- $cfg
  echo "== U=$1 HINT=$2 CTAS=$3"
  CVVAE_GN_U=$1 CVVAE_GN_HINT=$2 CVVAE_GN_CTAS=$3 timeout 300 python tools/bench_gn.py 5 2>&1 | grep -v Warn
done
python - <<'PY'
import torch
x=torch.empty((1,17,576,576,128),dtype=torch.half,device='cuda'); y=torch.empty_like(x)
fl=torch.empty(512<<20,dtype=torch.uint8,device='cuda')
ts=[]
for _ in range(5):
    fl.zero_(); s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record(); y.copy_(x); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ts.sort(); print("torch copy_ 1.44GB:", ts[2], "ms", x.numel()*4/ts[2]/1e6, "GB/s")
PY
} 2>&1 | tee gpurun_out/gn_sweep.log
