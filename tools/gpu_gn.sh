#!/bin/bash
# GroupNorm apply/stats bandwidth at the decoder's largest tensors next to torch's copy_ on the same box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 300 python tools/bench_gn.py 5 2>&1 | grep -v Warn
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
} 2>&1 | tee gpurun_out/gn_bench.log
