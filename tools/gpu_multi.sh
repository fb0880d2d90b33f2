#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary_multi.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary_multi.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary_multi.txt
  tail -n 3 "gpurun_out/$name.log" | cut -c1-1500 | tee -a gpurun_out/summary_multi.txt; }
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run bench_n1 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline
run bench_n2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3
N=${1:-2}
if [ "$N" -gt 2 ]; then
run bench_n$N 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3
run bench_ref_n$N 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 0
fi
