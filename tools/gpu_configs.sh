#!/bin/bash
# BASELINE configs 4 and 5 on a 2-GPU box: 33-frame 720p clip frame-sharded over 2 ranks; batch-32 256^2 sweep at N = 1, 2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name"; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $?"; tail -n 2 "gpurun_out/$name.log" | cut -c1-1200; }
run c5_n1 600 python tools/bench_batch.py
run c5_n2 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_batch.py
run c4_720p_n1 600 python bench.py --gpus 1 --height 720 --width 1280 --steps 3 --warmup 3 --no-cpu-baseline
run c4_720p_n2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --height 720 --width 1280 --steps 3 --warmup 3
