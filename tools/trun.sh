#!/bin/bash
# torchrun launch of bench.py exactly as the driver does it:  tools/trun.sh N [bench.py args]
n=$1; shift
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus "$n" "$@"
