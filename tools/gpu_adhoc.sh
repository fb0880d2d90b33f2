cd /root/repo; mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name"; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $?"; tail -n 1 "gpurun_out/$name.log" | cut -c1-420; }
run bench_n4 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3
run bench_n1b 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline
run bench_torch_cuda 900 python bench.py --impl torch-cuda --steps 3 --warmup 3 --no-cpu-baseline
