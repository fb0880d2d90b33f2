cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -q -x 2>&1 | tail -2
timeout 600 python tools/bench_conv.py --reps 3 2>&1 | grep '"layer"' | cut -c1-120 | tee gpurun_out/bench_conv.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
