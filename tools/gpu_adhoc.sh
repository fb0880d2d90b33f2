cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parallel_nccl.py -q -x 2>&1 | tail -15
