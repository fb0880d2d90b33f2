cd /root/repo; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -q -x -k full_size_parity 2>&1 | tail -15
cat gpurun_out/fullsize_parity_*.json
