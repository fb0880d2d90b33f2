cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "ours exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?"
tail -n 1 gpurun_out/bench_reference.log | cut -c1-600
