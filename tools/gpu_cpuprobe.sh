#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
cat /sys/fs/cgroup/cpu.max 2>/dev/null
for n in 16 32 64 128; do timeout 120 python tools/cpu_threads_probe.py $n; done
for n in 32 128; do OMP_NUM_THREADS=1 timeout 120 python tools/cpu_threads_probe.py $n; done
OMP_PROC_BIND=close OMP_PLACES=cores timeout 120 python tools/cpu_threads_probe.py 64
} 2>&1 | grep -v Warning | tee gpurun_out/cpu_probe.log
