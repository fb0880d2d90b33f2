#!/bin/bash
# compute-sanitizer passes over the small-shape GPU tests (memcheck: out-of-bounds / misaligned; synccheck: barrier misuse)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in ${SANITIZE_TOOLS:-memcheck synccheck}; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -q -x -k "not forced and not golden and not graph and not stress" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"
  grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned|Barrier" gpurun_out/sanitize_$tool.log | head -20
done
