cd /root/repo; mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_ops.py -q -x -k "not forced" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | head
grep -E "Error: Race|and (Read|Write) access" gpurun_out/sanitize_racecheck.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -20
