cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_ops.py -q -x -k "groupnorm or layernorm or softmax or temporal or data_movement or blend or video_pre or stacked" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | head
