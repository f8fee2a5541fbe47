mkdir -p gpurun_out
export J40B_TEST_SHORT=1
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lf_groups_sharing or wide_tokens or passes2_all" > gpurun_out/r3g_$tool.log 2>&1; echo "rc=$?"
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/r3g_$tool.log | head -8
done
