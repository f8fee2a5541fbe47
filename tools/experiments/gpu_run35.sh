mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lf_groups_sharing and 8" > gpurun_out/r3h_racecheck.log 2>&1; echo "rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r3h_racecheck.log
grep -E "Error: |Warning: " gpurun_out/r3h_racecheck.log | sed -E 's/\+0x[0-9a-f]+//g' | sort | uniq -c | sort -rn | head -40
