timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/c4_check.py 8192 > gpurun_out/c4_check.json 2> gpurun_out/c4_check.err; cat gpurun_out/c4_check.json; tail -2 gpurun_out/c4_check.err
