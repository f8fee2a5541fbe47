timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/c4_check.py 4096 > gpurun_out/c4_4096.json 2> gpurun_out/c4_4096.err; cat gpurun_out/c4_4096.json; tail -2 gpurun_out/c4_4096.err
