timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; tail -c 1500 gpurun_out/bench_v7.json; tail -2 gpurun_out/bench_v7.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v7_ref.json 2>> gpurun_out/bench_v7.err; cut -c1-300 gpurun_out/bench_v7_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
