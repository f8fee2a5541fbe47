# usage (on the GPU box, e.g. through gpurun): bash tools/gpu_check.sh
# the round's closing check: GPU parity suite, one short device-resident bench run, smoke()
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 12 --warmup 3 --skip-e2e > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_check.json") if l.startswith("{")][-1])
print("value %.0f Mpix/s, %.1f ms/step" % (d["value"], d["ms_per_step"]), {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()})
PY
tail -2 gpurun_out/bench_check.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
