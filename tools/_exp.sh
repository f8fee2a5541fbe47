timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_v8.json") if l.startswith("{")][-1])
print("value %.0f ms/step %.1f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v,1) for k,v in d["roofline"]["all_kernel_ms"].items()}, d["e2e"]["includes"][-70:])
PY
tail -2 gpurun_out/bench_v8.err
