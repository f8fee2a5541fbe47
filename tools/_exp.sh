run() { # name, args
  name=$1; shift
  timeout 280 python bench.py "$@" > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/x_$name.json") if l.startswith("{")][-1])
    print("$name", "value %.0f ms/step %.1f serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]), {k: round(v,1) for k,v in d["roofline"]["all_kernel_ms"].items()})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/x_$name.err").read()[-800:])
PY
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run lfopt --steps 24 --warmup 3 --skip-e2e
