run() { # name, args
  name=$1; shift
  timeout 280 python bench.py "$@" > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/x_$name.json"))
    print("$name", "value %.0f ms/step %.1f serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]), {k: round(v,1) for k,v in d["roofline"]["kernel_ms"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/x_$name.err").read()[-800:])
PY
  grep timeline gpurun_out/x_$name.err | head -${TL:-0}
  grep "phase" gpurun_out/x_$name.err | grep -v "total=0" | head -2
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
J40B_PHASE_DUMP=1 run s6l0 --steps 12 --warmup 3 --skip-e2e --lag 0
J40B_LF_STAGE=1 run s6l0stage --steps 12 --warmup 3 --skip-e2e --lag 0
run s12l0 --steps 24 --warmup 3 --skip-e2e --streams 12 --lag 0
