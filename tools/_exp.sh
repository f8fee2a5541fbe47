timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # name, env..., args
  name=$1; shift
  env "$@" timeout 280 python bench.py $ARGS > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/x_$name.json"))
    print("$name", "value %.0f ms/step %.1f serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]), {k: round(v,1) for k,v in d["roofline"]["kernel_ms"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/x_$name.err").read()[-800:])
PY
}
ARGS="--steps 12 --warmup 3"
run v10_default A=1
ARGS="--steps 12 --warmup 3 --skip-e2e"
run v10_lanes24 J40B_HF_LANES=24
run v10_lanes32 J40B_HF_LANES=32
run v10_lanes8 J40B_HF_LANES=8
ARGS="--steps 16 --warmup 3 --skip-e2e --streams 8"
run v10_s8 A=1
