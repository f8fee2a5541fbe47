timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
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
ARGS="--steps 12 --warmup 3 --streams 6 --skip-e2e"
run s6_default A=1
run s6_nostage_nocap J40B_LF_STAGE=0 J40B_LF_CAP=0
run s6_nostage_cap J40B_LF_STAGE=0 J40B_LF_CAP=256
run s6_warps4 J40B_LF_WARPS=4
ARGS="--steps 12 --warmup 3 --streams 3 --skip-e2e"
run s3_nostage_nocap J40B_LF_STAGE=0 J40B_LF_CAP=0
ARGS="--steps 6 --warmup 3 --streams 3"
run full_default A=1
