run() { # name, args
  name=$1; shift
  timeout 280 python bench.py "$@" > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/x_$name.json"))
    print("$name", "value %.0f ms/step %.1f serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]), {k: round(v,1) for k,v in d["roofline"]["all_kernel_ms"].items()})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/x_$name.err").read()[-800:])
PY
}
run base --steps 3 --warmup 3 --skip-e2e --streams 1
J40B_LF_CARVEOUT=100 run co100 --steps 3 --warmup 3 --skip-e2e --streams 1
J40B_LF_CARVEOUT=100 J40B_LF_STAGE=1 run co100stage --steps 3 --warmup 3 --skip-e2e --streams 1
J40B_LF_STAGE=1 run stage --steps 3 --warmup 3 --skip-e2e --streams 1
