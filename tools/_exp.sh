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
ARGS="--steps 12 --warmup 3"
run v8_default A=1
ARGS="--steps 5 --warmup 3"
run v8_default_s5 A=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lf_decode" -c 2 -o gpurun_out/prof_v8 python bench.py --steps 1 --warmup 1 --frames-per-gpu 8 --skip-e2e --streams 1 > gpurun_out/ncu_v8.log 2>&1; tail -1 gpurun_out/ncu_v8.log | cut -c1-200
