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
}
run s3 --steps 12 --warmup 3 --skip-e2e --streams 3
run s12 --steps 24 --warmup 3 --skip-e2e --streams 12
run f32s12 --steps 24 --warmup 3 --skip-e2e --streams 12 --frames-per-gpu 32
run f128s6 --steps 12 --warmup 3 --skip-e2e --streams 6 --frames-per-gpu 128
run f16s24 --steps 48 --warmup 3 --skip-e2e --streams 24 --frames-per-gpu 16
