run() { # name, args
  name=$1; shift
  timeout 400 python bench.py "$@" > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/x_$name.json") if l.startswith("{")][-1])
    print("$name", "value %.0f ms/step %.1f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"] if d["e2e"] else 0), d["e2e"]["includes"][-60:], d["device_bytes"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/x_$name.err").read()[-800:])
PY
  grep "e2e host" gpurun_out/x_$name.err
}
export J40B_TIMELINE=1
( time run sets2 --steps 24 --warmup 3 ) 2>&1 | grep -v "^$\|user\|sys"
run sets2s6 --steps 24 --warmup 3 --streams 6
