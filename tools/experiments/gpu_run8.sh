mkdir -p gpurun_out
for cfg in "rolling 2 12" "rolling 1 12" "rolling 2 8" "waves 3 8"; do
  set -- $cfg
  timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --e2e-mode $1 --e2e-sets $2 --streams $3 > gpurun_out/r2h_$1_$2_$3.json 2> gpurun_out/r2h_$1_$2_$3.err
  python - "$1_$2_$3" <<'PY'
import json, sys
f = "r2h_" + sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    print(f, "value %.0f Mpix/s, %.1f ms/step" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2), d["e2e"]["includes"][-70:]))
except Exception as e:
    print(f, "failed", e)
PY
done
