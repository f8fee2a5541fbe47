mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r3a_$tag.json 2> gpurun_out/r3a_$tag.err; grep timeline gpurun_out/r3a_$tag.err | head -30; }
echo "== 24 objects, one wave, LF only"
J40B_TIMELINE=1 run tl24 --steps 24 --streams 24 --debug-skip 3
echo "== same, 32 connections"
CUDA_DEVICE_MAX_CONNECTIONS=32 J40B_TIMELINE=1 run tl24c32 --steps 24 --streams 24 --debug-skip 3
echo "== 12 objects, 2 waves, everything"
J40B_TIMELINE=1 run tl12 --steps 24 --streams 12
python - <<'PY'
import json
for f in ("tl24","tl24c32","tl12"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r3a_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
