mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r3c_$tag.json 2> gpurun_out/r3c_$tag.err; grep "lf smhist" gpurun_out/r3c_$tag.err | tail -1 | cut -c1-2600; }
J40B_LF_SMHIST=1 run h8 --steps 24 --debug-skip 3
J40B_LF_SMHIST=1 J40B_LF_LANES=32 run h32 --steps 24 --debug-skip 3
python - <<'PY'
import json
for f in ("h8","h32"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r3c_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step" % d["ms_per_step"], "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
