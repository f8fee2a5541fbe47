mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r3f_$tag.json 2> gpurun_out/r3f_$tag.err; tail -1 gpurun_out/r3f_$tag.err; }
J40B_LF_SPLIT=1 run split --steps 24
J40B_LF_SPLIT=1 run split16 --steps 32 --streams 16
J40B_HF_LANES=16 run hf16 --steps 24
python - <<'PY'
import json
for f in ("split","split16","hf16"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r3f_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
