mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r2z_$tag.json 2> gpurun_out/r2z_$tag.err; tail -1 gpurun_out/r2z_$tag.err; }
J40B_CARVEOUT_LF=40 run lf40_lfonly --steps 24 --debug-skip 3
J40B_CARVEOUT_LF=40 J40B_LF_LANES=32 run lf40_g32_lfonly --steps 24 --debug-skip 3
J40B_CARVEOUT_LF=40 J40B_CARVEOUT_HF=40 run lfhf40_notile --steps 24 --debug-skip 1
J40B_CARVEOUT_LF=40 J40B_CARVEOUT_HF=40 run lfhf40_all --steps 24
python - <<'PY'
import json
for f in ("lf40_lfonly","lf40_g32_lfonly","lfhf40_notile","lfhf40_all"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2z_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
