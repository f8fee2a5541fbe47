mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r2y_$tag.json 2> gpurun_out/r2y_$tag.err; tail -1 gpurun_out/r2y_$tag.err; }
run c100 --steps 24
J40B_CARVEOUT=-1 run cdef --steps 24
J40B_LF_LANES=32 run c100_g32 --steps 24
run c100_lfonly --steps 24 --debug-skip 3
run c100_big8x4 --frames-per-gpu 256 --obj-frames 256 --streams 4 --steps 8
python - <<'PY'
import json
for f in ("c100","cdef","c100_g32","c100_lfonly","c100_big8x4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2y_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
