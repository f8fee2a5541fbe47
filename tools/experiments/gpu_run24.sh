mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lf_groups_sharing" 2>&1 | tail -5
run() { tag=$1; shift; timeout 600 python bench.py --gpus 1 --steps 24 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r2w_$tag.json 2> gpurun_out/r2w_$tag.err; tail -1 gpurun_out/r2w_$tag.err; }
J40B_LF_LANES=32 run g32
J40B_LF_LANES=16 run g16
run g8
run g8_lfonly --debug-skip 3
python - <<'PY'
import json
for f in ("g32","g16","g8","g8_lfonly"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2w_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
