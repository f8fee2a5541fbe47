mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lf_groups_sharing or vardct or batch" 2>&1 | tail -3
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r3b_$tag.json 2> gpurun_out/r3b_$tag.err; tail -1 gpurun_out/r3b_$tag.err; }
run lfonly --steps 24 --debug-skip 3
run all12 --steps 24
run all16 --steps 32 --streams 16
J40B_LF_LANES=32 run all12_g32 --steps 24
python - <<'PY'
import json
for f in ("lfonly","all12","all16","all12_g32"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r3b_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
