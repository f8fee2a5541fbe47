mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --steps 24 --warmup 3 "$@" > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err; tail -1 gpurun_out/r2v_$tag.err; }
run lf6 --debug-skip 3 --streams 6
run lf24 --debug-skip 3 --streams 24
run lf256x3 --debug-skip 3 --streams 3 --frames-per-gpu 256 --steps 9
run all256x3 --skip-e2e --skip-latency --streams 3 --frames-per-gpu 256 --steps 9
python - <<'PY'
import json
for f in ("lf6","lf24","lf256x3","all256x3"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2v_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
