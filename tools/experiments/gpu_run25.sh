mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py --gpus 1 --warmup 3 --skip-e2e --skip-latency "$@" > gpurun_out/r2x_$tag.json 2> gpurun_out/r2x_$tag.err; tail -1 gpurun_out/r2x_$tag.err; }
J40B_LF_LANES=32 run big32 --frames-per-gpu 256 --obj-frames 256 --streams 2 --steps 4
run big8 --frames-per-gpu 256 --obj-frames 256 --streams 2 --steps 4
run big8x4 --frames-per-gpu 256 --obj-frames 256 --streams 4 --steps 8
python - <<'PY'
import json
for f in ("big32","big8","big8x4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2x_{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
PROF_BENCH_ARGS="--frames-per-gpu 256 --obj-frames 256 --streams 1" bash tools/prof_r2.sh r2x_lfwp_big32 'k_lf_chan<\(int\)1, \(int\)32>' 0 J40B_LF_LANES=32
PROF_BENCH_ARGS="--frames-per-gpu 256 --obj-frames 256 --streams 1" bash tools/prof_r2.sh r2x_lfwp_big8 'k_lf_chan<\(int\)1, \(int\)8>' 0
