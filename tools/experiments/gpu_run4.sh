mkdir -p gpurun_out
timeout 500 python bench.py --steps 24 --warmup 3 --skip-latency > gpurun_out/r2d_warp.json 2> gpurun_out/r2d_warp.err; tail -2 gpurun_out/r2d_warp.err
for mode in warp lane; do
  J40B_LF_MODE=$mode timeout 900 python bench.py --workload c4 --steps 6 --warmup 3 --streams 4 --skip-e2e --skip-latency > gpurun_out/r2d_c4_$mode.json 2> gpurun_out/r2d_c4_$mode.err; tail -2 gpurun_out/r2d_c4_$mode.err
done
python - <<'PY'
import json
for f in ("r2d_warp", "r2d_c4_warp", "r2d_c4_lane"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2)))
    except Exception as e:
        print(f, "failed", e)
PY
