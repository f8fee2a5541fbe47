# round 2, second measurement: parity suite with the lane-per-stream decoders, bench in both LF modes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not c4_8192" 2>&1 | tail -6
for mode in warp lane; do
  J40B_LF_MODE=$mode timeout 500 python bench.py --steps 24 --warmup 3 --skip-e2e --skip-latency > gpurun_out/r2b_$mode.json 2> gpurun_out/r2b_$mode.err; tail -2 gpurun_out/r2b_$mode.err
done
J40B_TIMELINE=1 timeout 500 python bench.py --steps 24 --warmup 3 --skip-latency --streams 24 > gpurun_out/r2b_lane24.json 2> gpurun_out/r2b_lane24.err; tail -30 gpurun_out/r2b_lane24.err
python - <<'PY'
import json
for f in ("r2b_warp", "r2b_lane", "r2b_lane24"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["ceiling_gbs_per_gpu"],1), round(d["e2e"]["frac_of_ceiling"],2)))
    except Exception as e:
        print(f, "failed", e)
PY
