mkdir -p gpurun_out
J40B_LF_MODE=lane timeout 700 python bench.py --steps 96 --warmup 3 --skip-latency --skip-e2e --streams 24 > gpurun_out/r2f_lane24_96.json 2> gpurun_out/r2f_lane24_96.err; tail -2 gpurun_out/r2f_lane24_96.err
timeout 500 python bench.py --steps 48 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2f_warp12_48.json 2> gpurun_out/r2f_warp12_48.err; tail -2 gpurun_out/r2f_warp12_48.err
python - <<'PY'
import json
for f in ("r2f_lane24_96", "r2f_warp12_48"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, d["device_bytes"]/1e9)
    except Exception as e:
        print(f, "failed", e)
PY
