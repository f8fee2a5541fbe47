mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lane or wrap or intermediate or known_answer" 2>&1 | tail -4
for mode in lane; do
  J40B_LF_MODE=$mode timeout 500 python bench.py --steps 24 --warmup 3 --skip-e2e --skip-latency --streams 24 > gpurun_out/r2c_$mode.json 2> gpurun_out/r2c_$mode.err; tail -2 gpurun_out/r2c_$mode.err
done
python - <<'PY'
import json
for f in ("r2c_lane",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
