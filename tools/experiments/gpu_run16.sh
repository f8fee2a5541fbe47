mkdir -p gpurun_out
J40B_TILE_PERSIST=3 timeout 300 python -m pytest tests -m gpu -x -q -k "vardct or each_transform" 2>&1 | tail -2
for p in 2 3; do
J40B_TILE_PERSIST=$p timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2p_persist$p.json 2> gpurun_out/r2p_persist$p.err; tail -2 gpurun_out/r2p_persist$p.err
done
python - <<'PY'
import json
for f in ("r2p_persist2", "r2p_persist3"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, {k: round(v, 1) for k, v in d["roofline"]["stage_ms_in_region"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
