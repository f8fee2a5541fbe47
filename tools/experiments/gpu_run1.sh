# round 2, first measurement on the box: parity suite, the bench on both presets (device-resident only), one end-to-end run
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 400 python bench.py --steps 12 --warmup 3 --skip-e2e --skip-latency --preset light > gpurun_out/r2a_light.json 2> gpurun_out/r2a_light.err; tail -3 gpurun_out/r2a_light.err
timeout 600 python bench.py --steps 24 --warmup 3 > gpurun_out/r2a_d1.json 2> gpurun_out/r2a_d1.err; tail -3 gpurun_out/r2a_d1.err
python - <<'PY'
import json
for f in ("r2a_light", "r2a_d1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"]), d.get("latency"))
    except Exception as e:
        print(f, "failed", e)
PY
