mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "vardct or batch or passes" 2>&1 | tail -3
timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency > gpurun_out/r2j_d1.json 2> gpurun_out/r2j_d1.err; tail -2 gpurun_out/r2j_d1.err
timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e --streams 16 > gpurun_out/r2j_d1_s16.json 2> gpurun_out/r2j_d1_s16.err; tail -2 gpurun_out/r2j_d1_s16.err
timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e --preset light > gpurun_out/r2j_light.json 2> gpurun_out/r2j_light.err; tail -2 gpurun_out/r2j_light.err
python - <<'PY'
import json
for f in ("r2j_d1", "r2j_d1_s16", "r2j_light"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2)))
    except Exception as e:
        print(f, "failed", e)
PY
