mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not c4_8192" 2>&1 | tail -4
timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency > gpurun_out/r2l_d1.json 2> gpurun_out/r2l_d1.err; tail -2 gpurun_out/r2l_d1.err
timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e --streams 16 > gpurun_out/r2l_d1_s16.json 2> gpurun_out/r2l_d1_s16.err; tail -2 gpurun_out/r2l_d1_s16.err
python - <<'PY'
import json
for f in ("r2l_d1", "r2l_d1_s16"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, {k: round(v, 1) for k, v in d["roofline"]["stage_ms_in_region"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2)))
    except Exception as e:
        print(f, "failed", e)
PY
