mkdir -p gpurun_out
timeout 500 python bench.py --steps 24 --warmup 3 --skip-latency > gpurun_out/r2e_hf32.json 2> gpurun_out/r2e_hf32.err; tail -2 gpurun_out/r2e_hf32.err
J40B_HF_LANES=16 timeout 500 python bench.py --steps 24 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2e_hf16.json 2> gpurun_out/r2e_hf16.err; tail -2 gpurun_out/r2e_hf16.err
timeout 500 python bench.py --steps 24 --warmup 3 --skip-latency --skip-e2e --streams 16 > gpurun_out/r2e_hf32_s16.json 2> gpurun_out/r2e_hf32_s16.err; tail -2 gpurun_out/r2e_hf32_s16.err
python - <<'PY'
import json
for f in ("r2e_hf32", "r2e_hf16", "r2e_hf32_s16"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2)))
    except Exception as e:
        print(f, "failed", e)
PY
