mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 500 python bench.py --workload c2 --steps 12 --warmup 3 --skip-latency > gpurun_out/r2i_c2.json 2> gpurun_out/r2i_c2.err; tail -2 gpurun_out/r2i_c2.err
timeout 500 python bench.py --scaling strong --total-frames 32 --steps 6 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2i_strong32.json 2> gpurun_out/r2i_strong32.err; tail -2 gpurun_out/r2i_strong32.err
python - <<'PY'
import json
for f in ("r2i_c2", "r2i_strong32"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2)), d["config"]["workload"][:60])
    except Exception as e:
        print(f, "failed", e)
PY
