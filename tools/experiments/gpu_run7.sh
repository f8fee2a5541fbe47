mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not c4_8192" 2>&1 | tail -4
J40B_TIMELINE=1 J40B_E2E_SKIP=d2h timeout 500 python bench.py --steps 24 --warmup 3 --skip-latency > gpurun_out/r2g_nod2h.json 2> gpurun_out/r2g_nod2h.err; grep "e2e host" gpurun_out/r2g_nod2h.err
J40B_TIMELINE=1 timeout 500 python bench.py --steps 24 --warmup 3 --skip-latency --e2e-sets 1 > gpurun_out/r2g_sets1.json 2> gpurun_out/r2g_sets1.err; grep "e2e host" gpurun_out/r2g_sets1.err
python - <<'PY'
import json
for f in ("r2g_nod2h", "r2g_sets1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2), d["e2e"]["includes"]))
    except Exception as e:
        print(f, "failed", e)
PY
