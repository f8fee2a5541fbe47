mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 36 --warmup 3 > gpurun_out/r2s_split.json 2> gpurun_out/r2s_split.err; tail -2 gpurun_out/r2s_split.err
J40B_LF_SPLIT=0 timeout 600 python bench.py --steps 36 --warmup 3 --skip-e2e --skip-latency > gpurun_out/r2s_nosplit.json 2> gpurun_out/r2s_nosplit.err; tail -2 gpurun_out/r2s_nosplit.err
python - <<'PY'
import json
for f in ("r2s_split", "r2s_nosplit"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["frac_of_ceiling"],2)), d.get("latency"))
    except Exception as e:
        print(f, "failed", e)
PY
