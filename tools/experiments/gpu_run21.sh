mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -2 gpurun_out/r2t_bench.err
python - <<'PY'
import json
for f in ("r2t_bench",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "value %.0f Mpix/s, %.1f ms/step" % (d["value"], d["ms_per_step"]), "serial", round(r["serial_ms_per_step"],1), {k: round(v,1) for k,v in r["all_kernel_ms"].items()}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items()}, "e2e", d.get("e2e") and (round(d["e2e"]["value"]), d["e2e"].get("frac_of_ceiling")), d.get("latency"))
    except Exception as e:
        print(f, "failed", e)
PY
