mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "vardct or batch or each_transform or lane" 2>&1 | tail -3
timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2o_d1.json 2> gpurun_out/r2o_d1.err; tail -2 gpurun_out/r2o_d1.err
J40B_HF_STAGE=0 timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2o_d1_nostage.json 2> gpurun_out/r2o_d1_nostage.err; tail -2 gpurun_out/r2o_d1_nostage.err
python - <<'PY'
import json
for f in ("r2o_d1", "r2o_d1_nostage"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, {k: round(v, 1) for k, v in d["roofline"]["stage_ms_in_region"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
