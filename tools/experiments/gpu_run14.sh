mkdir -p gpurun_out
cp j40_b200/libj40b200.so /tmp/orig.so
for v in tile64; do
  cp exp_libs/$v.so j40_b200/libj40b200.so
  timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2n_$v.json 2> gpurun_out/r2n_$v.err; tail -2 gpurun_out/r2n_$v.err
done
cp /tmp/orig.so j40_b200/libj40b200.so
python - <<'PY'
import json
for f in ("r2n_tile64",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, {k: round(v, 1) for k, v in d["roofline"]["stage_ms_in_region"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
