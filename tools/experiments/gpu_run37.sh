mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_vardct or lf_groups_sharing or large_batch" 2>&1 | tail -2
timeout 200 python bench.py --gpus 1 --steps 24 --warmup 3 --skip-e2e --skip-latency > gpurun_out/r3j_hfpad.json 2> gpurun_out/r3j_hfpad.err; tail -1 gpurun_out/r3j_hfpad.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r3j_hfpad.json") if l.startswith("{")][-1])
r = d["roofline"]
print("%.1f ms/step, %.0f Mpix/s" % (d["ms_per_step"], d["value"]), "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
PY
