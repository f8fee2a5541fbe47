mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not c4_8192" 2>&1 | tail -3
timeout 500 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e > gpurun_out/r2q_d1.json 2> gpurun_out/r2q_d1.err; tail -2 gpurun_out/r2q_d1.err
python - <<'PY'
import json
for f in ("r2q_d1",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step, serial %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["serial_ms_per_step"]),
              {k: round(v, 1) for k, v in d["roofline"]["all_kernel_ms"].items()}, {k: round(v, 1) for k, v in d["roofline"]["stage_ms_in_region"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
bash tools/prof_r2.sh r2_final_lf_chan_wp 'k_lf_chan<\(int\)1>' 7
bash tools/prof_r2.sh r2_final_lf_chan_gen 'k_lf_chan<\(int\)5>' 9
