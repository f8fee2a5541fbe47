mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "vardct or each_transform or batch or wrap" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -3 gpurun_out/r2_final_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
timeout 600 python bench.py --steps 36 --warmup 3 --skip-latency --skip-e2e --preset light > gpurun_out/r2_final_light.json 2> gpurun_out/r2_final_light.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_final_traffic_list.csv \
    python bench.py --steps 1 --warmup 3 --skip-e2e --skip-latency --frames-per-gpu 32 --streams 1 > gpurun_out/r2_final_traffic.log 2>&1
python - <<'PY'
import json
for f in ("r2_final_bench", "r2_final_bench_reference", "r2_final_light"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value %.0f Mpix/s, %.1f ms/step" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e") and (round(d["e2e"]["value"]), d["e2e"].get("frac_of_ceiling")), d.get("latency"), d.get("roofline", {}).get("frac"), d.get("roofline", {}).get("step_frac"))
    except Exception as e:
        print(f, "failed", e)
PY
