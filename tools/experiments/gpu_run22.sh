mkdir -p gpurun_out
for sk in 1 2 4 3 6; do
timeout 600 python bench.py --gpus 1 --steps 24 --warmup 3 --debug-skip $sk > gpurun_out/r2u_skip$sk.json 2> gpurun_out/r2u_skip$sk.err; tail -1 gpurun_out/r2u_skip$sk.err
done
python - <<'PY'
import json
for f in ("r2u_skip1","r2u_skip2","r2u_skip4","r2u_skip3","r2u_skip6"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(f, "%.1f ms/step" % d["ms_per_step"], "alone", {k: round(v,1) for k,v in r["all_kernel_ms"].items() if v}, "in-region", {k: round(v,1) for k,v in r["stage_ms_in_region"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
