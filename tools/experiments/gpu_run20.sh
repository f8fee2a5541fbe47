mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 24 --warmup 3 --skip-latency > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err; tail -3 gpurun_out/r2_final_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 12 --warmup 3 --skip-latency --skip-e2e --scaling strong --total-frames 128 > gpurun_out/r2_final_bench_2gpu_strong128.json 2> gpurun_out/r2_final_bench_2gpu_strong128.err; tail -3 gpurun_out/r2_final_bench_2gpu_strong128.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_final_bench_2gpu_ref.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_final_bench_2gpu", "r2_final_bench_2gpu_strong128", "r2_final_bench_2gpu_ref"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d.get("n_gpus"), d.get("scaling"), "value %.0f Mpix/s, %.1f ms/step" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e") and (round(d["e2e"]["value"]), d["e2e"].get("frac_of_ceiling"), d["e2e"].get("ceiling_gbs_per_gpu")))
    except Exception as e:
        print(f, "failed", e)
PY
