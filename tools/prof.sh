# usage (on the GPU box): bash tools/prof.sh <tag> [kernels...]  -- writes text summaries to gpurun_out/ (reports are
# summarised on the box: three --set full reports with sources exceed gpurun's 64 MiB return limit)
tag=${1:-prof}; shift
kernels=${@:-k_lf_decode k_hf_group k_back_tile}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-e2e > gpurun_out/${tag}_bench_under_ncu.log 2>&1
reps=""
for k in $kernels; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/${tag}_$k \
      python bench.py --steps 1 --warmup 3 --skip-e2e --frames-per-gpu 16 --streams 1 > gpurun_out/${tag}_$k.log 2>&1
  python tools/ncu_hot_lines.py /tmp/${tag}_$k.ncu-rep 60 > gpurun_out/${tag}_${k}_hot.txt 2>&1
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  python tools/ncu_sass_mix.py /tmp/${tag}_$k.ncu-rep $k > gpurun_out/${tag}_${k}_sass.txt 2>&1
  reps="$reps /tmp/${tag}_$k.ncu-rep"
done
python - "$tag" $reps <<'PY'
import sys, subprocess, shutil, os
tag = sys.argv[1]
os.makedirs("profiles", exist_ok=True)
subprocess.run([sys.executable, "tools/ncu_summary.py", tag + "_box", f"gpurun_out/{tag}_launches.csv", *sys.argv[2:]], stdout=subprocess.DEVNULL)
shutil.copy(f"profiles/{tag}_box.txt", f"gpurun_out/{tag}_summary.txt")
PY
ls -la gpurun_out/
