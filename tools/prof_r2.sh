# usage (on the GPU box): bash tools/prof_r2.sh <tag> <kernel regex> [env assignments...]  -- one ncu --set full capture of the first
# launch of the kernel after warm-up in a small bench run; text summaries into gpurun_out/
tag=$1; k=$2; shift; shift
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o /tmp/${tag} \
    python bench.py --steps 1 --warmup 3 --skip-e2e --skip-latency --frames-per-gpu 32 --streams 1 > gpurun_out/${tag}.log 2>&1
python tools/ncu_hot_lines.py /tmp/${tag}.ncu-rep 70 > gpurun_out/${tag}_hot.txt 2>&1
python tools/ncu_sass_mix.py /tmp/${tag}.ncu-rep $k > gpurun_out/${tag}_sass.txt 2>&1
python tools/ncu_summary.py ${tag}_box /tmp/${tag}.ncu-rep > /dev/null 2>&1; cp profiles/${tag}_box.txt gpurun_out/${tag}_summary.txt
tail -32 gpurun_out/${tag}_summary.txt | head -34
