# usage (on the GPU box): bash tools/prof_r2.sh <tag> <kernel regex (demangled name)> <launches to skip> [env assignments...]
# one ncu --set full capture of one launch of the kernel in a small bench run; text summaries into gpurun_out/
tag=$1; k=$2; skip=$3; shift; shift; shift
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$k" -s $skip -c 1 -f -o /tmp/${tag} \
    python bench.py --steps 1 --warmup 3 --skip-e2e --skip-latency ${PROF_BENCH_ARGS:---frames-per-gpu 32 --streams 1} > gpurun_out/${tag}.log 2>&1
python tools/ncu_hot_lines.py /tmp/${tag}.ncu-rep 70 > gpurun_out/${tag}_hot.txt 2>&1
python tools/ncu_sass_mix.py /tmp/${tag}.ncu-rep "" > gpurun_out/${tag}_sass.txt 2>&1
python tools/ncu_summary.py ${tag}_box /tmp/${tag}.ncu-rep > /dev/null 2>&1; cp profiles/${tag}_box.txt gpurun_out/${tag}_summary.txt
head -12 gpurun_out/${tag}_summary.txt
