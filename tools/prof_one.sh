tag=r1_v8; k=k_lf_decode
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/${tag}_$k \
    python bench.py --steps 1 --warmup 3 --skip-e2e --frames-per-gpu 16 --streams 1 > gpurun_out/${tag}_$k.log 2>&1
python tools/ncu_hot_lines.py /tmp/${tag}_$k.ncu-rep 60 > gpurun_out/${tag}_${k}_hot.txt 2>&1
python tools/ncu_sass_mix.py /tmp/${tag}_$k.ncu-rep $k > gpurun_out/${tag}_${k}_sass.txt 2>&1
python tools/ncu_summary.py ${tag}_box /tmp/${tag}_$k.ncu-rep > /dev/null 2>&1; cp profiles/${tag}_box.txt gpurun_out/${tag}_summary.txt
tail -32 gpurun_out/${tag}_summary.txt | head -12
