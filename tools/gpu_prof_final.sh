mkdir -p gpurun_out
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_final_launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-e2e --skip-latency > gpurun_out/r2_final_bench_under_ncu.log 2>&1
python tools/ncu_summary.py r2_final_launches gpurun_out/r2_final_launches.csv > /dev/null 2>&1; cp profiles/r2_final_launches.txt gpurun_out/
bash tools/prof_r2.sh r2_final_lf_chan_wp 'k_lf_chan<1>' 7
bash tools/prof_r2.sh r2_final_lf_chan_wide 'k_lf_chan<3>' 9
bash tools/prof_r2.sh r2_final_hf 'k_hf_group' 2
bash tools/prof_r2.sh r2_final_tile 'k_back_tile' 2
