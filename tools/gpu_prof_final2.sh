mkdir -p gpurun_out
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_final2_launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-e2e --skip-latency > gpurun_out/r2_final2_bench_under_ncu.log 2>&1
python tools/ncu_summary.py r2_final2_launches gpurun_out/r2_final2_launches.csv > /dev/null 2>&1; cp profiles/r2_final2_launches.txt gpurun_out/
# DRAM bytes per launch of one decode of 32 frames
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final2_traffic_list.csv \
    python bench.py --steps 1 --warmup 3 --skip-e2e --skip-latency --frames-per-gpu 32 --streams 1 > gpurun_out/r2_final2_traffic_under_ncu.log 2>&1
python tools/ncu_traffic_from_list.py gpurun_out/r2_final2_traffic_list.csv 32 d1 && cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
bash tools/prof_r2.sh r2_final2_lf_chan_wp8 'k_lf_chan<\(int\)1, \(int\)8>' 7
bash tools/prof_r2.sh r2_final2_hf 'k_hf_group' 2
bash tools/prof_r2.sh r2_final2_tile 'k_back_tile' 2
