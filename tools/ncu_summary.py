#!/usr/bin/env python
"""Summarises ncu artefacts brought back in gpurun_out/ into small text files under profiles/ (tracked).
usage: python tools/ncu_summary.py <tag> [launches.csv] [prof_x.ncu-rep ...]"""
import csv
import subprocess
import sys
from collections import OrderedDict

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_misc_per_issue_active.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_active.avg"]


def main():
    tag = sys.argv[1]
    out = []
    for path in sys.argv[2:]:
        if path.endswith(".csv"):
            rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
            agg = OrderedDict()
            for r in rows:
                agg.setdefault(r[4], []).append(float(r[-1]))
            total = sum(sum(v) for v in agg.values())
            out.append(f"# launch list {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)")
            for k, v in agg.items():
                out.append(f"{k:60s} launches={len(v):3d} avg_ms={sum(v) / len(v) / 1e6:10.3f} share={100 * sum(v) / total:5.1f}%")
        else:
            txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(txt.splitlines()))
            hdr, units = rows[0], rows[1]
            out.append(f"# full capture {path} (ncu --set full --clock-control none)")
            for r in rows[2:]:
                out.append(f"kernel: {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
                for w in WANT:
                    if w in hdr:
                        out.append(f"  {w:86s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
    text = "\n".join(out) + "\n"
    open(f"profiles/{tag}.txt", "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
