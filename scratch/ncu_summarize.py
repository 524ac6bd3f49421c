"""Compact per-kernel summary of `ncu -i x.ncu-rep --page raw --csv` files (kept under profiles/).

usage: python scratch/ncu_summarize.py gpurun_out/TAG_*_raw.csv > profiles/rNN_TAG_ncu_summary.txt
"""
import csv
import re
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe (DFMA/DADD/DMUL) active %"),
    ("sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "DMMA tensor pipe active %"),
    ("sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe cycles active %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor pipe inst %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-instr"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-instr"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-instr"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
]


def main(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        rows = [r for r in rows if len(r) > 20]
        if len(rows) < 3:
            print("# %s: no kernels" % path)
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print("# %s" % path)
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            name = re.sub(r"\(.*", "", name)
            print("kernel %s  grid %s block %s" % (name, r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
            for key, label in KEEP:
                if key in col and r[col[key]] != "":
                    print("  %-44s %s %s" % (label, r[col[key]], units[col[key]]))
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
