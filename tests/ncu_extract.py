"""Reduce an .ncu-rep to a small text table of the metrics the roofline discussion needs (run on the GPU box so that
only the summary travels back):  python tests/ncu_extract.py in.ncu-rep out.txt"""
import csv, subprocess, sys
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    for k, row in enumerate(rows[2:]):
        f.write(f"--- launch {k}\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"  {w:90s} {row[i]} {units[i]}\n")
