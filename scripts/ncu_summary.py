"""Summarise an .ncu-rep (read here, no GPU): one row per captured launch with the metrics the roofline needs."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_i%"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ]
cols = [(hdr.index(a), b) for a, b in want if a in hdr]
units = rows[1]
print(" | ".join(b + ("[" + units[i] + "]" if units[i] else "") for i, b in cols))
for r in rows[2:]:
    print(" | ".join((r[i][:48] if b == "kernel" else r[i]) for i, b in cols))
