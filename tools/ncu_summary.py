"""Condense an .ncu-rep (ncu --set full) into the per-kernel CSV kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_ncu_full.csv"""
import csv, subprocess, sys
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
idx = [H.index(w) for w in WANT if w in H]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([H[i] for i in idx])
    w.writerow([rows[1][i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
print("wrote", sys.argv[2], len(rows) - 2, "launches")
