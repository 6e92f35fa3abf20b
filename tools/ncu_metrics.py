"""Prints the headline metrics of every kernel in an .ncu-rep (from `ncu -i X --page raw --csv`)."""
import csv, subprocess, sys
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[0]; units = rows[1]
for r in rows[2:]:
    print("---")
    for w in WANT:
        if w in H:
            i = H.index(w); print(f"{w:85s} {r[i]} {units[i]}")
    for w in H:
        if "issue_stalled" in w and "per_warp_active" in w and "not_issued" not in w:
            v = float(r[H.index(w)])
            if v > 3.0: print(f"  stall {w.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):30s} {v:.1f} %")
