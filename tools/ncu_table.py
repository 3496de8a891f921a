"""Per-kernel key metrics of an .ncu-rep (all kernels):  python tools/ncu_table.py file.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.avg.per_cycle_active", "ipc/smsp"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("launch__grid_size", "grid"), ("launch__waves_per_multiprocessor", "waves"),
        ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long"), ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "st_short"),
        ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "st_mio"), ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg"),
        ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_bar"), ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "st_wait"),
        ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "st_math"), ("smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "st_notsel")]
idx = [(hdr.index(k), n) for k, n in cols if k in hdr]
for r in rows[2:]:
    print("  ".join(f"{n}={r[i][:60]}{rows[1][i] if n in ('rd','wr','us') else ''}" for i, n in idx))
