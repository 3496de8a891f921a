"""Key metrics + top stall lines of an .ncu-rep (first kernel):  python tools/ncu_summary.py file.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]
for k in want:
    if k in hdr:
        i = hdr.index(k); print(f"{k:75s} {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
iS = h.index("# Samples"); isrc = h.index("Source")
stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[iS]) for r in data if r[iS].isdigit())
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[iS]) if r[iS].isdigit() else 0)[:ntop]:
    st = {h[i][6:]: int(r[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*int(r[iS])/tot:5.1f}%  {r[isrc].strip()[:64]:64s} {st}")
