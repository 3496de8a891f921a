"""Per-kernel table from an .ncu-rep (ncu -i ... --page raw --csv): time, DRAM bytes, throughputs, occupancy, registers,
issue activity, top stall reasons.  usage: python tools/ncu_table.py file.ncu-rep [name-filter] [--json OUT IMAGES SIZE]
(--json: also write {kernel, images, size, dram_bytes_per_launch, ...} of the LAST matching launch -- bench.py reads
profiles/r02_ncu_conv_dominant.json for roofline.traffic)"""
import csv, io, json, subprocess, sys
argv = sys.argv[1:]
jout = None
if "--json" in argv:
    i = argv.index("--json")
    jout, j_images, j_size = argv[i + 1], int(argv[i + 2]), int(argv[i + 3])
    argv = argv[:i] + argv[i + 4:]
rep = argv[0]
flt = argv[1] if len(argv) > 1 else ""
last = None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default
def f(r, name):
    try:
        return float(g(r, name).replace(",", ""))
    except ValueError:
        return float("nan")
stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or
              (h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"))]
stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith("ratio") and "not_issued" not in h]
for r in data:
    name = g(r, "Kernel Name")
    if flt and flt not in name:
        continue
    t_unit = units[col["gpu__time_duration.sum"]]
    dur = f(r, "gpu__time_duration.sum")
    dur_us = dur / 1e3 if t_unit in ("nsecond", "ns") else dur * (1e3 if t_unit.startswith("ms") else 1)
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    ru, wu = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= sc.get(ru, 1); wr *= sc.get(wu, 1)
    last = {"kernel": name, "duration_us": dur_us, "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": rd + wr,
            "tensor_pipe_pct": f(r, "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active")
            if "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
            "dram_pct": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "source": rep.split("/")[-1],
            "note": "ncu --set full --clock-control none, one launch; cold-cache and serialised"}
    stalls = sorted(((f(r, h), h.split("issue_stalled_")[1].split("_per")[0].replace(".ratio", "")) for h in stall_cols), reverse=True)[:4]
    print(f"{name[:60]:60s} {dur_us:9.1f}us rd {rd/1e6:8.1f}MB wr {wr/1e6:8.1f}MB dram% {f(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"sm% {f(r,'sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} l1% {f(r,'l1tex__throughput.avg.pct_of_peak_sustained_active'):5.1f} "
          f"occ% {f(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} regs {g(r,'launch__registers_per_thread')} "
          f"issue% {f(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} inst {f(r,'smsp__inst_executed.sum'):.3g} "
          f"bankconf {f(r,'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):.3g} grid {g(r,'launch__grid_size')} | " +
          " ".join(f"{n}={v:.1f}" for v, n in stalls))
if jout and last:
    last.update(images=j_images, size=j_size)
    with open(jout, "w") as fh:
        json.dump(last, fh, indent=1)
