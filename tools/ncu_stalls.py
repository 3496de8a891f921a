"""Stall-reason totals + hottest SASS lines of ONE kernel in an .ncu-rep:  python tools/ncu_stalls.py rep launch_index [ntop]"""
import csv, subprocess, sys, io, collections
rep, skip = sys.argv[1], sys.argv[2]; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 14
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
print(" ".join(rows[0])[:120] if hi > 0 else "")
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
iS = h.index("# Samples"); isrc = h.index("Source")
stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[iS]) for r in data if r[iS].isdigit())
agg = collections.Counter()
for r in data:
    for i in stall:
        if r[i].isdigit(): agg[h[i][6:]] += int(r[i])
print("total samples", tot, " instructions", len(data))
print("stall totals:", ", ".join(f"{k}={100*v/max(1,sum(agg.values())):.1f}%" for k, v in agg.most_common(9)))
ops = collections.Counter()
iE = h.index("Instructions Executed") if "Instructions Executed" in h else None
for r in data:
    op = r[isrc].strip().split()
    if not op: continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    if iE is not None and r[iE].isdigit(): ops[o.split(".")[0]] += int(r[iE])
print("warp-insts by opcode:", ", ".join(f"{k}={v}" for k, v in ops.most_common(14)), " total", sum(ops.values()))
for r in sorted(data, key=lambda r: -int(r[iS]) if r[iS].isdigit() else 0)[:ntop]:
    st = {h[i][6:]: int(r[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*int(r[iS])/max(1,tot):5.1f}%  {r[isrc].strip()[:64]:64s} {st}")
