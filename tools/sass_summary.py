"""Per-kernel SASS opcode summary of libipdm_b200.so (cuobjdump -sass): the tcgen05 / TMEM / TMA mnemonics that prove the
tensor-core path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit), plus
instruction counts.  usage: python tools/sass_summary.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "inverseproblemwithdiffusionmodel_b200", "libipdm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMAPF", "UTCBAR", "SYNCS", "LDGSTS", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"{'kernel':90s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
for (mangled, c), name in zip(per.items(), names):
    name = re.sub(r"^void ", "", name).replace("ipdm::", "")
    name = re.sub(r"\(.*", "", name)
    print(f"{name[:90]:90s} {c['_total']:7d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    tot.update(c)
print(f"{'TOTAL (' + str(len(per)) + ' kernels)':90s} {tot['_total']:7d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
