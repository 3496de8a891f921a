"""Per-kernel breakdown of one ALD step from an ncu launch list (gpurun_out/launches.csv)."""
import csv, re, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines=[l for l in open(path) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
idx=[i for i,r in enumerate(rows) if 'k_conv_first' in r['Kernel Name']]
# the LAST complete step of the run: the first forwards carry one-off work (buffer fills, the range audit, graph priming)
seg=rows[idx[-2]:idx[-1]] if len(idx) > 1 else rows[idx[0]:]
agg=collections.defaultdict(lambda:[0,0.0])
for row in seg:
    name=re.sub(r'\(.*','',row['Kernel Name']).replace('ipdm::','').replace('void ','')
    t=float(row['Metric Value'].replace(',',''))
    if row['Metric Unit']=='ns': t/=1e3
    agg[name][0]+=1; agg[name][1]+=t
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    if v[1]/tot < 0.001: continue
    print(f"{k[:50]:50s} n={v[0]:4d} total={v[1]/1e3:8.3f} ms share={v[1]/tot*100:5.1f}%")
print(f"one step (serialised, cold cache): {tot/1e3:.3f} ms, {len(seg)} launches")
