#!/bin/bash
# A/B of library variants under variants/*.so against the shipped build (forward / adjoint / step at the two judged points)
O=gpurun_out; mkdir -p $O
show() { grep '"batch": 64' $1 | grep 'R": 40' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['coils'],d['size'],'fwd',d['fwd_ms'],d['fwd_frac'],'adjm',d['adj_masked_ms'],'step',d['step_ms'])"; }
python tools/bench_sense.py > $O/ab_base.jsonl 2>&1; echo base; show $O/ab_base.jsonl
for f in variants/*.so; do
  IPDM_B200_LIB=/root/repo/$f python tools/bench_sense.py > $O/ab_$(basename $f .so).jsonl 2>&1; echo $f; show $O/ab_$(basename $f .so).jsonl
done
python tools/bench_sense.py > $O/ab_base2.jsonl 2>&1; echo base again; show $O/ab_base2.jsonl
