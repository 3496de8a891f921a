#!/bin/bash
O=gpurun_out; mkdir -p $O
for d in 0 1 2 3; do
  IPDM_COLS_ONE=$d python tools/bench_sense.py > $O/r2_sweep_one$d.jsonl 2>&1; echo "one=$d"; grep '"batch": 64' $O/r2_sweep_one$d.jsonl | grep 'R": 40' | cut -c1-420
done
