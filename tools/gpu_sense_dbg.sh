#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "sense" > $O/r2_t14.log 2>&1; tail -1 $O/r2_t14.log | cut -c1-250
for d in 0 1 2 3 0; do
  IPDM_SENSE_DBG=$d python tools/bench_sense.py > $O/r2_sweep_dbg$d.jsonl 2>&1; echo "dbg=$d"; grep '"batch": 64' $O/r2_sweep_dbg$d.jsonl | grep 'R": 40' | cut -c1-420
done
