#!/bin/bash
# A/B of the SENSE kernels: current build vs the IPDM_PLAN_GW=4 variant (variants/libipdm_gw4.so)
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "sense" > $O/r2_t13.log 2>&1; tail -1 $O/r2_t13.log | cut -c1-250
python tools/bench_sense.py > $O/r2_sweep_j.jsonl 2>&1; grep '"batch": 64' $O/r2_sweep_j.jsonl | grep 'R": 40' | cut -c1-420
if [ -f variants/libipdm_gw4.so ]; then
IPDM_B200_LIB=/root/repo/variants/libipdm_gw4.so python -m pytest tests/test_gpu_parity.py -x -q -k "sense" 2>&1 | tail -1
IPDM_B200_LIB=/root/repo/variants/libipdm_gw4.so python tools/bench_sense.py > $O/r2_sweep_j_gw4.jsonl 2>&1; grep '"batch": 64' $O/r2_sweep_j_gw4.jsonl | grep 'R": 40' | cut -c1-420
fi
