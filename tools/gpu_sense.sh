#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_sense python -m pytest tests/test_gpu_parity.py -q -x -k "fft or sense or fused_step or sampler or philox"
run sense_sweep python tools/bench_sense.py ${SWEEP:-}
if [ -n "$PROF" ]; then tools/gpu_prof_sense.sh > gpurun_out/prof_sense.log 2>&1; fi
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/t_sense.log; cut -c1-330 gpurun_out/sense_sweep.log
