#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_3d python -m pytest tests/test_gpu_parity.py -q -x -k "conv or ncsn3d or cine or scorenet or ngf128"
run configs python tools/bench_configs.py
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_3d.log; cut -c1-420 gpurun_out/configs.log
