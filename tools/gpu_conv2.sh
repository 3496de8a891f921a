#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_conv python -m pytest tests/test_gpu_parity.py -q -x -k "conv or scorenet or ngf128 or ncsn3d"
run igemm python tools/bench_igemm.py 28
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_conv.log; head -15 gpurun_out/igemm.log; tail -n 1 gpurun_out/bench.log | cut -c1-250
