#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_conv python -m pytest tests/test_gpu_parity.py -q -x -k "conv or scorenet or ngf128"
run igemm_on python tools/bench_igemm.py 28 1
run igemm_off python tools/bench_igemm.py 28 0
run ends python tools/bench_ends.py
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_conv.log; grep res+ gpurun_out/igemm_on.log; echo; grep res+ gpurun_out/igemm_off.log; cat gpurun_out/ends.log; tail -n 1 gpurun_out/bench.log | cut -c1-300
