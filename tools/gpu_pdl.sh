#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
export IPDM_CONV_PDL=1
run t_pdl python -m pytest tests/test_gpu_parity.py -q -x -k "conv or scorenet or ngf128 or ncsn3d or sampler or graph"
run bench_pdl python bench.py --steps 10 --warmup 3 --no-cpu-baseline
export IPDM_CONV_PDL=0
run bench_nopdl python bench.py --steps 10 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_pdl.log; tail -n 1 gpurun_out/bench_pdl.log | cut -c1-200; tail -n 1 gpurun_out/bench_nopdl.log | cut -c1-200
