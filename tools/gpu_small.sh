#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_small python -m pytest tests/test_gpu_parity.py -q -x -k "maxpool or scorenet or ngf128 or bilinear or conv_direct"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/t_small.log; python tools/launch_breakdown.py gpurun_out/launches.csv | grep -E "maxpool|bilinear|conv_last|conv_first|act_to|instnorm|one step"
