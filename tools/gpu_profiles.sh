#!/bin/bash
# profile artifacts of the current build: launch list of one step, ncu --set full of the two conv variants and of the SENSE kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launches rc=$?"
tools/gpu_prof_conv.sh
POINT="4 256 64 40" tools/gpu_prof_sense.sh
