#!/bin/bash
mkdir -p gpurun_out
for m in res f16; do
python tools/prof_one.py $m > gpurun_out/prof_one_$m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo_$m python tools/prof_one.py $m > gpurun_out/ncu_halo_$m.log 2>&1
echo "$m rc=$?"
done
ls -la gpurun_out/*.ncu-rep
