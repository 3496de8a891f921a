#!/bin/bash
# ncu evidence for one round: launch list of one ALD step, then a full-set capture of the top kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --chains ${CHAINS:-14}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv_igemm -s 2 -c 2 -o gpurun_out/prof_igemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
ls -la gpurun_out
