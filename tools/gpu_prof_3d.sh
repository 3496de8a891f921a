#!/bin/bash
mkdir -p gpurun_out
python tools/prof_3d.py > gpurun_out/p3d.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_3d.csv python tools/prof_3d.py > gpurun_out/ncu_p3d.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_3d.csv
