#!/bin/bash
# ncu --set full of every SENSE kernel at one sweep point (second repetition = warm)
mkdir -p gpurun_out
ARGS="${POINT:-4 256 64 40}"
python tools/prof_sense.py $ARGS > gpurun_out/ps.log 2>&1 || { cat gpurun_out/ps.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k2_|k_fwd|k_adj|k_ald' --launch-skip 7 -c 7 -f -o gpurun_out/prof_sense python tools/prof_sense.py $ARGS > gpurun_out/ncu_ps.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_ps.log
