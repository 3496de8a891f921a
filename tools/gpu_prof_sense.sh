#!/bin/bash
# ncu --set full of the plan (pruned) SENSE kernels at two sweep points, second repetition
O=gpurun_out; mkdir -p $O
for P in "32 512 64 40:big" "4 256 64 40:small"; do
  ARGS=${P%%:*}; TAG=${P##*:}
  python tools/prof_sense.py $ARGS > $O/ps.log 2>&1 || { cat $O/ps.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:'kp_' --launch-skip 5 -c 5 -f -o $O/r2_sense_${TAG}_k python tools/prof_sense.py $ARGS > $O/ncu_ps_$TAG.log 2>&1
  echo "$TAG ncu rc=$?"
done
