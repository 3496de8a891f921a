#!/bin/bash
# full GPU suite + bench; optional launch list (LAUNCHES=1)
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_all python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:---no-cpu-baseline}
if [ -n "$LAUNCHES" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "launches rc=$?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt; tail -n 12 gpurun_out/t_all.log; tail -n 2 gpurun_out/smoke.log; tail -n 1 gpurun_out/bench.log | cut -c1-300
