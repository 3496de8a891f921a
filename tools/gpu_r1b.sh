#!/bin/bash
# HEAD verification: GPU suite, smoke, default bench (with cpu_baseline), reference arm, SENSE sweep
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_all python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
run sense_sweep python tools/bench_sense.py
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_all.log; tail -n 2 gpurun_out/smoke.log; tail -n 1 gpurun_out/bench.log | cut -c1-300; tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
