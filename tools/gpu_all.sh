#!/bin/bash
# full GPU suite, smoke, default bench (+ reference arm when REF=1), SENSE sweep
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_all python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py ${BENCH_ARGS:-}
if [ -n "$REF" ]; then run bench_ref python bench.py --impl reference --steps 2 --warmup 1; fi
run sense_sweep python tools/bench_sense.py
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_all.log; tail -n 2 gpurun_out/smoke.log; tail -n 1 gpurun_out/bench.log | cut -c1-300
