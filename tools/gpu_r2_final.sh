#!/bin/bash
# Reduced end-of-round check: GPU tests, A/B of the dense halo kernel against the previous build (if variants/lib_prevhalo.so is
# there), default bench, the small configs, the launch list of one step.
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2_tests_gpu.log 2>&1; echo "tests rc=$? $(tail -1 $O/r2_tests_gpu.log)"
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
show() { head -6 | python -c "
import sys,json
print([(json.loads(l)['mode'][:14], json.loads(l)['ms']) for l in sys.stdin])"; }
echo current; python tools/bench_igemm.py 28 2>&1 | show
if [ -f variants/lib_prevhalo.so ]; then echo previous-halo; IPDM_B200_LIB=/root/repo/variants/lib_prevhalo.so python tools/bench_igemm.py 28 2>&1 | show; fi
python tools/bench_pooled_conv.py 2>&1 | head -3
python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; echo "bench rc=$?"; cut -c1-200 $O/r2_bench_1gpu.json
for c in cfg1 cfg2-B1 cfg4-none cfg4-tv cfg4-diffusion; do
  python bench.py --config $c > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err; echo "$c rc=$?"; tail -1 $O/r2_bench_$c.json | cut -c1-160
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/r2_launches_one_step.csv $CMD > $O/r2_ncu_launches.log 2>&1; echo "launches rc=$?"
python tools/launch_breakdown.py $O/r2_launches_one_step.csv | tail -1
