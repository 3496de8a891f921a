#!/bin/bash
# Round-2 validation on one B200: GPU tests, smoke, every bench configuration / arm, launch list, ncu captures.
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_tests_gpu.log 2>&1; echo "tests rc=$? $(tail -1 $O/r2_tests_gpu.log)"
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 $O/r2_bench_1gpu.json
python bench.py --impl reference > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/r2_bench_reference.json
python bench.py --impl reference-gpu > $O/r2_bench_reference_gpu.json 2> $O/r2_bench_reference_gpu.err; echo "refgpu rc=$?"; cut -c1-300 $O/r2_bench_reference_gpu.json
for c in cfg1 cfg4-none cfg4-tv cfg4-diffusion cfg2-B1 cfg5-sweep; do
  python bench.py --config $c > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err; echo "$c rc=$?"; tail -1 $O/r2_bench_$c.json | cut -c1-300
done
python bench.py --impl reference --config cfg5-sweep --steps 3 --warmup 1 > $O/r2_bench_cfg5-sweep_reference.json 2>&1
python tools/bench_sense.py > $O/r2_sense_sweep.jsonl 2>&1; echo "sweep rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/r2_launches_one_step.csv $CMD > $O/r2_ncu_launches.log 2>&1; echo "launches rc=$?"
python tools/prof_one.py t16 > $O/prof_one_t16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o $O/r2_conv_t16 python tools/prof_one.py t16 > $O/ncu_t16.log 2>&1; echo "ncu t16 rc=$?"
# summaries are made HERE (the reports themselves are too big to travel back: 64 MiB limit on gpurun_out)
python tools/ncu_summary.py $O/r2_conv_t16.ncu-rep > $O/r02_ncu_conv_halo_t16_mode.txt 2>&1
python tools/ncu_table.py $O/r2_conv_t16.ncu-rep "" --json $O/r02_ncu_conv_dominant.json 28 256 > /dev/null 2>&1
rm -f $O/r2_conv_t16.ncu-rep
bash tools/gpu_prof_sense.sh
for t in big small; do
  { python tools/ncu_table.py $O/r2_sense_${t}_k.ncu-rep; for i in 0 1 2 3 4; do echo; python tools/ncu_stalls.py $O/r2_sense_${t}_k.ncu-rep $i 12; done; } > $O/r02_ncu_sense_$t.txt 2>&1
  rm -f $O/r2_sense_${t}_k.ncu-rep
done
rm -f $O/*.ncu-rep
du -sh $O
