#!/bin/bash
# Multi-GPU lines of round 2.  usage: bash tools/gpu_r2_multi.sh N [full]   (full: also --strong and cfg 3 at the full schedule)
N=${1:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err; echo "bench N=$N rc=$?"; tail -1 $O/r2_bench_${N}gpu.json | cut -c1-330
$TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/r2_bench_${N}gpu_reference.json 2> $O/r2_bench_${N}gpu_reference.err; echo "ref N=$N rc=$?"; tail -1 $O/r2_bench_${N}gpu_reference.json | cut -c1-200
$TR bench.py --strong --gpus $N --steps 10 --warmup 3 > $O/r2_bench_${N}gpu_strong.json 2> $O/r2_bench_${N}gpu_strong.err; echo "strong N=$N rc=$?"; tail -1 $O/r2_bench_${N}gpu_strong.json | cut -c1-330
if [ "${2:-}" = full ]; then
  $TR tools/run_posterior.py --chains 105 --out $O/posterior_full > $O/r2_posterior_105chains_${N}gpu_full.json 2> $O/r2_posterior_full.err; echo "posterior rc=$?"; tail -1 $O/r2_posterior_105chains_${N}gpu_full.json | cut -c1-600
  rm -rf $O/posterior_full/*.pt
fi
