#!/bin/bash
# One GPU-box session: run the GPU suite in isolated groups (a trapped kernel must not poison the
# other groups), then smoke, then a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run t_fft_sense python -m pytest tests/test_gpu_parity.py -q -k "library or fft or sense_and_prox or sense_full or fused_step or posterior" 
run t_conv python -m pytest tests/test_gpu_parity.py -q -k "conv"
run t_small python -m pytest tests/test_gpu_parity.py -q -k "scorenet_small or sampler_uncond or sampler_sense or sampler_cine or philox or graph_path"
run t_ngf128 python -m pytest tests/test_gpu_parity.py -q -k "ngf128"
run smoke python -c "import __graft_entry__ as g; g.smoke()"
if [ -f bench.py ]; then run bench python bench.py --steps 3 --warmup 3; fi
cat gpurun_out/summary.txt
tail -n 25 gpurun_out/t_fft_sense.log gpurun_out/t_conv.log gpurun_out/t_small.log gpurun_out/t_ngf128.log gpurun_out/smoke.log
