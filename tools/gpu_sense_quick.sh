#!/bin/bash
# SENSE unit tests + the two judged sweep points (forward / masked adjoint / fused step)
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "sense or chain_noise or ald or sampler or no_kernel_writes" 2>&1 | tail -1
python tools/bench_sense.py > $O/r2_sweep_quick.jsonl 2>&1
grep '"batch": 64' $O/r2_sweep_quick.jsonl | grep 'R": 40' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['coils'],d['size'],'fwd',d['fwd_ms'],d['fwd_frac'],'adjm',d['adj_masked_ms'],'step',d['step_ms'],d['step_frac'],'step(injected noise)',d.get('step_injected_noise_ms'),d.get('step_injected_noise_frac'))"
