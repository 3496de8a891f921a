#!/bin/bash
# compute-sanitizer over the unit tests of the shared-memory FFT kernels (SENSE, pruned and general) and the TMEM / TMA
# convolution kernels (race / memory checks).  Run on a GPU box:  bash tools/sanitize.sh  -> gpurun_out/sanitize_*.log
# memcheck: out-of-bounds / misaligned global and shared accesses; racecheck: shared-memory hazards between warps of a CTA;
# synccheck: divergent barriers.  The kernels under test are small (the sanitizer runs them 10-100x slower); every tool
# has its own time limit (default 600 s) so that a slow selection cannot eat the box.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LIMIT=${SANITIZE_LIMIT:-600}
SEL_MEM='sense_pruned_plan_variants or sense_two_pass_engine_variants or fft_sizes_vs_oracle or conv_16bit_residual_stream or conv_halo_output_modes or chain_noise or instnorm_plus_isolated or sense_plan_falls_back'
SEL_RACE='sense_pruned_plan_variants or conv_16bit_residual_stream or chain_noise or sense_plan_falls_back'
SKIP='not 256-256-128-128 and not 20000'          # leave the largest shapes out: hours under the sanitizer
: > gpurun_out/sanitize_summary.txt
for tool in memcheck racecheck synccheck; do
  if [ $tool = memcheck ]; then SEL=$SEL_MEM; else SEL=$SEL_RACE; fi
  timeout $LIMIT compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "($SEL) and $SKIP" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$? (124 = time limit of $LIMIT s reached; 9 = sanitizer errors)" | tee -a gpurun_out/sanitize_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3 | tee -a gpurun_out/sanitize_summary.txt
done
