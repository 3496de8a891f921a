#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "configs rc=$?"; cat gpurun_out/configs.log | cut -c1-500
