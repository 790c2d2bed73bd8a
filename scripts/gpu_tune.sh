#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/tune_fused.py 4096 16384 > gpurun_out/tune.log 2>&1
cat gpurun_out/tune.log
