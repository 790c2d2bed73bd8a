#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/quick_bench.py --fused --dry 4096 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 1 -o gpurun_out/fused_dry -f python scripts/quick_bench.py --fused --dry 4096 > gpurun_out/ncu_dry.log 2>&1
tail -1 gpurun_out/ncu_dry.log
