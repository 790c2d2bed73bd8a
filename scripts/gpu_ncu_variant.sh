#!/bin/bash
mkdir -p gpurun_out
V=${1:-5}
HG_FUSED_VARIANT=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 70 -c 1 -o gpurun_out/fused_v$V -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_v$V.log 2>&1
tail -2 gpurun_out/ncu_v$V.log | cut -c1-300
