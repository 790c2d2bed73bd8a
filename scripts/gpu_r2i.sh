#!/bin/bash
# round 2: what the thermal outflow path costs (marks forced off / on), and a full capture of the three-group kernel
mkdir -p gpurun_out
timeout 300 python scripts/exp_thermal.py 4096 2>&1 | tee gpurun_out/exp_thermal.log
HG_FUSED_VARIANT=15 timeout 300 python scripts/exp_thermal.py 4096 2>&1 | tee gpurun_out/exp_thermal_v15.log
bash scripts/gpu_ncu_variant.sh 15
