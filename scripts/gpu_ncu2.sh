#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 128" "3 256"; do
  set -- $cfg
  HG_FUSED_VARIANT=$1 HG_FUSED_SEG=$2 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_v$1_s$2.log 2>&1
  HG_FUSED_VARIANT=$1 HG_FUSED_SEG=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 70 -c 1 -o gpurun_out/fused_v$1_s$2 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_v$1.log 2>&1
done
grep -h -o '"value": [0-9.]*, "unit": "Gcell-steps/s", "n_gpus"\|"kernel_ms": [0-9.]*' gpurun_out/bench_v*.log
