#!/bin/bash
# usage: gpu_bench_variants.sh "v seg" "v seg" ...   (bench state, kernel_ms per variant; ncu of the first)
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  HG_FUSED_VARIANT=$1 HG_FUSED_SEG=$2 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_v$1_s$2.log 2>&1
  echo "variant $1 seg $2: $(grep -h -o '"value": [0-9.]*, "unit": "Gcell-steps/s", "n_gpus"\|"kernel_ms": [0-9.]*\|"ms_per_step": [0-9.]*, "higher' gpurun_out/bench_v$1_s$2.log | tr '\n' ' ')"
done
