#!/bin/bash
# VARIANTS="0 5 8" bash scripts/gpu_variants_env.sh: GPU tests + device-timed bench per HG_FUSED_VARIANT (same box: A/B)
mkdir -p gpurun_out
for v in $VARIANTS; do
HG_FUSED_VARIANT=$v timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
HG_FUSED_VARIANT=$v timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ws_v$v.log 2>&1
echo "variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/ws_v$v.log | tr '\n' ' ')"
done
