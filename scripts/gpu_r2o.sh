#!/bin/bash
# round 2: rows per CTA at 16384^2 (uniform segments; HG_FUSED_SEG), one GPU
mkdir -p gpurun_out
for s in 0 2731 1366 1093 911 683 456 342; do
echo "seg $s: $(HG_FUSED_SEG=$s timeout 300 python scripts/quick_bench.py --fused 16384 2>&1 | grep -o 'N=[0-9]* .*ms/step' | tr '\n' ';')"
done
