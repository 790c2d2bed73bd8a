#!/bin/bash
mkdir -p gpurun_out
for seg in 0 342 205 171 147 128 114 93 86; do
  HG_FUSED_SEG=$seg timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 1 > gpurun_out/seg_$seg.log 2>&1
  echo "seg $seg: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/seg_$seg.log | tr '\n' ' ')"
done
