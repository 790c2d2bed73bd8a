#!/bin/bash
# warp groups interleaved in pairs of warps (each scheduler runs one loop only): parity + time
mkdir -p gpurun_out
for v in 5 25 26 5; do
HG_FUSED_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_grid.py -m gpu -x -q 2>&1 | tail -1
HG_FUSED_VARIANT=$v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/a_v$v.log 2>&1
echo "variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/a_v$v.log | tr '\n' ' ')"
done
