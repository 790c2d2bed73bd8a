#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -1
for rep in 1 2; do
timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/b_v5.log 2>&1
echo "default: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/b_v5.log | tr '\n' ' ')"
done
echo "quick: $(timeout 300 python scripts/quick_bench.py --fused 4096 16384 2>&1 | grep -o 'N=[0-9]* .*ms/step' | tr '\n' ';')"
