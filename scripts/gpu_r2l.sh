#!/bin/bash
# measurement only: how the step time follows the hydraulic group's work (stage B / erosion removed: WRONG results)
mkdir -p gpurun_out
for lib in nob noero; do for v in 5 18; do
HG_B200_LIB=$PWD/variants/lib_$lib.so HG_FUSED_VARIANT=$v timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/exp_${lib}_v$v.log 2>&1
echo "$lib variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/exp_${lib}_v$v.log | tr '\n' ' ')"
done; done
