#!/bin/bash
# steady-state loop of ONE warp group unrolled (HG_UNROLL_H / HG_UNROLL_T): parity subset + time, same box
mkdir -p gpurun_out
for lib in base uh2 ut2 uht2 uh3 base; do
if [ $lib = base ]; then unset HG_B200_LIB; else export HG_B200_LIB=$PWD/variants/lib_$lib.so; fi
timeout 300 python -m pytest tests/test_gpu_grid.py -m gpu -x -q 2>&1 | tail -1
timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/z_$lib.log 2>&1
echo "$lib: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/z_$lib.log | tr '\n' ' ')"
done
