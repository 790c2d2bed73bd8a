#!/bin/bash
# same-box A/B of the K = 1 shortcut (variants/lib_nok.so = without it)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py tests/test_gpu_fuzz.py tests/test_gpu_sizes.py -m gpu -x -q 2>&1 | tail -1
for rep in 1 2; do for lib in nok new; do
if [ $lib = nok ]; then export HG_B200_LIB=$PWD/variants/lib_nok.so; else unset HG_B200_LIB; fi
timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/c_$lib.log 2>&1
echo "$lib: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/c_$lib.log | tr '\n' ' ') quick: $(timeout 300 python scripts/quick_bench.py --fused 4096 16384 2>&1 | grep -o 'N=[0-9]* .*ms/step' | tr '\n' ';')"
done; done
