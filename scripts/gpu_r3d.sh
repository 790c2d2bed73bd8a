#!/bin/bash
# same-box A/B of the K = 1 shortcut on the bench workload at 16384^2 (the north-star size)
mkdir -p gpurun_out
for rep in 1 2; do for lib in nok new; do
if [ $lib = nok ]; then export HG_B200_LIB=$PWD/variants/lib_nok.so; else unset HG_B200_LIB; fi
timeout 400 python bench.py --width 16384 --rows-per-gpu 16384 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/d_$lib.log 2>&1
echo "$lib 16384^2: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/d_$lib.log | tr '\n' ' ')"
done; done
