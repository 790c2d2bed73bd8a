#!/bin/bash
# same box A/B: the library of commit 66388d1 (end of the first session) against the current one, at the driver's flags and at 300 steps
mkdir -p gpurun_out
for rep in 1 2; do for lib in old new; do for fl in "20 5" "300 20"; do set -- $fl
if [ $lib = old ]; then export HG_B200_LIB=$PWD/variants/lib_r02g.so; else unset HG_B200_LIB; fi
timeout 300 python bench.py --steps $1 --warmup $2 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/ab_${lib}_$1.log 2>&1
echo "$lib steps $1 warmup $2: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/ab_${lib}_$1.log | tr '\n' ' ')"
done; done; done
