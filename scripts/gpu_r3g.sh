#!/bin/bash
# ring offsets carried and rotated instead of recomputed each row: parity + same-box A/B (variants/lib_norot.so = recomputed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py tests/test_gpu_fuzz.py tests/test_gpu_particles_slabs.py -m gpu -x -q 2>&1 | tail -1
for rep in 1 2; do for lib in norot new; do
if [ $lib = norot ]; then export HG_B200_LIB=$PWD/variants/lib_norot.so; else unset HG_B200_LIB; fi
timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/g_$lib.log 2>&1
echo "$lib: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/g_$lib.log | tr '\n' ' ')"
done; done
