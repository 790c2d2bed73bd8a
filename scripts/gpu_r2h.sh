#!/bin/bash
# round 2, three-warp-group kernel (k_fused_ws3): parity subset + device-timed bench per variant, droplet tail variant
mkdir -p gpurun_out
for v in 5 13 14 15 16 17; do
HG_FUSED_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py -m gpu -x -q 2>&1 | tail -1
HG_FUSED_VARIANT=$v timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ws3_v$v.log 2>&1
echo "variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/ws3_v$v.log | tr '\n' ' ')"
done
for d in 0 2; do
HG_DROPS_VARIANT=$d timeout 600 python -m pytest tests/test_gpu_particles_slabs.py -m gpu -x -q -k "not slab" 2>&1 | tail -1
HG_DROPS_VARIANT=$d timeout 300 python scripts/particle_bench.py 100 2>&1 | tee gpurun_out/ws3_drops_$d.log
done
