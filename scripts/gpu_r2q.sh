#!/bin/bash
# round 2, after the one-division atan and the sqrt-free momentum cut-off: whole GPU suite, kernel time, droplet dispatch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/q2_v5.log 2>&1
echo "default: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/q2_v5.log | tr '\n' ' ')"
timeout 300 python scripts/particle_bench.py 100 2>&1 | tee gpurun_out/q2_drops.log
