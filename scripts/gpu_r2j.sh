#!/bin/bash
# round 2, queued thermal outflow (k_fused_q): parity subset + device-timed bench per variant
mkdir -p gpurun_out
for v in 18 19 20 5; do
HG_FUSED_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py -m gpu -x -q 2>&1 | tail -3
HG_FUSED_VARIANT=$v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/q_v$v.log 2>&1
echo "variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/q_v$v.log | tr '\n' ' ')"
HG_FUSED_VARIANT=$v timeout 200 python scripts/exp_thermal.py 4096 2>&1 | tee gpurun_out/exp_thermal_v$v.log
done
