#!/bin/bash
# round 2: gradient table in the rain kernel: parity (rain tests + whole grid file) and time per launch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py tests/test_gpu_sizes.py -m gpu -x -q 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_rain -c 6 --csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras 2>/dev/null | grep k_rain | awk -F'","' '{print $5, $(NF)}' | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/w_v5.log 2>&1
echo "default: $(grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/w_v5.log | head -3 | tr '\n' ' ')"
