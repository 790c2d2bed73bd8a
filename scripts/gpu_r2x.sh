#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py tests/test_gpu_sizes.py tests/test_gpu_particles_slabs.py -m gpu -x -q 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none -k regex:k_rain -s 2 -c 1 -o /tmp/rain_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/ncu_rain.log 2>&1
ncu -i /tmp/rain_full.ncu-rep --page raw --csv > gpurun_out/rain_raw.csv
