#!/bin/bash
# GPU visit: sanity (tests, bench), full ncu capture with source counters, pipe micro-benchmark.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 70 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
nvcc -arch=sm_100a -O3 -o /tmp/pipe_bench scripts/scratch/pipe_bench.cu && timeout 120 /tmp/pipe_bench > gpurun_out/pipe_bench.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench.log; head -14 gpurun_out/pipe_bench.txt
