#!/bin/bash
# round 2, final 1-GPU visit of the second session: GPU suite, smoke, bench line (+ extras), reference arm, launch list, full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 70 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench.log | cut -c1-400
