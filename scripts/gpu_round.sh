#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, one full ncu capture of the fused kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
ldconfig -p | grep -E 'libEGL|libGLX_nvidia|libnvidia-egl|libGL' > gpurun_out/gl_probe.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python scripts/quick_bench.py 1024 4096 16384 > gpurun_out/quick.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 70 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
