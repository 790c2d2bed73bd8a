#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (ours + reference arm), ncu launch list, one full ncu capture of
# the fused kernel.  /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh'; then
# python scripts/ncu_summary.py <tag> to turn gpurun_out/ into profiles/<tag>_*.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python scripts/quick_bench.py --fused 1024 4096 16384 > gpurun_out/quick.log 2>&1
bash scripts/gpu_profile.sh
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
ls -la gpurun_out
