#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python scripts/quick_bench.py --fused 1024 4096 8192 16384 2>&1 | tee gpurun_out/quick.log
HG_FUSED_VARIANT=0 timeout 600 python scripts/quick_bench.py --fused 1024 4096 16384 2>&1 | tee gpurun_out/quick_v0.log
