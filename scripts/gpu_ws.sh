#!/bin/bash
# warp-specialised variants: parity tests, then timing per variant / segment height
mkdir -p gpurun_out
HG_FUSED_VARIANT=5 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ws.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ws.log
tail -3 gpurun_out/pytest_ws.log
for cfg in "0 0" "5 0" "5 256" "5 342" "5 512" "6 0"; do
  set -- $cfg
  HG_FUSED_VARIANT=$1 HG_FUSED_SEG=$2 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ws_v$1_s$2.log 2>&1
  echo "variant $1 seg $2: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/ws_v$1_s$2.log | tr '\n' ' ')"
done
