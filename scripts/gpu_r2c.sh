#!/bin/bash
# round 2, visit C: the two-columns-per-thread kernel after the FFMA2-fusion fix: parity of the whole grid suite + time
mkdir -p gpurun_out
HG_FUSED_VARIANT=10 timeout 1200 python -m pytest tests -m gpu -q -k "grid or sizes or fuzz or golden or drift or slab or checkpoint or host_driver" 2>&1 | tail -8 > gpurun_out/pytest_v10.log
cat gpurun_out/pytest_v10.log
for v in 5 10; do
HG_FUSED_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ws_v$v.log 2>&1
echo "variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/ws_v$v.log | tr '\n' ' ')"
done
HG_FUSED_VARIANT=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 70 -c 1 -o gpurun_out/fused_v10 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_v10.log 2>&1
