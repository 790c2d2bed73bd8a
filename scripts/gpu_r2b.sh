#!/bin/bash
# round 2, visit B: the two-columns-per-thread kernel (variants 10..12) against the one-column kernel (5): parity + time + instruction counts
mkdir -p gpurun_out
for v in 10; do
HG_FUSED_VARIANT=$v timeout 900 python -m pytest tests -m gpu -x -q -k "grid or sizes or fuzz or golden or drift or slab" 2>&1 | tail -3
done
for v in 5 10 11 12; do
HG_FUSED_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ws_v$v.log 2>&1
echo "variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/ws_v$v.log | tr '\n' ' ')"
done
for v in 5 10; do
HG_FUSED_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 70 -c 1 -o gpurun_out/fused_v$v -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_v$v.log 2>&1
done
ls -la gpurun_out | tail -5
