#!/bin/bash
# after the packed-arithmetic trims: the variants again, bench workload (4096^2) and quick_bench (4096^2, 16384^2)
mkdir -p gpurun_out
for v in 5 18 22; do
HG_FUSED_VARIANT=$v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/y_v$v.log 2>&1
echo "variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/y_v$v.log | tr '\n' ' ') quick: $(HG_FUSED_VARIANT=$v timeout 300 python scripts/quick_bench.py --fused 4096 16384 2>&1 | grep -o 'N=[0-9]* .*ms/step' | tr '\n' ';')"
done
