#!/bin/bash
# round 2: 192-column strips and the queued kernel at 4096^2 and 16384^2
mkdir -p gpurun_out
HG_FUSED_VARIANT=24 timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_ref_golden.py -m gpu -x -q 2>&1 | tail -1
for v in 5 18 22 24; do
echo "variant $v: $(HG_FUSED_VARIANT=$v timeout 300 python scripts/quick_bench.py --fused 4096 16384 2>&1 | grep -o 'N=[0-9]* .*ms/step' | tr '\n' ';')"
done
