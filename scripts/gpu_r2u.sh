#!/bin/bash
# fused halo push A/B on N GPUs (timing only)
N=${1:-4}
mkdir -p gpurun_out
for p in 1 0 1 0; do
HG_FUSED_PUSH=$p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 20 --e2e-steps 0 --no-extras --no-cpu-baseline > gpurun_out/push${p}_$N.log 2>&1
echo "fused push $p, $N GPUs: $(grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/push${p}_$N.log | head -3 | tr '\n' ' ')"
done
