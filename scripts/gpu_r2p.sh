#!/bin/bash
# round 2: fused halo push A/B on N GPUs: cross-process slab parity, then the weak-scaling headline with and without it
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check_push_$N.log 2>&1; echo "exit $?" >> gpurun_out/mgpu_check_push_$N.log
grep -v "^W\|^\*\*\*" gpurun_out/mgpu_check_push_$N.log | tail -4
for p in 1 0 1 0; do
HG_FUSED_PUSH=$p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 20 --e2e-steps 0 --no-extras --no-cpu-baseline > gpurun_out/push${p}_$N.log 2>&1
echo "fused push $p, $N GPUs: $(grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*\|"halo_errors": [0-9]*' gpurun_out/push${p}_$N.log | head -4 | tr '\n' ' ')"
done
