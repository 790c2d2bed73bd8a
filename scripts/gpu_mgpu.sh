#!/bin/bash
# N-GPU visit: slab parity against the whole-map run, then the weak-scaling bench line.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "exit $?" >> gpurun_out/mgpu_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 20 --e2e-steps 4 > gpurun_out/bench_$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_$N.log
grep -v "^W\|^\*\*\*" gpurun_out/mgpu_check_$N.log | tail -6; tail -3 gpurun_out/bench_$N.log | cut -c1-1500
