#!/bin/bash
# BASELINE configs 3 and 5 on N GPUs of one box (manual runs; bench.py defaults are config[1]).
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
# config 3: 16384^2 strong-scaled over N slabs
timeout 900 $( [ $N = 1 ] && echo "python bench.py" || echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py" ) --gpus $N --width 16384 --rows-per-gpu $((16384 / N)) --steps 100 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/cfg3_$N.log 2>&1
grep -o '"value": [0-9.]*, "unit": "Gcell-steps/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*\|"halo_errors": [0-9]*' gpurun_out/cfg3_$N.log | tr '\n' ' '; echo
if [ $N = 8 ]; then
  # config 5: 65536^2 over 8 slabs of 8192 rows (38.7 GB of planes per GPU)
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --width 65536 --rows-per-gpu 8192 --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/cfg5_8.log 2>&1
  grep -o '"value": [0-9.]*, "unit": "Gcell-steps/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*\|"halo_errors": [0-9]*' gpurun_out/cfg5_8.log | tr '\n' ' '; echo
  tail -3 gpurun_out/cfg5_8.log | cut -c1-300
fi
