#!/bin/bash
# bench line + ncu launch list + one full capture of the fused kernel (profiles/<tag>_*)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 70 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_rain -c 1 -o /tmp/rain_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_rain.log 2>&1
ncu -i /tmp/rain_full.ncu-rep --page raw --csv > gpurun_out/rain_raw.csv
tail -1 gpurun_out/bench.log | cut -c1-300
