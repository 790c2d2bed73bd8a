#!/bin/bash
# One `ncu --set full` launch of every kernel in steady state (scripts/all_kernels.py brackets them with
# cudaProfilerStart/Stop).  The raw page is exported on the box (the report itself can exceed what gpurun copies
# back).  Back home: python scripts/ncu_kernels_summary.py <tag>
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --profile-from-start off -c 120 \
    -o /tmp/all_kernels -f python scripts/all_kernels.py > gpurun_out/all_kernels.log 2>&1
tail -3 gpurun_out/all_kernels.log
ncu -i /tmp/all_kernels.ncu-rep --page raw --csv > gpurun_out/all_kernels_raw.csv
ls -la /tmp/all_kernels.ncu-rep gpurun_out/all_kernels_raw.csv
[ $(stat -c %s /tmp/all_kernels.ncu-rep) -lt 40000000 ] && cp /tmp/all_kernels.ncu-rep gpurun_out/
du -sh gpurun_out
