#!/bin/bash
mkdir -p gpurun_out
for h in 8192 1024; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/drops_ncu_$h.csv python scripts/drops_profile.py $h 10 > /dev/null 2>&1
done
