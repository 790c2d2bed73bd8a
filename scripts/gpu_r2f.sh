#!/bin/bash
# round 2, visit F: ncu launch list + full capture of the fused kernel for the r02 profiles; compute-sanitizer over the code that is new this round
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_ws -s 70 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_full.log 2>&1
SEL="droplet_slabs or slab_checkpoint or particle_erode_sparse or particle_tail or pack_device or slabs_on_one_gpu or fused_one_step or host_async"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_r02_$tool.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitize_r02_$tool.log
  tail -3 gpurun_out/sanitize_r02_$tool.log
done
