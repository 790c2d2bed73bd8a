#!/bin/bash
mkdir -p gpurun_out
cp hydro_gen_b200/libhydrogen_b200.so /tmp/lib_base.so
for so in /tmp/lib_base.so variants/lib_*.so; do
  n=$(basename $so .so)
  cp $so hydro_gen_b200/libhydrogen_b200.so
  for v in $VARIANTS; do
    HG_FUSED_VARIANT=$v timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 1 > gpurun_out/var_${n}_v$v.log 2>&1
    echo "$n variant $v: $(grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*' gpurun_out/var_${n}_v$v.log | tr '\n' ' ')"
  done
done
cp /tmp/lib_base.so hydro_gen_b200/libhydrogen_b200.so
