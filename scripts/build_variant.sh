#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>: builds gpurun_variants/lib_<name>.so from a scratch copy of csrc
set -e
name=$1; shift
root=$(cd $(dirname $0)/.. && pwd)
tmp=/tmp/hgvar_$name; rm -rf $tmp; mkdir -p $tmp/hydro_gen_b200 $root/variants
cp -r $root/hydro_gen_b200/csrc $tmp/hydro_gen_b200/; cp -r $root/include $tmp/
rm -f $tmp/hydro_gen_b200/csrc/*.o
make -s -C $tmp/hydro_gen_b200/csrc -j8 EXTRA="$*" > $tmp/build.log 2>&1 || (grep -B2 -A6 "error" $tmp/build.log | head -40; exit 1)
cp $tmp/hydro_gen_b200/libhydrogen_b200.so $root/variants/lib_$name.so
grep -A3 "k_fused_stepILi128ELi4" $tmp/hydro_gen_b200/csrc/hg_fused.ptxas.log | grep -E "Used|spill" 
