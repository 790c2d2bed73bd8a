#!/bin/bash
# droplet mode: how often the processing order is rebuilt (HG_DROPS_REBIN), 4 Mi droplets on 8192^2
for r in 2 4 8 16 32; do
echo "rebin $r: $(HG_DROPS_REBIN=$r timeout 300 python scripts/particle_bench.py 100 2>&1 | grep -o '4[ab] [a-z]* (hmap [0-9]*): [0-9.]* ms/step' | tr '\n' ';')"
done
