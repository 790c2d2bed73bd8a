#!/bin/bash
# compute-sanitizer over a small wet run of the product path (init, rain, fused step, far fix-up, pack/unpack)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from hydro_gen_b200 import Context
ctx = Context(128)
m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
r = ctx.get_rain(); r.period = 2; ctx.set_rain(r)
ctx.gen_heightmap()
ctx.run(6, 0.015, 0.015, True)
print("far", ctx.far_fetch_count(), "sum", float(ctx.download(0).sum()))
ctx.close()
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|far ' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
