#!/bin/bash
# compute-sanitizer over the second session's new code paths: the queued-outflow kernel (shared-memory queue, service warp),
# the three-group kernel, the fused halo push on three connected slabs of one process, the rain kernel's tables
mkdir -p gpurun_out
cat > /tmp/san_case2.py <<'PY'
import sys, os; sys.path.insert(0, '.')
import numpy as np
from hydro_gen_b200 import Context
def run_one(W, H):
    ctx = Context(W, H)
    m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
    r = ctx.get_rain(); r.period = 2; ctx.set_rain(r)
    ctx.gen_heightmap()
    ctx.run(6, 0.015, 0.015, True)
    print("variant", os.environ.get("HG_FUSED_VARIANT"), "far", ctx.far_fetch_count(), "sum", float(ctx.download(0).sum()))
    ctx.close()
def run_slabs():
    W, H, n = 200, 96, 3
    ctxs = [Context(W, H, row0=k * 32, rows=32) for k in range(n)]
    for k, c in enumerate(ctxs): c.connect_local(ctxs, k)
    for c in ctxs:
        m = c.get_map(); m.seed = 1234.5; c.set_map(m)
        r = c.get_rain(); r.period = 2; c.set_rain(r)
        c.gen_heightmap()
    for s in range(1, 7):
        for c in ctxs: c.run(1, s * 0.015, 0.0, True)
    for c in ctxs: c.sync()
    print("slabs sum", sum(float(c.download(0).sum()) for c in ctxs), "errors", [c.slab_errors() for c in ctxs])
    for c in ctxs: c.close()
if sys.argv[1] == "slabs": run_slabs()
else: run_one(264, 120)
PY
for tool in memcheck racecheck; do
  for v in 18 13 5; do
    HG_FUSED_VARIANT=$v timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case2.py one > gpurun_out/sanitize2_${tool}_v$v.log 2>&1
    echo "== $tool variant $v: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|^variant ' gpurun_out/sanitize2_${tool}_v$v.log | tr '\n' ' ')"
  done
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case2.py slabs > gpurun_out/sanitize2_${tool}_slabs.log 2>&1
  echo "== $tool slabs (fused push): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|^slabs ' gpurun_out/sanitize2_${tool}_slabs.log | tr '\n' ' ')"
done
