"""Run under HG_FMAD=1 by tests/test_fmad_build.py: the opt-in CONTRACTED build of the library (-fmad=true,
libhydrogen_b200_fmad.so) against the non-contracting oracle.  Prints one JSON line with the north star's two gates:
(i) per-field max relative error after ONE step from identical wet inputs, (ii) mass / mean-height drift after 1000
steps.  Test infrastructure (imports the oracle)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from hydro_gen_b200 import Context, _lib
from tests.util import DT_TIME, FIELDS, SEED, copy_state, max_rel_err, wet_world
from tests.test_drift import PERIOD, STEPS, _metrics, _run_oracle, _time

out = {"version": _lib.load().hg_version().decode()}
# (i) one step from a wet state in which every branch is live
w = wet_world(256, 300)
ctx = Context(256)
copy_state(w, ctx)
ctx.dispatch_grid(); w.dispatch_grid()
out["one_step"] = {name: max_rel_err(ctx.download(FIELDS[name]), w.get(FIELDS[name])) for name in ("heightmap", "flux", "sediment")}
out["one_step_bit_identical"] = all(np.array_equal(ctx.download(FIELDS[n]).view(np.uint32), w.get(FIELDS[n]).view(np.uint32)) for n in ("heightmap", "flux", "sediment"))
ctx.close(); w.close()
# (ii) 1000 steps
n = 256
H, S = _run_oracle(n, False)
ctx = Context(n)
m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
r = ctx.get_rain(); r.period = PERIOD; ctx.set_rain(r)
ctx.gen_heightmap()
for s in range(1, STEPS + 1):
    ctx.run(1, _time(s), 0.0, True)
mass_rel, terr_abs = _metrics(ctx.download(0), ctx.download(3), H, S)
out["drift_1000"] = {"mass_rel": mass_rel, "mean_abs_terrain": terr_abs}
ctx.close()
print(json.dumps(out))
