"""Long run of the bench workload (development aid): step rate per chunk, mass, far fetches — to see that the balanced
partition stays stable as the terrain erodes.  python scripts/soak.py [chunks] [steps per chunk]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hydro_gen_b200 import Context

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 10
per = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ctx = Context(4096)
m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
r = ctx.get_rain(); r.period = 16; ctx.set_rain(r)
ctx.gen_heightmap()
t = 0.015
for c in range(chunks):
    ms, k_ms = ctx.run_profiled(per, t, 0.015, True)
    t += per * 0.015
    mass = ctx.mass()
    print(f"steps {(c + 1) * per:6d}: {ms / per:.4f} ms/step, kernel {k_ms:.4f} ms, far/step {ctx.far_fetch_count() / per:.0f}, "
          f"mass rock {mass[0]:.6e} dirt {mass[1]:.6e} water {mass[2]:.4e} sed {mass[3] + mass[4]:.4e}, finite {np.isfinite(mass).all()}", flush=True)
ctx.close()
