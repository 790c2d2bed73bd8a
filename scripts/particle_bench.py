"""BASELINE config 4: droplet mode, 4 Mi droplets on 8192^2 (development aid; not the bench.py line).
4a: the reference's default hmap_dims = (1024, 1024): droplets live in a 1022^2 corner, ~4 per cell, heavy contention.
4b: hmap_dims = (8192, 8192): sparse.   python scripts/particle_bench.py [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydro_gen_b200 import Context, _lib

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N, COUNT = 8192, 4 * 1024 * 1024
for name, hmap in (("4a faithful (hmap 1024)", 1024), ("4b corrected (hmap 8192)", 8192)):
    ctx = Context(N, particle_count=COUNT, erosion_type=_lib.HG_PARTICLES)
    m = ctx.get_map(); m.seed = 1234.5; m.hmap_dims[0], m.hmap_dims[1] = hmap, hmap; ctx.set_map(m)
    ctx.gen_heightmap()
    ctx.run(20, 0.015, 0.015, True)
    ctx.sync()
    ctx.timer_start()
    ctx.run(steps, 21 * 0.015, 0.015, True)
    ms = ctx.timer_stop() / steps
    print(f"{name}: {ms:.3f} ms/step, {COUNT / ms / 1e3:.1f} Mdroplet-steps/s ({20 * COUNT / ms / 1e6:.1f} G atomic updates/s), {N * N / ms / 1e6:.2f} Gcell-steps/s of the grid part "
          f"(thermal x2 + smoothing + momentum decay), launches/step {ctx.launch_count / (steps + 20):.1f}", flush=True)
    ctx.close()
