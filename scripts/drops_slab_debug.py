"""Droplet slabs in one process against the whole map (development aid): how many values differ and by how much."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hydro_gen_b200 import Context, _lib, slabs
W, H, n, count, steps = [int(a) for a in sys.argv[1:6]]
rows = H // n
def setup(c):
    m = c.get_map(); m.seed = 1234.5; m.hmap_dims[0], m.hmap_dims[1] = W, H; c.set_map(m); c.gen_heightmap()
whole = Context(W, H, particle_count=count, erosion_type=_lib.HG_PARTICLES); setup(whole)
parts = [Context(W, H, particle_count=count, erosion_type=_lib.HG_PARTICLES, row0=k * rows, rows=rows) for k in range(n)]
for i, s in enumerate(parts): s.connect_local(parts, i)
for s in parts: setup(s)
for k in range(1, steps + 1):
    whole.dispatch_particle(k * 0.015, True)
    for s in parts: s.dispatch_particle(k * 0.015, True)
    if k % 10 == 0 or k == steps:
        for s in parts: s.sync()
        got = slabs.merge_droplets([s.download_particles() for s in parts], [s.particle_owners() for s in parts])
        want = whole.download_particles()
        dp = (np.frombuffer(got.tobytes(), np.uint32) != np.frombuffer(want.tobytes(), np.uint32)).sum()
        msg = f"step {k}: droplet words differing {dp}"
        for f in (0, 2):
            a = np.concatenate([s.download(f) for s in parts], axis=0); b = whole.download(f)
            d = a.view(np.uint32) != b.view(np.uint32)
            msg += f"; field {f}: {int(d.sum())} values differ, max abs {np.abs(a - b).max():.3e}"
        print(msg, flush=True)
