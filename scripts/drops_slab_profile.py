"""Droplet slabs in ONE process (peer pointers on one GPU) under ncu: per-kernel durations of the sharded dispatch
(development aid).  ncu --metrics gpu__time_duration.sum python scripts/drops_slab_profile.py [n_slabs] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydro_gen_b200 import Context, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
N, COUNT = 8192, 4 * 1024 * 1024
rows = N // n
parts = [Context(N, N, particle_count=COUNT, erosion_type=_lib.HG_PARTICLES, row0=k * rows, rows=rows) for k in range(n)]
for i, s in enumerate(parts): s.connect_local(parts, i)
for s in parts:
    m = s.get_map(); m.seed = 1234.5; m.hmap_dims[0], m.hmap_dims[1] = N, N; s.set_map(m); s.gen_heightmap()
for k in range(1, steps + 1):
    for s in parts: s.dispatch_particle(k * 0.015, True)
for s in parts: s.sync()
print("errors", [s.slab_errors() for s in parts])
