"""Droplet mode under ncu (development aid): a few Erosion::dispatch_particle steps of BASELINE config 4 (4 Mi droplets on
8192^2) for hmap_dims given on the command line.  ncu --metrics ... python scripts/drops_profile.py 8192"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydro_gen_b200 import Context, _lib
hmap = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N, COUNT = 8192, 4 * 1024 * 1024
ctx = Context(N, particle_count=COUNT, erosion_type=_lib.HG_PARTICLES)
m = ctx.get_map(); m.seed = 1234.5; m.hmap_dims[0], m.hmap_dims[1] = hmap, hmap; ctx.set_map(m)
ctx.gen_heightmap()
ctx.run(steps, 0.015, 0.015, True)
ctx.sync()
ctx.close()
