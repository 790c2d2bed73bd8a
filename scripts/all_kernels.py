"""Launch every kernel of the library a few times in steady state, between cudaProfilerStart/Stop, so that
    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/all_kernels python scripts/all_kernels.py
captures one representative launch of each (scripts/gpu_all_kernels.sh; summarised by scripts/ncu_kernels_summary.py
into profiles/<tag>_all_kernels.txt).  Grid kernels at BASELINE config[1] (4096^2), droplets at config[4] (4 Mi on 8192^2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hydro_gen_b200 import Context, _lib

rt = torch.cuda.cudart()
N = int(os.environ.get("HG_N", 4096))

GRID = os.environ.get("HG_SKIP_GRID", "0") != "1"
ctx = Context(N if GRID else 64)
m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
r = ctx.get_rain(); r.period = 4; ctx.set_rain(r)
ctx.gen_heightmap()
ctx.run(64 if GRID else 0, 0.0, 0.015, True)            # wet the terrain; the planner settles
ctx.sync()
img = ctx.download(0)
if GRID: rt.cudaProfilerStart()
ctx.run(4, 64 * 0.015, 0.015, True)      # k_rain, k_far_fixup, k_fused_ws, k_plan_segments
ctx.set_schedule(_lib.SCHEDULE_PASSES)
ctx.run(1, 68 * 0.015, 0.015, False)     # the eight 1:1 pass kernels
ctx.set_schedule(_lib.SCHEDULE_FUSED)
ctx.upload(0, img)                       # k_unpack
ctx.download(0, img)                     # k_pack
ctx.mass()                               # k_mass
ctx.gen_heightmap()                      # k_heightmap
ctx.sync()
rt.cudaProfilerStop()
ctx.close()

D, COUNT = 8192, 4 * 1024 * 1024
HMAP = int(os.environ.get("HG_HMAP", D))      # 1024 = the reference's default hmap_dims: ~4 droplets per cell in a corner
ctx = Context(D, particle_count=COUNT, erosion_type=_lib.HG_PARTICLES)
m = ctx.get_map(); m.seed = 1234.5; m.hmap_dims[0], m.hmap_dims[1] = HMAP, HMAP; ctx.set_map(m)
ctx.gen_heightmap()
ctx.run(40, 0.015, 0.015, True)          # droplets in flight, momentum map populated
ctx.sync()
rt.cudaProfilerStart()
ctx.run(2, 41 * 0.015, 0.015, True)      # k_particle_move, k_particle_erode, k_fused_ws<..., DROPS>
ctx.sync()
rt.cudaProfilerStop()
ctx.close()
print("done")
