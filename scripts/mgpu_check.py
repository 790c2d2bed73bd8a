"""Multi-GPU parity check (run under torchrun, one rank per GPU): every rank steps its row slab
with NVLink halo pushes AND the whole map on its own GPU, and compares its rows bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from hydro_gen_b200 import Context, slabs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# MGPU_W / MGPU_ROWS: e.g. 4096 / 1280 makes every slab (and the whole map) large enough for the balanced partition
W, H, STEPS = int(os.environ.get("MGPU_W", 1024)), int(os.environ.get("MGPU_ROWS", 512)) * world, int(os.environ.get("MGPU_STEPS", 48))
row0, rows = slabs.slab_rows(H, world, rank)

def setup(ctx):
    m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
    r = ctx.get_rain(); r.period = 8; ctx.set_rain(r)
    e = ctx.get_erosion(); e.d_t = 0.01; ctx.set_erosion(e)      # livelier water: far fetches across slab borders
    ctx.gen_heightmap()

slab = Context(W, H, device=local, row0=row0, rows=rows)
slabs.connect_ring(slab, dist, world, rank)
setup(slab)
dist.barrier()
slab.run(STEPS, 0.015, 0.015, True)
slab.sync()
whole = Context(W, H, device=local)
setup(whole)
whole.run(STEPS, 0.015, 0.015, True)
ok = True
for f, name in ((0, "heightmap"), (1, "flux"), (3, "sediment")):
    a, b = slab.download(f), whole.download(f)[row0:row0 + rows]
    same = np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ok &= same
    if not same:
        print(f"rank {rank}: {name} differs in {(a.view(np.uint32) != b.view(np.uint32)).sum()} values", flush=True)
print(f"rank {rank}/{world}: rows [{row0},{row0 + rows}) {'bit-identical to the whole-map run' if ok else 'MISMATCH'}; "
      f"halo errors {slab.slab_errors()}, far cells {slab.far_fetch_count()}", flush=True)
dist.barrier()
slab.close(); whole.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
