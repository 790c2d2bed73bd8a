import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydro_gen_b200 import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = Context(n)
m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
ctx.gen_heightmap()
ctx.dispatch_grid_rain(0.5)
ctx.dispatch_grid()
ctx.sync()
print("ok", ctx.mass())
