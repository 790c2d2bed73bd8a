"""Experiment (development aid): how much of the fused step's time is the thermal outflow path?  The bench workload's
state after the pre-roll, then dispatch_grid timed with (a) the default talus angles, (b) Kalpha so large that nothing
is ever marked, (c) Kalpha = 0 (every downhill neighbour marked).  python scripts/exp_thermal.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydro_gen_b200 import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = Context(n)
m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
r = ctx.get_rain(); r.period = 16; ctx.set_rain(r)
ctx.gen_heightmap()
ctx.run(64, 0.015, 0.015, True)
r.period = 1 << 30; ctx.set_rain(r)
e0 = ctx.get_erosion()
ka = (e0.Kalpha[0], e0.Kalpha[1])
for name, k in (("default", ka), ("never marked", (10.0, 10.0)), ("default again", ka), ("always marked", (0.0, 0.0)), ("default 3", ka)):
    e = ctx.get_erosion(); e.Kalpha[0], e.Kalpha[1] = k; ctx.set_erosion(e)
    for _ in range(5): ctx.dispatch_grid()
    ctx.sync()
    ctx.timer_start()
    for _ in range(40): ctx.dispatch_grid()
    ms = ctx.timer_stop() / 40
    print(f"{name:16s} Kalpha={k}: {ms:.3f} ms/step", flush=True)
ctx.close()
