"""Quick device-side timing of one grid step for both schedules (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hydro_gen_b200 import Context, _lib

def run(n, schedule, steps=20, warm=5, dry=False):
    ctx = Context(n)
    m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
    r = ctx.get_rain(); r.period = (1 << 30) if dry else 4; ctx.set_rain(r)
    ctx.set_schedule(schedule)
    ctx.gen_heightmap()
    ctx.run(warm * 4, 0.015, 0.015, True)      # a few rains so water is live
    r.period = 1 << 30; ctx.set_rain(r)
    ctx.sync()
    ctx.timer_start()
    for _ in range(steps):
        ctx.dispatch_grid()
    ms = ctx.timer_stop() / steps
    far = ctx.far_fetch_count()
    cells = n * n
    print(f"N={n} schedule={'fused' if schedule == 0 else 'passes'}: {ms:.3f} ms/step, {cells / ms / 1e6:.2f} Gcell-steps/s, "
          f"{72 * cells / ms / 1e6:.0f} GB/s algorithmic ({72 * cells / ms / 1e6 / 6463.7 * 100:.1f}% of measured HBM), far cells/step {far / (steps + warm * 4):.0f}")
    ctx.close()

if __name__ == "__main__":
    dry = "--dry" in sys.argv
    fused_only = "--fused" in sys.argv
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [1024, 4096]
    for n in sizes:
        if not fused_only:
            run(n, 1, dry=dry)
        run(n, 0, dry=dry)
