"""GPU tuning aid: time the fused step for every CTA-shape variant x segment length, after checking
each against the PASSES schedule bit for bit.  usage: tune_fused.py [sizes...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hydro_gen_b200 import Context, _lib

def make(n, h=None, variant=None, seg=None):
    for k, v in (("HG_FUSED_VARIANT", variant), ("HG_FUSED_SEG", seg)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    ctx = Context(n, h)
    m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
    r = ctx.get_rain(); r.period = 4; ctx.set_rain(r)
    return ctx

def check(variant, seg):
    W, H = 1024, 640
    ref = make(W, H); ref.set_schedule(_lib.SCHEDULE_PASSES); ref.gen_heightmap(); ref.run(24, 0.015, 0.015, True)
    ctx = make(W, H, variant, seg); ctx.gen_heightmap(); ctx.run(24, 0.015, 0.015, True)
    ok = all(np.array_equal(ctx.download(f).view(np.uint32)[..., :3], ref.download(f).view(np.uint32)[..., :3]) for f in (0, 1, 3))
    ctx.close(); ref.close()
    return ok

def timeit(n, variant, seg, steps=30):
    ctx = make(n, None, variant, seg)
    ctx.gen_heightmap()
    ctx.run(40, 0.015, 0.015, True)
    r = ctx.get_rain(); r.period = 1 << 30; ctx.set_rain(r)
    ctx.run(5, 1.0, 0.015, True); ctx.sync()
    k_ms = ctx.profile_fused(steps)
    ctx.timer_start()
    for _ in range(steps): ctx.dispatch_grid()
    ms = ctx.timer_stop() / steps
    far = ctx.far_fetch_count() / (steps * 2 + 45)
    ctx.close()
    return k_ms, ms, far

if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [4096]
    names = ["128x4", "128x3", "192x2", "224x2", "224x1"]
    for v in range(5):
        for seg in (128, 256, 512):
            ok = check(v, seg)
            line = f"variant {names[v]} seg {seg}: parity {'OK' if ok else 'FAIL'}"
            for n in sizes:
                k_ms, ms, far = timeit(n, v, seg)
                line += f" | N={n}: kernel {k_ms:.3f} ms, step {ms:.3f} ms, {n*n/k_ms/1e6:.1f} Gcell/s ({72*n*n/k_ms/1e6/6550.4*100:.1f}%), far/step {far:.0f}"
            print(line, flush=True)
