"""GPU parity at the sizes bench.py measures (BASELINE configs 2-4), through the C ABI, against the oracle.

The small-map tests cover every branch; these cover what only a large map reaches: the bench workload itself
(4096^2: 36 strips, the balanced partition re-cut every step), a >= 16384-column band (142 strips, the uniform
segment path, plane offsets that need more than 26 bits) and the droplet kernels with 64-bit texel indices
and counts >= 2^22.  The oracle runs these in seconds on the host cores (0.08 Gcell-steps/s).
"""
import numpy as np
import pytest

import oracle
from hydro_gen_b200 import Context, _lib
from tests.util import DT_TIME, FIELDS, SEED, assert_bit_equal, copy_state, max_rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north star: per-field max relative error after 1 step


def _grid_pair(W, H, period):
    ref = oracle.World(W, H, seed=SEED)
    ref.gen_heightmap()
    ref.rain.period = period
    ctx = Context(W, H)
    m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
    r = ctx.get_rain(); r.period = period; ctx.set_rain(r)
    ctx.gen_heightmap()
    return ctx, ref


def _compare_grid(ctx, ref, what):
    for name in ("heightmap", "flux", "sediment"):
        got, want = ctx.download(FIELDS[name]), ref.get(FIELDS[name])
        assert_bit_equal(got, want, f"{what}: {name}")
        assert max_rel_err(got, want) <= TOL


def test_bench_workload_4096_bit_exact(built):
    """BASELINE config[1] exactly as bench.py builds it (generated 4096^2 terrain, seed 1234.5, main-loop steps
    through hg_run with rain): 6 steps with rain on steps 2, 4 and 6 so water, flux and sediment are live from the
    second step on; H, F and S equal the oracle's bit for bit, the generated terrain included."""
    ctx, ref = _grid_pair(4096, 4096, 2)
    _compare_grid(ctx, ref, "generated terrain 4096^2")
    ctx.far_fetch_count()
    for s in range(1, 7):
        t = float(np.float32(s) * np.float32(DT_TIME))
        ref.step(t)
        ctx.run(1, t, 0.0, True)
    _compare_grid(ctx, ref, "4096^2, 6 main-loop steps")
    assert ctx.steps == ref.steps == 6
    assert ctx.download(0)[..., 2].max() > 0.0
    ctx.close(); ref.close()


def test_wide_band_16384_columns_bit_exact(built):
    """16384 columns (142 strips of the fused kernel; the balanced partition is off at this width, so this is the
    uniform-segment path of BASELINE configs 3 and 5) x 512 rows, 6 wet steps."""
    ctx, ref = _grid_pair(16384, 512, 2)
    for s in range(1, 7):
        t = float(np.float32(s) * np.float32(DT_TIME))
        ref.step(t)
        ctx.run(1, t, 0.0, True)
    _compare_grid(ctx, ref, "16384x512, 6 steps")
    ctx.close(); ref.close()


def _particle_pair(n, count, hmap):
    ref = oracle.World(n, particle_count=count, erosion_type=1, seed=SEED)
    ref.gen_heightmap()
    ref.map.hmap_dims[0], ref.map.hmap_dims[1] = hmap, hmap
    ctx = Context(n, particle_count=count, erosion_type=_lib.HG_PARTICLES)
    ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(ref.map)))
    ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(ref.erosion)))
    ctx.gen_heightmap()
    return ctx, ref


def _multi_hit_mask(parts, n):
    """texels touched by the corners of more than one live droplet (positions after the move pass)"""
    live = parts["iters"] != 0
    bx = parts["position"][live, 0].astype(np.int64)
    by = parts["position"][live, 1].astype(np.int64)
    hits = np.zeros(n * n, np.int32)
    for dx, dy in ((0, 0), (1, 0), (1, 1), (0, 1)):
        x, y = bx + dx, by + dy
        ok = (x >= 0) & (x < n) & (y >= 0) & (y < n)
        np.add.at(hits, y[ok] * n + x[ok], 1)
    return (hits > 1).reshape(n, n), live


def test_droplets_262144_on_2048_bit_exact_where_order_free(built):
    """particle.glsl + particle_erosion.glsl with 262 144 droplets on a 2048^2 map (hmap_dims = map, the sparse
    regime of BASELINE config 4).  The move pass is pointwise: the droplet buffer must equal the oracle's byte for
    byte.  The erode pass is order-free on every texel that only one droplet touches in the step: there H.rgb and
    the momentum accumulator must match bit for bit; texels shared by several droplets depend on the order of the
    additions (lock order in the reference) and must agree within float reassociation."""
    n, count = 2048, 262144
    ctx, ref = _particle_pair(n, count, n)
    for k in range(1, 4):
        t = float(np.float32(k) * np.float32(DT_TIME))
        H0, M0 = ctx.download(0), ctx.download(2)
        ref.set(0, H0); ref.set(2, M0)                       # identical state before every droplet step
        ref_parts = ref.particles()
        ref_parts[...] = ctx.download_particles()
        ctx.dispatch_particle_pass(0, t, True); ref.particle_pass(0, t, True)
        gp, wp = ctx.download_particles(), ref.particles().copy()
        assert gp.tobytes() == wp.tobytes(), f"droplets differ after move {k}"
        shared, live = _multi_hit_mask(gp, n)
        assert live.sum() == count and 0 < shared.sum() < 0.2 * n * n
        ctx.dispatch_particle_pass(1, t, True); ref.particle_pass(1, t, True)
        for f, chans in ((0, (0, 1, 2)), (2, (0, 1, 2, 3))):
            got, want = ctx.download(f), ref.get(f)
            for ch in chans:
                g, w = got[..., ch], want[..., ch]
                assert np.array_equal(g[~shared].view(np.uint32), w[~shared].view(np.uint32)), f"step {k}: field {f} channel {ch} differs on an order-free texel"
                assert np.abs(g - w).max() <= 2e-5 * (np.abs(w).max() + 1e-30)
        gp, wp = ctx.download_particles(), ref.particles()
        assert np.array_equal(gp["iters"], wp["iters"]) and np.array_equal(gp["to_kill"], wp["to_kill"])
        # a droplet's carried sediment depends on the texel values only through the exhausted-layer clamp
        np.testing.assert_allclose(gp["sediment"], wp["sediment"], rtol=1e-5, atol=1e-9)
    ctx.close(); ref.close()


def test_droplets_4Mi_on_8192(built):
    """BASELINE config 4 at full size: 4 194 304 droplets on 8192^2 (64-bit texel indices, count = 2^22), two
    Erosion::dispatch_particle steps with hmap_dims = map.  Droplet life cycle (iters, to_kill) equals the oracle's;
    positions and fields agree within float reassociation of the contended texels (the reference is lock-order
    dependent there); the total of what the droplets put on the map matches."""
    n, count = 8192, 4194304
    ctx, ref = _particle_pair(n, count, n)
    assert_bit_equal(ctx.download(0), ref.get(0), "generated terrain 8192^2")
    for k in range(1, 3):
        t = float(np.float32(k) * np.float32(DT_TIME))
        ctx.dispatch_particle(t, True); ref.dispatch_particle(t, True)
    gp, wp = ctx.download_particles(), ref.particles()
    assert np.array_equal(gp["iters"], wp["iters"]) and np.array_equal(gp["to_kill"], wp["to_kill"])
    np.testing.assert_allclose(gp["position"], wp["position"], rtol=0, atol=1e-3)
    got, want = ctx.download(0), ref.get(0)
    for ch, name in enumerate(("rock", "dirt", "water")):
        scale = np.abs(want[..., ch]).max() + 1e-30
        assert np.abs(got[..., ch] - want[..., ch]).max() / scale < 2e-5, name
    for ch in range(3):
        a, b = got[..., ch].sum(dtype=np.float64), want[..., ch].sum(dtype=np.float64)
        assert abs(a - b) <= 1e-7 * abs(b) + 1e-6, f"total of channel {ch}: {a} vs {b}"
    gm, wm = ctx.download(2), ref.get(2)
    assert np.abs(gm - wm).max() <= 2e-5 * (np.abs(wm).max() + 1e-30)
    ctx.close(); ref.close()
