"""GPU parity tests of the grid erosion path, through the C ABI.

The CUDA library is compiled -fmad=false and the oracle -ffp-contract=off, and both use
include/hg_defined_math.h for atan/exp/sin, so the bar here is stronger than the north
star's 1e-5: every field must match the oracle BIT FOR BIT.  The north-star tolerance is
asserted as well (max_rel_err <= 1e-5) so the stated gate is visible in the test.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from hydro_gen_b200 import Context, _lib
from tests.util import DT_TIME, FIELDS, SEED, assert_bit_equal, copy_state, max_rel_err, wet_world

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north star: per-field max relative error after 1 step


@pytest.fixture(scope="module")
def wet256():
    w = wet_world(256, 300)
    yield w
    w.close()


def _compare(ctx, ref, names, what):
    for name in names:
        got, want = ctx.download(FIELDS[name]), ref.get(FIELDS[name])
        assert max_rel_err(got, want) <= TOL, f"{what}: {name} beyond the north-star tolerance"
        assert_bit_equal(got, want, f"{what}: {name}")


def test_heightmap_init_bit_exact(built):
    """heightmap.glsl via hg_gen_heightmap vs the oracle (default map settings, injected seed)."""
    for n, seed in ((256, SEED), (64, 77.25)):
        ctx = Context(n)
        m = ctx.get_map(); m.seed = seed; ctx.set_map(m)
        ctx.gen_heightmap()
        ref = oracle.World(n, seed=seed); ref.gen_heightmap()
        _compare(ctx, ref, ("heightmap", "flux", "sediment"), f"init {n}")
        ctx.close(); ref.close()


@pytest.mark.parametrize("variant", ["uplift_terrace", "round_slope_warp2"])
def test_heightmap_init_variants(built, variant):
    """every branch of heightmap.glsl:85-131 (uplift/ridged sfbm, masks, terrace, 2-level warp)"""
    n = 128
    ctx = Context(n)
    ref = oracle.World(n, seed=SEED)
    m = ctx.get_map(); m.seed = SEED
    if variant == "uplift_terrace":
        m.uplift, m.terrace, m.terrace_scale = 1, 6, 0.5
    else:
        m.mask_round, m.mask_slope, m.domain_warp, m.mask_exp = 1, 1, 2, 0
    ctx.set_map(m)
    C.memmove(C.addressof(ref.map), C.addressof(m), C.sizeof(m))
    ctx.gen_heightmap(); ref.gen_heightmap()
    _compare(ctx, ref, ("heightmap",), variant)
    ctx.close(); ref.close()


def test_rain_bit_exact(built, wet256):
    ctx = Context(256)
    ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(wet256.map)))
    copy_state(wet256, ctx)
    ref = wet_world(256, 0)
    for f in ("heightmap", "flux", "sediment"):
        ref.set(FIELDS[f], wet256.get(FIELDS[f]))
    for t in (0.123, 4.5):
        ctx.dispatch_grid_rain(t)
        ref.dispatch_grid_rain(t)
    _compare(ctx, ref, ("heightmap",), "rain")
    ctx.close(); ref.close()


def _clone_oracle(src, n):
    ref = oracle.World(n, seed=SEED)
    for f in ("heightmap", "flux", "velocity", "sediment"):
        ref.set(FIELDS[f], src.get(FIELDS[f]))
    ref.rain.period = src.rain.period
    ref.steps = src.steps
    return ref


def test_passes_each_dispatch_bit_exact(built, wet256):
    """PASSES schedule: after every one of the reference's dispatches (erosion.cpp:158-200)
    all materialised textures equal the oracle's, including V, TC, TD and H.a."""
    ctx = Context(256)
    ctx.set_schedule(_lib.SCHEDULE_PASSES)
    copy_state(wet256, ctx, ("heightmap", "flux", "velocity", "sediment"))
    ref = _clone_oracle(wet256, 256)
    for p, names in ((oracle.PASS_FLUX, ("heightmap", "flux", "velocity")),
                     (oracle.PASS_EROSION, ("heightmap", "sediment")),
                     (oracle.PASS_SEDIMENT, ("heightmap", "sediment")),
                     (oracle.PASS_THERMAL, ("heightmap", "thermal_c", "thermal_d")),
                     (oracle.PASS_SMOOTH, ("heightmap",))):
        ctx.dispatch_pass(p)
        ref.run_pass(p)
        _compare(ctx, ref, names, f"pass {p}")
    ctx.close(); ref.close()


def test_fused_one_step_bit_exact(built, wet256):
    """The product path: one fused step == the oracle's 8-pass step, from a wet state in
    which every branch is live (north-star gate (i))."""
    ctx = Context(256)
    copy_state(wet256, ctx)
    ref = _clone_oracle(wet256, 256)
    ctx.dispatch_grid(); ref.dispatch_grid()
    _compare(ctx, ref, ("heightmap", "flux", "sediment"), "fused step")
    assert ctx.launch_count > 0
    ctx.close(); ref.close()


def test_fused_run_with_rain_and_far_fetch(built, wet256):
    """48 main-loop iterations (rain every 16) stay bit-identical, and the far-fetch path
    (back-trace leaving the +-1 window) is actually exercised."""
    ctx = Context(256)
    copy_state(wet256, ctx)
    r = ctx.get_rain(); r.period = 16; ctx.set_rain(r)
    ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(wet256.map)))
    ctx.steps = wet256.steps
    ref = _clone_oracle(wet256, 256)
    ctx.far_fetch_count()
    t0 = (wet256.steps + 1) * DT_TIME
    ctx.run(48, t0, DT_TIME, True)
    for k in range(48):
        ref.step(np.float32(t0) + np.float32(k) * np.float32(DT_TIME))
    _compare(ctx, ref, ("heightmap", "flux", "sediment"), "48 steps")
    assert ctx.far_fetch_count() > 0, "far-fetch path never taken: the test state is too tame"
    assert ctx.steps == ref.steps
    ctx.close(); ref.close()


@pytest.mark.parametrize("shape", [(8, 8), (16, 8), (8, 40), (136, 24), (120, 264), (248, 72)])
def test_fused_ragged_sizes(built, shape):
    """Smallest legal map, non-square maps, widths around the strip width (116 owned columns)
    and heights around the row-segment length: strip / segment edges and the map border."""
    W, H = shape
    ref = wet_world(H, 40, period=4, width=W)
    ctx = Context(W, H)
    src = _clone_like(ref, W, H)
    copy_state(src, ctx)
    for _ in range(3):
        ctx.dispatch_grid(); src.dispatch_grid()
    _compare(ctx, src, ("heightmap", "flux", "sediment"), f"{W}x{H}")
    ctx.close(); ref.close(); src.close()


def _clone_like(src, W, H):
    ref = oracle.World(W, H, seed=SEED)
    for f in ("heightmap", "flux", "velocity", "sediment"):
        ref.set(FIELDS[f], src.get(FIELDS[f]))
    return ref


def test_fused_equals_passes_schedule(built, wet256):
    """Both schedules of the library give the same bits (and can be switched between steps)."""
    a, b = Context(256), Context(256)
    b.set_schedule(_lib.SCHEDULE_PASSES)
    copy_state(wet256, a); copy_state(wet256, b, ("heightmap", "flux", "velocity", "sediment"))
    for _ in range(4):
        a.dispatch_grid(); b.dispatch_grid()
    for name in ("heightmap", "flux", "sediment"):
        assert_bit_equal(a.download(FIELDS[name]), b.download(FIELDS[name]), f"fused vs passes: {name}")
    a.set_schedule(_lib.SCHEDULE_PASSES)
    a.dispatch_grid(); b.dispatch_grid()
    assert_bit_equal(a.download(0), b.download(0), "after switching schedule")
    a.close(); b.close()


def test_dry_default_config(built):
    """BASELINE config 1 as the reference runs it: period 512, so 60 steps never rain and
    only thermal + smoothing act."""
    n = 256
    ctx = Context(n)
    m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
    ctx.gen_heightmap()
    ctx.run(60, DT_TIME, DT_TIME, True)
    ref = oracle.World(n, seed=SEED); ref.gen_heightmap()
    for s in range(1, 61):
        ref.step(s * DT_TIME)
    _compare(ctx, ref, ("heightmap", "flux", "sediment"), "dry 60 steps")
    assert ctx.download(0)[..., 2].max() == 0.0
    ctx.close(); ref.close()


def test_steep_terrain_thermal_marks(built):
    """Talus angles lowered so thermal slippage marks many neighbours (atan path, sharpness,
    bk division) in both layers; dry."""
    n = 128
    ref = oracle.World(n, seed=SEED); ref.gen_heightmap()
    H = ref.get(0)
    rng = np.random.default_rng(5)
    H[..., 0] += rng.random((n, n), dtype=np.float32) * 6.0
    H[..., 1] += rng.random((n, n), dtype=np.float32) * 2.0
    H[..., 3] = H[..., 0] + H[..., 1] + H[..., 2]
    ref.set(0, H)
    ref.erosion.Kalpha[0], ref.erosion.Kalpha[1] = 0.9, 0.3
    ctx = Context(n)
    ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(ref.erosion)))
    copy_state(ref, ctx)
    for _ in range(5):
        ctx.dispatch_grid(); ref.dispatch_grid()
    _compare(ctx, ref, ("heightmap", "flux", "sediment"), "steep thermal")
    ctx.close(); ref.close()


def test_settings_hot_reload_between_steps(built, wet256):
    """Erosion_settings / Rain_settings::push_data between steps (rendering.cpp:200-201,255-257: the UI pushes the
    whole struct while the simulation runs): new parameters apply from the next dispatch on, with no re-creation —
    including the ones baked into the step parameters (d_t, Kalpha -> talus tangents) and the rain period."""
    ctx = Context(256)
    copy_state(wet256, ctx)
    ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(wet256.map)))
    ctx.steps = wet256.steps
    ref = _clone_oracle(wet256, 256)
    t = np.float32((wet256.steps + 1) * DT_TIME)
    for d_t, kc, kalpha, period, amount in ((0.001, 0.2, (1.3, 0.6), 16, 0.01), (0.004, 0.35, (0.8, 0.4), 3, 0.02),
                                            (0.0005, 0.1, (1.0, 0.7), 5, 0.005)):
        e = ref.erosion
        e.d_t, e.Kc, e.Kalpha[0], e.Kalpha[1] = d_t, kc, kalpha[0], kalpha[1]
        ref.rain.period, ref.rain.amount = period, amount
        ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(ref.erosion)))
        ctx.set_rain(_lib.RainData.from_buffer_copy(bytes(ref.rain)))
        for _ in range(7):
            ctx.run(1, float(t), DT_TIME, True)
            ref.step(t)
            t = np.float32(t + np.float32(DT_TIME))
        _compare(ctx, ref, ("heightmap", "flux", "sediment"), f"after push_data(d_t={d_t}, period={period})")
    assert ctx.steps == ref.steps
    ctx.close(); ref.close()


@pytest.mark.parametrize("kspeed", [1e-30, 3e8])
def test_thermal_outflow_generic_division(built, kspeed):
    """Kspeed so small (or so large) that S = d_t*Kspeed*sharpness*H/2 leaves the range in which the
    kernel's shared-reciprocal division is proven exact: hg_thermal_outflow must take its generic
    IEEE-division path (out of line on the device) and still match the oracle bit for bit, denormal
    outflows included.  One step for the large value (the terrain explodes afterwards)."""
    n = 128
    ref = oracle.World(n, seed=SEED); ref.gen_heightmap()
    H = ref.get(0)
    rng = np.random.default_rng(9)
    H[..., 0] += rng.random((n, n), dtype=np.float32) * 6.0
    H[..., 1] += rng.random((n, n), dtype=np.float32) * 2.0
    H[..., 3] = H[..., 0] + H[..., 1] + H[..., 2]
    ref.set(0, H)
    ref.erosion.Kalpha[0], ref.erosion.Kalpha[1] = 0.9, 0.3
    ref.erosion.Kspeed[0], ref.erosion.Kspeed[1] = kspeed, kspeed
    ctx = Context(n)
    ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(ref.erosion)))
    copy_state(ref, ctx)
    for _ in range(3 if kspeed < 1 else 1):
        ctx.dispatch_grid(); ref.dispatch_grid()
    got = ctx.download(0)
    assert np.isfinite(got).all()
    if kspeed > 1:
        assert not np.array_equal(got[..., :2], H[..., :2])      # something did flow (1e-30 outflows vanish in the sum)
    _compare(ctx, ref, ("heightmap", "flux", "sediment"), f"generic thermal division, Kspeed {kspeed}")
    ctx.close(); ref.close()


def test_exhausted_dirt_layer(built, wet256):
    """Thin dirt + aggressive dissolving: the 'layer went negative -> continue into rock'
    branch of hydro_erosion.glsl:68-77."""
    n = 256
    ref = _clone_oracle(wet256, n)
    H = ref.get(0); H[..., 1] = 1e-4; H[..., 3] = H[..., 0] + H[..., 1] + H[..., 2]; ref.set(0, H)
    ref.erosion.Ks[1] = 50.0; ref.erosion.Kc = 5.0
    ctx = Context(n)
    ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(ref.erosion)))
    copy_state(ref, ctx)
    for _ in range(3):
        ctx.dispatch_grid(); ref.dispatch_grid()
    _compare(ctx, ref, ("heightmap", "flux", "sediment"), "exhausted dirt")
    ctx.close(); ref.close()


def test_errors_are_loud(built):
    L = _lib.load()
    assert not L.hg_create(100, 100, 0, 0, 0) and b"multiple of 8" in L.hg_last_error()
    assert not L.hg_create(64, 64, 0, 1, 0)      # particle mode without droplets
    ctx = Context(64)
    with pytest.raises(_lib.HydrogenError):
        ctx.download(_lib.FIELD_VELOCITY)        # not materialised by the fused schedule
    with pytest.raises(_lib.HydrogenError):
        ctx.dispatch_particle(0.0)
    with pytest.raises(_lib.HydrogenError):
        ctx.dispatch_pass(0)
    ctx.close()


def test_mass_and_roundtrip(built, wet256):
    ctx = Context(256)
    copy_state(wet256, ctx)
    H, S = wet256.get(0), wet256.get(3)
    want = np.array([H[..., 0].sum(dtype=np.float64), H[..., 1].sum(dtype=np.float64), H[..., 2].sum(dtype=np.float64),
                     S[..., 0].sum(dtype=np.float64), S[..., 1].sum(dtype=np.float64)])
    np.testing.assert_allclose(ctx.mass(), want, rtol=1e-12)
    assert_bit_equal(ctx.download(0), H, "upload/download round trip")
    ctx.close()


def test_step_host_async_pipelined(built, wet256):
    """hg_step_host_async: host images in -> one fused step -> host images out, three calls in
    flight back to back; every output equals the plain upload/dispatch/download result."""
    from hydro_gen_b200 import PinnedBuffer
    names = ("heightmap", "flux", "sediment")
    ref_ctx = Context(256)
    copy_state(wet256, ref_ctx)
    ref_ctx.dispatch_grid()
    want = [ref_ctx.download(FIELDS[n]) for n in names]
    ctx = Context(256)
    pins = [PinnedBuffer((256, 256, 4)) for _ in names]
    outs = [[PinnedBuffer((256, 256, 4)) for _ in names] for _ in range(3)]
    for p, n in zip(pins, names):
        p.array[...] = wet256.get(FIELDS[n])
    for k in range(3):
        ctx.step_host_async([p.array for p in pins], [p.array for p in outs[k]])
    ctx.sync()
    for k in range(3):
        for p, w_, n in zip(outs[k], want, names):
            assert_bit_equal(p.array, w_, f"pipelined host step {k}: {n}")
    for p in pins + [q for o in outs for q in o]:
        p.free()
    ctx.close(); ref_ctx.close()


def test_balanced_partition_large_slab_bit_exact(built):
    """A slab large enough for the balanced partition of the fused step (k_plan_segments: 3 CTAs per SM, every
    strip re-cut each step into segments of equal forecast cost from the durations the previous step's CTAs
    reported).  The cut changes from step to step; the fields must not: 8 wet steps at 4096 x 1280 against the
    oracle, bit for bit, and against the same run with uniform segments (HG_FUSED_SEG)."""
    import os
    W, H = 4096, 1280
    ref = oracle.World(W, H, seed=SEED)
    ref.gen_heightmap()
    ref.rain.period = 2
    ctx = Context(W, H)
    m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
    r = ctx.get_rain(); r.period = 2; ctx.set_rain(r)
    ctx.gen_heightmap()
    for s in range(1, 9):
        t = float(np.float32(s) * np.float32(DT_TIME))
        ref.step(t)
        ctx.run(1, t, 0.0, True)
    _compare(ctx, ref, ("heightmap", "flux", "sediment"), "balanced partition, 8 steps")
    assert ctx.far_fetch_count() >= 0
    ctx.close(); ref.close()


def test_pack_device_is_the_texture_image(built, wet256):
    """hg_pack_device -- the device-side half of hg_publish_gl (the other half is a copy into the mapped GL array,
    which needs an OpenGL context) -- writes the field as an RGBA32F image into caller-provided DEVICE memory: it must
    be byte for byte what hg_download returns, H.a included, for the grid fields and for the droplet mode's
    texture-layout images."""
    import torch
    ctx = Context(256)
    copy_state(wet256, ctx)
    ctx.dispatch_grid()
    for name in ("heightmap", "flux", "sediment"):
        dev = torch.full((256, 256, 4), -1.0, dtype=torch.float32, device="cuda")
        ctx.pack_device(FIELDS[name], dev.data_ptr())
        ctx.sync()
        assert_bit_equal(dev.cpu().numpy(), ctx.download(FIELDS[name]), f"pack_device {name}")
    H = ctx.download(0)
    assert np.array_equal(H[..., 3], (H[..., 0] + H[..., 1]) + H[..., 2])      # H.a as the last writer of a step leaves it
    ctx.close()
    p = Context(128, particle_count=1024, erosion_type=_lib.HG_PARTICLES)
    m = p.get_map(); m.seed = SEED; m.hmap_dims[0], m.hmap_dims[1] = 128, 128; p.set_map(m)
    p.gen_heightmap()
    for k in range(3):
        p.dispatch_particle((k + 1) * DT_TIME, True)
    for name in ("heightmap", "velocity"):
        dev = torch.zeros((128, 128, 4), dtype=torch.float32, device="cuda")
        p.pack_device(FIELDS[name], dev.data_ptr())
        p.sync()
        assert_bit_equal(dev.cpu().numpy(), p.download(FIELDS[name]), f"pack_device {name} (droplet mode)")
    p.close()
