"""The oracle restatement (oracle/hg_oracle.c) against the REFERENCE'S OWN shaders, compiled for the CPU
from /root/reference/glsl/*.glsl by oracle/refshader/build_ref.py (oracle/_ref/libhg_refshaders.so; the
built library travels to the GPU box, the sources are only needed to build it).

Every dispatch of Erosion::dispatch_grid (src/erosion.cpp:158-200) and Erosion::dispatch_grid_rain is run
by both on the same inputs; all six textures must agree BIT FOR BIT after each dispatch, on a wet state
in which every branch is live, and after whole multi-step runs.  What this pins: expression structure and
association order, constants, branch structure, neighbour indexing, border rules, texture bindings and
swap order — everything the restatement could have got wrong.  (Built-ins GLSL leaves open — atan, the
formulas of min/max/mix/... — are the documented ones of include/hg_defined_math.h on both sides.)"""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import refshaders
from tests.util import DT_TIME, SEED, assert_bit_equal, wet_world

pytestmark = pytest.mark.skipif(not refshaders.available(), reason="oracle/_ref/libhg_refshaders.so not built (needs /root/reference)")

FIELDS = ("heightmap", "flux", "velocity", "sediment", "thermal_c", "thermal_d")


def _ref_from(orc):
    """RefWorld with the oracle's current read textures and settings."""
    e = oracle.ErosionData.from_buffer_copy(bytes(orc.erosion))
    r = oracle.RainData.from_buffer_copy(bytes(orc.rain))
    m = oracle.MapSettingsData.from_buffer_copy(bytes(orc.map))
    ref = refshaders.RefWorld(orc.W, e, r, m)
    for k, f in enumerate(FIELDS):
        getattr(ref, f).read[...] = orc.get(k)
    return ref


def _compare(ref, orc, what, fields=FIELDS):
    for k, f in enumerate(FIELDS):
        if f in fields:
            assert_bit_equal(orc.get(k), getattr(ref, f).read, f"{what}: {f}")


@pytest.fixture(scope="module")
def wet():
    w = wet_world(96, 200, period=8)
    yield w
    w.close()


def test_each_dispatch_matches_the_reference_shaders(wet):
    orc = wet
    ref = _ref_from(orc)
    for which, name in enumerate(refshaders.RefWorld.PASSES):
        getattr(ref, name)()
        orc.run_pass(which)
        _compare(ref, orc, name)
    # the wet state exercises the hydraulics; check it is not a trivial comparison
    assert orc.get(0)[..., 2].max() > 0 and np.abs(orc.get(1)).max() > 0 and orc.get(3)[..., :2].max() > 0


def test_rain_matches_the_reference_shader(wet):
    orc = wet
    ref = _ref_from(orc)
    t = 201 * DT_TIME
    ref.dispatch_grid_rain(t)
    orc.dispatch_grid_rain(t)
    _compare(ref, orc, "rain", ("heightmap",))


@pytest.mark.parametrize("variant", ["default", "uplift_terrace", "round_slope_warp2", "exp_power_warp1"])
def test_heightmap_init_matches_the_reference_shader(variant):
    """heightmap.glsl (State::World::gen_heightmap, src/state.cpp:116-147) with the noise functions it reaches
    (simplex_noise.glsl: hash, noised, perlfbm, erosion_perlfbm, gln_simplex, gln_sfbm) in every mask / warp /
    terrace branch, against the oracle."""
    n = 64
    orc = oracle.World(n, seed=SEED)
    m = orc.map
    if variant == "uplift_terrace":
        m.uplift, m.terrace, m.terrace_scale = 1, 6, 0.5
    elif variant == "round_slope_warp2":
        m.mask_round, m.mask_slope, m.domain_warp = 1, 1, 2
    elif variant == "exp_power_warp1":
        m.mask_exp, m.mask_power, m.domain_warp, m.uplift = 1, 1, 1, 1
    ref = _ref_from(orc)
    ref.gen_heightmap()
    orc.gen_heightmap()
    _compare(ref, orc, f"heightmap init ({variant})", ("heightmap", "flux", "velocity", "sediment"))
    H = orc.get(0)
    assert H[..., 0].max() > H[..., 0].min() and (H[..., 1] > 0).all()
    orc.close()


@pytest.mark.parametrize("variant", ["default", "steep_fast", "thin_dirt"])
def test_multi_step_runs_match_the_reference_shaders(variant):
    """main-loop iterations (src/main.cpp:310-321) from the generated terrain: rain when due, then the 8 dispatches"""
    n = 64
    orc = oracle.World(n, seed=SEED)
    orc.gen_heightmap()
    orc.rain.period = 4
    if variant == "steep_fast":
        orc.erosion.Kalpha[0], orc.erosion.Kalpha[1] = 0.5, 0.2
        orc.erosion.d_t = 0.02
        orc.rain.amount = 0.5
    if variant == "thin_dirt":
        H = orc.get(0); H[..., 1] = 1e-4; H[..., 3] = H[..., 0] + H[..., 1] + H[..., 2]; orc.set(0, H)
        orc.erosion.Ks[1] = 50.0; orc.erosion.Kc = 5.0
        orc.rain.amount = 0.3
    ref = _ref_from(orc)
    thermal = 0.0
    for s in range(1, 25):
        t = float(np.float32(s) * np.float32(DT_TIME))
        ref.step(s, t)
        orc.step(t)
        _compare(ref, orc, f"{variant} step {s}")
        thermal = max(thermal, float(np.abs(orc.get(4)).max()), float(np.abs(orc.get(5)).max()))
    assert orc.get(0)[..., 2].max() > 0
    if variant == "steep_fast":
        assert thermal > 0          # thermal outflow (both neighbour classes) was live
    orc.close()


@pytest.mark.parametrize("hmap", [64, 48])
def test_droplet_mode_matches_the_reference_shaders(hmap):
    """Erosion::dispatch_particle (src/erosion.cpp:132-156): particle.glsl, particle_erosion.glsl (spin lock on
    the r32ui lock map), thermal x2, smoothing with the momentum map.  The shaders' invocations run in id order
    here, which is the order the oracle defines for contended texels; 1024 droplets on 48^2 / 64^2 cells collide
    constantly, so the lock path and the layer-exhaustion clamp are live.
    hmap_dims <= map size, as in every configuration the reference can reach (main.cpp:210 makes them equal;
    droplets die 2 cells inside hmap_dims).  With hmap_dims LARGER than the map (tried: the default (1024, 1024)
    on this 64^2 map) droplets walk off the map and deposit momentum on its border texels, which smoothing.glsl
    never rewrites (early return, :27-33): the reference then reads back whatever the other ping-pong texture
    held two steps earlier, while the oracle and the product define those texels as 0 (oracle/hg_oracle.c,
    smooth_pass) -- the two agree for 8 steps and then part; that configuration is outside the contract."""
    n, count = 64, 1024
    orc = oracle.World(n, particle_count=count, erosion_type=1, seed=SEED)
    orc.gen_heightmap()
    if hmap:
        orc.map.hmap_dims[0], orc.map.hmap_dims[1] = hmap, hmap
    e = oracle.ErosionData.from_buffer_copy(bytes(orc.erosion))
    ref = refshaders.RefWorld(n, e, oracle.RainData.from_buffer_copy(bytes(orc.rain)),
                              oracle.MapSettingsData.from_buffer_copy(bytes(orc.map)), particle_count=count)
    for k, f in enumerate(FIELDS):
        getattr(ref, f).read[...] = orc.get(k)
    for s in range(1, 13):
        t = float(np.float32(s) * np.float32(DT_TIME))
        rain = s < 9                                   # later steps: no respawn, droplets die out
        ref.dispatch_particle(t, rain)
        orc.dispatch_particle(t, rain)
        assert ref.particle_buffer[:count * 48].tobytes() == orc.particles().tobytes(), f"droplets differ after step {s}"
        _compare(ref, orc, f"droplet step {s}", ("heightmap", "velocity"))
    p = orc.particles()
    assert (p["iters"] > 0).any() and (p["to_kill"] != 0).any()
    orc.close()


# ---- the committed golden vectors ARE the reference shaders' outputs ------------------------------------
def _ref_for_case(cases, name):
    from tests.test_golden import apply_params
    Hm = cases[f"{name}/in/H"]
    assert Hm.shape[0] == Hm.shape[1]
    w = oracle.World(8)                              # only for the default settings blocks
    e = apply_params(oracle.ErosionData.from_buffer_copy(bytes(w.erosion)), cases[f"{name}/params"])
    ref = refshaders.RefWorld(Hm.shape[1], e, oracle.RainData.from_buffer_copy(bytes(w.rain)), oracle.MapSettingsData.from_buffer_copy(bytes(w.map)))
    w.close()
    return ref


@pytest.mark.parametrize("name", ["default_wet", "steep_thermal"])       # the square cases (RefWorld textures are n x n)
def test_committed_golden_vectors_are_the_reference_shaders_outputs(name):
    """tests/golden/grid_cases.npz was written by an independent numpy restatement before the reference
    shaders could be run here; the shaders reproduce every image in it bit for bit."""
    import os
    from tests.test_golden import PASS_CHECKS
    cases = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid_cases.npz"))
    ref = _ref_for_case(cases, name)
    for k, f in (("H", "heightmap"), ("F", "flux"), ("V", "velocity"), ("S", "sediment")):
        getattr(ref, f).read[...] = cases[f"{name}/in/{k}"]
    for (stage, _, fields), method in zip(PASS_CHECKS, refshaders.RefWorld.PASSES):
        getattr(ref, method)()
        for j, f in enumerate(fields):
            assert_bit_equal(getattr(ref, f).read, cases[f"{name}/step1/{stage}/{j}"], f"{name} after {stage}: {f}")
    ref.dispatch_grid(); ref.dispatch_grid()
    for k, f in (("H", "heightmap"), ("F", "flux"), ("V", "velocity"), ("S", "sediment")):
        assert_bit_equal(getattr(ref, f).read, cases[f"{name}/step3/out/{k}"], f"{name} step 3: {f}")


@pytest.mark.parametrize("seed", range(10))
def test_random_parameters_match_the_reference_shaders(seed):
    """randomised: small square maps (down to 8 x 8), random erosion / rain parameters and map seed, terrain from
    heightmap.glsl, six main-loop iterations: oracle == reference shaders after every iteration"""
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.choice((8, 16, 24, 40, 64)))
    orc = oracle.World(n, seed=float(rng.uniform(0.0, 5000.0)))
    e = orc.erosion
    e.Kc = float(rng.uniform(0.01, 2.0)); e.d_t = float(rng.uniform(0.001, 0.03)); e.G = float(rng.uniform(1.0, 20.0))
    e.Ke = float(rng.uniform(0.0, 0.5)); e.ENERGY_KEPT = float(rng.uniform(0.5, 1.0)); e.Kconv = float(rng.uniform(0.0, 0.1))
    for i in range(2):
        e.Kalpha[i] = float(rng.uniform(0.1, 1.2)); e.Ks[i] = float(rng.uniform(0.001, 2.0)); e.Kd[i] = float(rng.uniform(0.001, 2.0))
        e.Kspeed[i] = float(rng.uniform(0.1, 30.0))
    orc.rain.period = int(rng.integers(1, 4)); orc.rain.amount = float(rng.uniform(0.001, 1.0)); orc.rain.drops = float(rng.uniform(0.005, 0.3))
    ref = _ref_from(orc)
    ref.gen_heightmap(); orc.gen_heightmap()
    _compare(ref, orc, f"seed {seed}: init", ("heightmap",))
    for s in range(1, 7):
        t = float(np.float32(s) * np.float32(DT_TIME))
        ref.step(s, t)
        orc.step(t)
        _compare(ref, orc, f"seed {seed} n={n} step {s}")
    orc.close()
