"""Randomised parity: small maps of odd shapes (down to 8 x 8, narrower than one strip, shorter than the pipeline is
deep), random erosion / rain parameters, a few main-loop iterations through the C ABI against the oracle, bit for bit.
Seeds are fixed: the cases are the same on every run."""
import numpy as np
import pytest

import oracle
from hydro_gen_b200 import Context, _lib
from tests.util import DT_TIME, assert_bit_equal

pytestmark = pytest.mark.gpu
SIZES = (8, 16, 24, 40, 64, 72, 112, 120, 128, 136, 232, 248, 264)


@pytest.mark.parametrize("seed", range(16))
def test_random_shapes_and_parameters(built, seed):
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.choice(SIZES)), int(rng.choice(SIZES))
    ref = oracle.World(W, H, seed=float(rng.uniform(0.0, 5000.0)))
    e = ref.erosion
    e.Kc = float(rng.uniform(0.01, 2.0)); e.d_t = float(rng.uniform(0.001, 0.03)); e.G = float(rng.uniform(1.0, 20.0))
    e.Ke = float(rng.uniform(0.0, 0.5)); e.ENERGY_KEPT = float(rng.uniform(0.5, 1.0)); e.Kconv = float(rng.uniform(0.0, 0.1))
    for i in range(2):
        e.Kalpha[i] = float(rng.uniform(0.1, 1.2)); e.Ks[i] = float(rng.uniform(0.001, 2.0)); e.Kd[i] = float(rng.uniform(0.001, 2.0))
        e.Kspeed[i] = float(rng.uniform(0.1, 30.0))
    ref.rain.period = int(rng.integers(1, 4)); ref.rain.amount = float(rng.uniform(0.001, 1.0)); ref.rain.drops = float(rng.uniform(0.005, 0.3))
    ref.gen_heightmap()
    ctx = Context(W, H)
    ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(ref.map)))
    ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(e)))
    ctx.set_rain(_lib.RainData.from_buffer_copy(bytes(ref.rain)))
    ctx.gen_heightmap()
    assert_bit_equal(ctx.download(0), ref.get(0), f"seed {seed} {W}x{H}: init")
    for s in range(1, 7):
        t = float(np.float32(s) * np.float32(DT_TIME))
        ref.step(t)
        ctx.run(1, t, 0.0, True)
    for fid, name in ((0, "heightmap"), (1, "flux"), (3, "sediment")):
        assert_bit_equal(ctx.download(fid), ref.get(fid), f"seed {seed} {W}x{H}: {name}")
    ctx.close(); ref.close()


@pytest.mark.parametrize("variant,shape", [(10, (264, 200)), (13, (264, 200)), (15, (136, 72)), (18, (264, 200)), (18, (120, 40)), (19, (400, 136)),
                                           (21, (264, 200)), (22, (264, 200)), (22, (400, 72)), (24, (400, 136)), (25, (264, 200)), (26, (136, 72))])
def test_tuning_variants_of_the_fused_kernel_are_bit_exact(built, monkeypatch, variant, shape):
    """The non-default forms of the fused step kernel kept in the tree as measured variants (HG_FUSED_VARIANT, read when a
    context is created): two columns per thread (10), three warp groups (13, 15), queued thermal outflow with one / two
    service warps (18, 19; 24 on 192-column strips), warp groups swapped (21) or interleaved in pairs of warps (25; 26 is its in-order control), 192-column strips (22) -- each against the
    oracle over 10 main-loop steps with rain, bit for bit, on maps that are not a multiple of the strip width."""
    W, H = shape
    monkeypatch.setenv("HG_FUSED_VARIANT", str(variant))
    ref = oracle.World(W, H, seed=77.25)
    ref.rain.period = 3
    ref.gen_heightmap()
    ctx = Context(W, H)
    ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(ref.map)))
    ctx.set_rain(_lib.RainData.from_buffer_copy(bytes(ref.rain)))
    ctx.gen_heightmap()
    for s in range(1, 11):
        t = float(np.float32(s) * np.float32(DT_TIME))
        ref.step(t)
        ctx.run(1, t, 0.0, True)
    for fid, name in ((0, "heightmap"), (1, "flux"), (3, "sediment")):
        assert_bit_equal(ctx.download(fid), ref.get(fid), f"variant {variant} {W}x{H}: {name}")
    ctx.close(); ref.close()
