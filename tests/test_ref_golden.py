"""tests/golden/refshader_runs.npz holds outputs of THE REFERENCE ITSELF: its own compute shaders compiled for
the CPU and driven like its main loop (tests/golden/make_ref_golden.py, run where /root/reference exists).
These tests need neither the reference nor the compiled shaders: the oracle (CPU) and the CUDA path through the
C ABI (GPU) must reproduce the committed images bit for bit after 1, 8 and 16 main-loop iterations with rain
every 4 steps, on three seeded 48x48 cases (default parameters; steep talus + fast water; exhausted dirt)."""
import os

import numpy as np
import pytest

import oracle
from tests.test_golden import apply_params
from tests.util import assert_bit_equal

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ("default", "steep_fast", "thin_dirt")
CHECKPOINTS = (1, 8, 16)
DT_TIME = 0.015


@pytest.fixture(scope="module")
def runs():
    return np.load(os.path.join(HERE, "golden", "refshader_runs.npz"))


def _time(s):
    return float(np.float32(s) * np.float32(DT_TIME))


def _apply_rain(r, v):
    r.amount, r.mountain_thresh, r.mountain_multip, r.period, r.drops = float(v[0]), float(v[1]), float(v[2]), int(v[3]), float(v[4])
    return r


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_shader_runs(runs, name):
    H = runs[f"{name}/in/H"]
    w = oracle.World(H.shape[1])
    apply_params(w.erosion, runs[f"{name}/params"])
    _apply_rain(w.rain, runs[f"{name}/rain"])
    assert w.map.max_height == runs[f"{name}/max_height"]
    w.set(0, H)
    for s in range(1, max(CHECKPOINTS) + 1):
        w.step(_time(s))
        if s in CHECKPOINTS:
            for k, fid in (("H", 0), ("F", 1), ("S", 3)):
                assert_bit_equal(w.get(fid), runs[f"{name}/step{s}/{k}"], f"{name} step {s}: {k}")
    assert w.get(0)[..., 2].max() > 0
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_reference_shader_runs(built, runs, name):
    from hydro_gen_b200 import Context
    H = runs[f"{name}/in/H"]
    ctx = Context(H.shape[1])
    ctx.set_erosion(apply_params(ctx.get_erosion(), runs[f"{name}/params"]))
    ctx.set_rain(_apply_rain(ctx.get_rain(), runs[f"{name}/rain"]))
    assert ctx.get_map().max_height == runs[f"{name}/max_height"]
    ctx.upload(0, H)
    ctx.upload(1, np.zeros_like(H)); ctx.upload(3, np.zeros_like(H))
    for s in range(1, max(CHECKPOINTS) + 1):
        ctx.run(1, _time(s), 0.0, True)
        if s in CHECKPOINTS:
            for k, fid in (("H", 0), ("F", 1), ("S", 3)):
                assert_bit_equal(ctx.download(fid), runs[f"{name}/step{s}/{k}"], f"{name} step {s}: {k}")
    ctx.close()


def test_oracle_reproduces_reference_heightmap_and_droplets(runs):
    w = oracle.World(64, seed=1234.5)
    w.gen_heightmap()
    assert_bit_equal(w.get(0), runs["init64/H"], "heightmap.glsl at 64x64")
    w.close()
    n, count = 128, 64
    w = oracle.World(n, particle_count=count, erosion_type=1, seed=1234.5)
    w.set(0, runs["drops/in/H"])
    w.map.hmap_dims[0], w.map.hmap_dims[1] = n, n
    for s in range(1, 6):
        w.dispatch_particle(_time(int(runs["drops/t0"]) + s), True)
    assert w.particles().tobytes() == runs["drops/step5/particles"].tobytes()
    assert_bit_equal(w.get(0), runs["drops/step5/H"], "droplets: heightmap")
    assert_bit_equal(w.get(2), runs["drops/step5/M"], "droplets: momentum map")
    w.close()


@pytest.mark.gpu
def test_cuda_reproduces_reference_heightmap_and_droplets(built, runs):
    from hydro_gen_b200 import Context, _lib
    ctx = Context(64)
    m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
    ctx.gen_heightmap()
    assert_bit_equal(ctx.download(0), runs["init64/H"], "heightmap.glsl at 64x64")
    ctx.close()
    n, count = 128, 64
    ctx = Context(n, particle_count=count, erosion_type=_lib.HG_PARTICLES)
    m = ctx.get_map(); m.seed = 1234.5; m.hmap_dims[0], m.hmap_dims[1] = n, n; ctx.set_map(m)
    ctx.upload(0, runs["drops/in/H"])
    ctx.upload(2, np.zeros_like(runs["drops/in/H"]))
    for s in range(1, 6):      # a spawn time for which no two droplets share a texel: the result is order-free
        ctx.dispatch_particle(_time(int(runs["drops/t0"]) + s), True)
    assert np.asarray(ctx.download_particles()).tobytes() == runs["drops/step5/particles"].tobytes()
    assert_bit_equal(ctx.download(0), runs["drops/step5/H"], "droplets: heightmap")
    assert_bit_equal(ctx.download(2), runs["drops/step5/M"], "droplets: momentum map")
    ctx.close()
