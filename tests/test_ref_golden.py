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
