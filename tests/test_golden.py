"""Committed known-answer vectors (tests/golden/, made by tests/golden/make_golden.py).

grid_cases.npz comes from an INDEPENDENT numpy restatement of the shaders
(tests/golden/grid_restatement.py); oracle_runs.npz are seed-fixed oracle dumps.
CPU tests: the oracle reproduces both files bit for bit (this is what pins the oracle).
GPU tests: the CUDA path (PASSES after every dispatch, FUSED after 1 and 3 steps)
reproduces the same files without the oracle being involved at all.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from tests.util import FIELDS, assert_bit_equal

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ("default_wet", "big_dt_fast_water", "steep_thermal", "thin_dirt_dry_patches")
PNAMES = ("Kc", "Kalpha0", "Kalpha1", "Kconv", "Ks0", "Ks1", "Kd0", "Kd1", "Ke", "ENERGY_KEPT", "Kspeed0", "Kspeed1", "G", "d_t")


@pytest.fixture(scope="module")
def cases():
    return np.load(os.path.join(HERE, "golden", "grid_cases.npz"))


@pytest.fixture(scope="module")
def runs():
    return np.load(os.path.join(HERE, "golden", "oracle_runs.npz"))


def apply_params(e, p):
    """p: the 14 floats stored per case, in PNAMES order, into an Erosion_data struct."""
    v = dict(zip(PNAMES, (float(x) for x in p)))
    e.Kc, e.Kconv, e.Ke, e.ENERGY_KEPT, e.G, e.d_t = v["Kc"], v["Kconv"], v["Ke"], v["ENERGY_KEPT"], v["G"], v["d_t"]
    e.Kalpha[0], e.Kalpha[1] = v["Kalpha0"], v["Kalpha1"]
    e.Ks[0], e.Ks[1] = v["Ks0"], v["Ks1"]
    e.Kd[0], e.Kd[1] = v["Kd0"], v["Kd1"]
    e.Kspeed[0], e.Kspeed[1] = v["Kspeed0"], v["Kspeed1"]
    return e


def load_oracle(cases, name):
    Hm = cases[f"{name}/in/H"]
    w = oracle.World(Hm.shape[1], Hm.shape[0])
    apply_params(w.erosion, cases[f"{name}/params"])
    for k, f in (("H", "heightmap"), ("F", "flux"), ("V", "velocity"), ("S", "sediment")):
        w.set(FIELDS[f], cases[f"{name}/in/{k}"])
    return w


PASS_CHECKS = (("flux", oracle.PASS_FLUX, ("heightmap", "flux", "velocity")),
               ("erosion", oracle.PASS_EROSION, ("heightmap", "sediment")),
               ("sediment", oracle.PASS_SEDIMENT, ("heightmap", "sediment")),
               ("thermal1", oracle.PASS_THERMAL, ("heightmap", "thermal_c", "thermal_d")),
               ("smooth", oracle.PASS_SMOOTH, ("heightmap",)))


# ------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_independent_restatement_per_pass(cases, name):
    w = load_oracle(cases, name)
    for stage, p, fields in PASS_CHECKS:
        w.run_pass(p)
        for j, f in enumerate(fields):
            assert_bit_equal(w.get(FIELDS[f]), cases[f"{name}/step1/{stage}/{j}"], f"{name} after {stage}: {f}")
    w.close()


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_independent_restatement_3_steps(cases, name):
    w = load_oracle(cases, name)
    w.dispatch_grid()
    for k, f in (("H", "heightmap"), ("F", "flux"), ("V", "velocity"), ("S", "sediment"), ("TC", "thermal_c"), ("TD", "thermal_d")):
        assert_bit_equal(w.get(FIELDS[f]), cases[f"{name}/step1/out/{k}"], f"{name} step 1: {f}")
    w.dispatch_grid(); w.dispatch_grid()
    for k, f in (("H", "heightmap"), ("F", "flux"), ("V", "velocity"), ("S", "sediment")):
        assert_bit_equal(w.get(FIELDS[f]), cases[f"{name}/step3/out/{k}"], f"{name} step 3: {f}")
    w.close()


def test_golden_cases_hit_the_branches(cases):
    """the fixtures are only worth something if the hard branches are live in them"""
    c = cases
    assert (c["thin_dirt_dry_patches/in/H"][..., 2] == 0).sum() > 50                      # dry cells: K = min(1, 0/0)
    assert (c["thin_dirt_dry_patches/step1/erosion/0"][..., 1] == 0).sum() > (c["thin_dirt_dry_patches/in/H"][..., 1] == 0).sum()   # dirt exhausted
    assert (c["steep_thermal/step1/thermal0/1"] > 0).sum() > 100                          # rock-layer marks
    assert (c["steep_thermal/step1/thermal1/2"] > 0).sum() > 100                          # diagonal outflow, dirt layer
    v = c["big_dt_fast_water/step1/flux/2"]
    assert (np.abs(v[..., :2]) * 0.02 > 2.0).sum() > 20                                   # back-traces of > 2 cells
    h0, h1 = c["default_wet/step1/thermal1/0"], c["default_wet/step1/smooth/0"]
    assert (h0[1:-1, 1:-1, :2] != h1[1:-1, 1:-1, :2]).any()                               # smoothing changed terrain


def test_oracle_reproduces_committed_runs(runs):
    w = oracle.World(64, seed=1234.5)
    w.gen_heightmap()
    assert_bit_equal(w.get(0), runs["init64/H"], "init 64")
    w.rain.period = 8
    for s in range(1, 41):
        w.step(s * 0.015)
    for k, fid in (("H", 0), ("F", 1), ("S", 3)):
        assert_bit_equal(w.get(fid), runs[f"run64_40/{k}"], f"run64_40 {k}")
    w.close()


# ------------------------------------------------------------------ GPU: the CUDA path
def _gpu_ctx(cases, name, schedule):
    from hydro_gen_b200 import Context
    Hm = cases[f"{name}/in/H"]
    ctx = Context(Hm.shape[1], Hm.shape[0])
    ctx.set_schedule(schedule)
    ctx.set_erosion(apply_params(ctx.get_erosion(), cases[f"{name}/params"]))
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_passes_match_golden(built, cases, name):
    from hydro_gen_b200 import _lib
    ctx = _gpu_ctx(cases, name, _lib.SCHEDULE_PASSES)
    for k, f in (("H", "heightmap"), ("F", "flux"), ("V", "velocity"), ("S", "sediment")):
        ctx.upload(FIELDS[f], cases[f"{name}/in/{k}"])
    for stage, p, fields in PASS_CHECKS:
        ctx.dispatch_pass(p)
        for j, f in enumerate(fields):
            assert_bit_equal(ctx.download(FIELDS[f]), cases[f"{name}/step1/{stage}/{j}"], f"{name} after {stage}: {f}")
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_fused_matches_golden(built, cases, name):
    from hydro_gen_b200 import _lib
    ctx = _gpu_ctx(cases, name, _lib.SCHEDULE_FUSED)
    for k, f in (("H", "heightmap"), ("F", "flux"), ("S", "sediment")):
        ctx.upload(FIELDS[f], cases[f"{name}/in/{k}"])
    ctx.dispatch_grid()
    for k, f in (("H", "heightmap"), ("F", "flux"), ("S", "sediment")):
        assert_bit_equal(ctx.download(FIELDS[f]), cases[f"{name}/step1/out/{k}"], f"{name} fused step 1: {f}")
    ctx.dispatch_grid(); ctx.dispatch_grid()
    for k, f in (("H", "heightmap"), ("F", "flux"), ("S", "sediment")):
        assert_bit_equal(ctx.download(FIELDS[f]), cases[f"{name}/step3/out/{k}"], f"{name} fused step 3: {f}")
    ctx.close()


@pytest.mark.gpu
def test_cuda_reproduces_committed_runs(built, runs):
    from hydro_gen_b200 import Context
    for tag, shape, seed, period, steps in (("64", (64, 64), 1234.5, 8, 40), ("48x80", (48, 80), 77.25, 4, 24)):
        ctx = Context(shape[0], shape[1])
        m = ctx.get_map(); m.seed = seed; ctx.set_map(m)
        r = ctx.get_rain(); r.period = period; ctx.set_rain(r)
        ctx.gen_heightmap()
        assert_bit_equal(ctx.download(0), runs[f"init{tag}/H"], f"init {tag}")
        for s in range(1, steps + 1):          # time = s * 0.015 in double, as the generator passes it
            ctx.run(1, s * 0.015, 0.0, True)
        for k, fid in (("H", 0), ("F", 1), ("S", 3)):
            assert_bit_equal(ctx.download(fid), runs[f"run{tag}_{steps}/{k}"], f"run{tag} {k}")
        ctx.close()
