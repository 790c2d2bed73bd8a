"""Shared helpers for the parity tests: oracle states and bitwise comparison."""
import numpy as np

import oracle

SEED = 1234.5
DT_TIME = 0.015
FIELDS = {"heightmap": 0, "flux": 1, "velocity": 2, "sediment": 3, "thermal_c": 4, "thermal_d": 5}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(got, want, what):
    g, w = bits(got), bits(want)
    if not np.array_equal(g, w):
        bad = g != w
        idx = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.size} values differ; first at {tuple(idx)}: "
                             f"got {got[tuple(idx)]!r} want {want[tuple(idx)]!r}; max abs diff "
                             f"{np.nanmax(np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64)))}")


def max_rel_err(got, want):
    """SURVEY.md §8d gate (i): max |d| / max(|ref|, 1e-6 * field max), elementwise."""
    want = np.asarray(want, np.float64)
    got = np.asarray(got, np.float64)
    eps = 1e-6 * max(np.abs(want).max(), 1e-30)
    return float((np.abs(got - want) / np.maximum(np.abs(want), eps)).max())


def wet_world(n, steps, period=16, width=None, seed=SEED, fma=False):
    """Oracle world after `steps` main-loop iterations with rain every `period` steps, so water,
    flux, velocity and sediment are all live (BASELINE config 1b)."""
    w = oracle.World(width or n, n, seed=seed, fma=fma)
    w.gen_heightmap()
    w.rain.period = period
    for s in range(1, steps + 1):
        w.step(s * DT_TIME)
    return w


def copy_state(src, dst_ctx, fields=("heightmap", "flux", "sediment")):
    for f in fields:
        dst_ctx.upload(FIELDS[f], src.get(FIELDS[f]))
