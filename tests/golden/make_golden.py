#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/.  TEST INFRASTRUCTURE.

  python tests/golden/make_golden.py            # rewrites grid_cases.npz and oracle_runs.npz

grid_cases.npz   known-answer vectors from the independent numpy restatement
                 (grid_restatement.py): seeded 16x16 / 24x16 states built to hit every branch
                 (dry cells with K = NaN, exhausted dirt layer, thermal marks in both layers,
                 smoothing extrema, back-traces that leave the 3x3 neighbourhood and hit the
                 map edge), the images after each of the 8 dispatches of one step and after
                 3 whole steps.
oracle_runs.npz  seed-fixed oracle dumps (64x64: heightmap init; 40 main-loop iterations with
                 rain every 8 steps; 48x80 non-square) so the GPU tests can also run against
                 committed files, and so a change of the oracle itself is caught.

The reference has no fixtures of its own.  These files were the first pin of the
oracle; tests/golden/make_ref_golden.py adds outputs of the reference's own shaders compiled for
the CPU, which also reproduce grid_cases.npz bit for bit.  Neither generator here touches
/root/reference at run time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import grid_restatement as gr   # noqa: E402

f32 = np.float32

CASES = {
    # name: (W, H, rng seed, Params overrides, state style)
    "default_wet": (16, 16, 1, {}, "wet"),
    "big_dt_fast_water": (24, 16, 2, {"d_t": 0.02, "G": 9.81, "Ke": 0.2}, "torrent"),
    "steep_thermal": (16, 16, 3, {"Kalpha": (0.35, 0.2), "Kspeed": (5.0, 20.0), "d_t": 0.01}, "rough"),
    "thin_dirt_dry_patches": (16, 24, 4, {"d_t": 0.05, "Ks": (0.3, 0.5), "Kc": 1.0}, "thin"),
}


def make_state(W, H, seed, style):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    rock = 20.0 + 6.0 * np.sin(xx * 0.7) * np.cos(yy * 0.5) + rng.uniform(0, 1.5, (H, W))
    dirt = rng.uniform(0.2, 1.0, (H, W))
    water = rng.uniform(0.0, 0.6, (H, W))
    flux = rng.uniform(0.0, 2.0, (H, W, 4))
    sed = rng.uniform(0.0, 0.05, (H, W, 2))
    if style == "torrent":
        water = rng.uniform(0.0, 0.02, (H, W))          # shallow, fast: |v dt| of several cells
        flux = rng.uniform(0.0, 30.0, (H, W, 4))
        flux[rng.random((H, W, 4)) < 0.3] = 0.0
        sed = rng.uniform(0.0, 0.5, (H, W, 2))
    elif style == "rough":
        rock = 30.0 + rng.uniform(0, 12.0, (H, W))       # slopes far above both talus angles
        rock[4:7, 5:9] += 25.0
        dirt = rng.uniform(0.0, 3.0, (H, W))
        dirt[rng.random((H, W)) < 0.2] = 0.0
    elif style == "thin":
        dirt = rng.uniform(0.0, 2e-4, (H, W))            # erosion exhausts the dirt layer
        dirt[rng.random((H, W)) < 0.3] = 0.0
        water = rng.uniform(0.0, 1.0, (H, W))
        water[rng.random((H, W)) < 0.35] = 0.0          # dry cells: K = min(1, 0/0)
        flux = rng.uniform(0.0, 8.0, (H, W, 4))
        flux[water == 0.0] = 0.0
    Hm = np.zeros((H, W, 4), f32)
    Hm[..., 0], Hm[..., 1], Hm[..., 2] = rock, dirt, water
    Hm[..., 3] = Hm[..., 0] + Hm[..., 1] + Hm[..., 2]    # H.a as smoothing/rain leave it: (r+g)+b
    F = flux.astype(f32)
    F[:, 0, 0] = 0; F[:, -1, 1] = 0; F[-1, :, 2] = 0; F[0, :, 3] = 0   # what the flux pass itself guarantees
    V = np.zeros((H, W, 4), f32)
    V[..., :3] = rng.uniform(-1, 1, (H, W, 3))           # overwritten by the flux pass; w passes through
    S = np.zeros((H, W, 4), f32)
    S[..., :2] = sed
    return Hm, F, V, S


def build_grid_cases():
    out = {}
    for name, (W, H, seed, over, style) in CASES.items():
        P = gr.Params(**over)
        Hm, F, V, S = make_state(W, H, seed, style)
        out[f"{name}/params"] = np.array([P.Kc, P.Kalpha[0], P.Kalpha[1], P.Kconv, P.Ks[0], P.Ks[1], P.Kd[0], P.Kd[1],
                                          P.Ke, P.ENERGY_KEPT, P.Kspeed[0], P.Kspeed[1], P.G, P.d_t], f32)
        for k, a in zip("HFVS", (Hm, F, V, S)):
            out[f"{name}/in/{k}"] = a
        trace = {}
        h, f_, v, s, tc, td = gr.grid_step(Hm, F, V, S, P, trace)
        for stage, imgs in trace.items():
            for j, a in enumerate(imgs):
                out[f"{name}/step1/{stage}/{j}"] = a
        for k, a in zip(("H", "F", "V", "S", "TC", "TD"), (h, f_, v, s, tc, td)):
            out[f"{name}/step1/out/{k}"] = a
        for _ in range(2):
            h, f_, v, s, tc, td = gr.grid_step(h, f_, v, s, P)
        for k, a in zip(("H", "F", "V", "S"), (h, f_, v, s)):
            out[f"{name}/step3/out/{k}"] = a
        print(f"{name}: {W}x{H} done; water sum {float(h[..., 2].sum()):.6f}")
    return out


def check_atan():
    xs = np.concatenate([np.linspace(0, 50, 20001), np.logspace(-6, 6, 4001)]).astype(f32)
    got = np.array([gr.atan_defined(x) for x in xs], f32)
    ref = np.arctan(xs.astype(np.float64))
    ulp = np.abs(got.astype(np.float64) - ref) / np.spacing(np.maximum(np.abs(ref), 1e-30).astype(f32)).astype(np.float64)
    assert ulp.max() <= 4.0, ulp.max()
    print(f"atan_defined vs arctan: max {ulp.max():.2f} ulp over {xs.size} points")


def build_oracle_runs():
    import oracle
    out = {}
    w = oracle.World(64, seed=1234.5)
    w.gen_heightmap()
    out["init64/H"] = w.get(0)
    w.rain.period = 8
    for s in range(1, 41):
        w.step(s * 0.015)
    for k, fid in (("H", 0), ("F", 1), ("S", 3)):
        out[f"run64_40/{k}"] = w.get(fid)
    w.close()
    w = oracle.World(48, 80, seed=77.25)
    w.gen_heightmap()
    out["init48x80/H"] = w.get(0)
    w.rain.period = 4
    for s in range(1, 25):
        w.step(s * 0.015)
    for k, fid in (("H", 0), ("F", 1), ("S", 3)):
        out[f"run48x80_24/{k}"] = w.get(fid)
    w.close()
    return out


if __name__ == "__main__":
    check_atan()
    np.savez_compressed(os.path.join(HERE, "grid_cases.npz"), **build_grid_cases())
    np.savez_compressed(os.path.join(HERE, "oracle_runs.npz"), **build_oracle_runs())
    for n in ("grid_cases.npz", "oracle_runs.npz"):
        print(n, os.path.getsize(os.path.join(HERE, n)), "bytes")
