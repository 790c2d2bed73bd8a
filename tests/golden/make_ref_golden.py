#!/usr/bin/env python
"""Generates tests/golden/refshader_runs.npz: outputs of THE REFERENCE ITSELF, run here.  TEST INFRASTRUCTURE.

  python tests/golden/make_ref_golden.py        (needs /root/reference/glsl; builds oracle/_ref/libhg_refshaders.so)

The reference's own compute shaders, compiled for the CPU (oracle/refshader/build_ref.py) and driven like
src/main.cpp:310-321 / src/erosion.cpp:76-200 (oracle/refshaders.py), advance seeded 48x48 states; the inputs
and the heightmap / flux / sediment images after 1, 8 and 16 main-loop iterations are stored, plus the terrain
heightmap.glsl generates at 64x64 and a 5-step sparse droplet run (heightmap, momentum map, droplet buffer).  The oracle
(CPU test) and the CUDA path (GPU test, where /root/reference does not exist) must reproduce them bit for bit
(tests/test_refshaders.py).  Initial terrains come from the oracle's heightmap generator: they are inputs here,
not something under test."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle                       # noqa: E402
from oracle import refshaders       # noqa: E402

N, PERIOD, DT_TIME, SEED = 48, 4, 0.015, 1234.5
CHECKPOINTS = (1, 8, 16)
PNAMES = ("Kc", "Kalpha0", "Kalpha1", "Kconv", "Ks0", "Ks1", "Kd0", "Kd1", "Ke", "ENERGY_KEPT", "Kspeed0", "Kspeed1", "G", "d_t")
CASES = {"default": {}, "steep_fast": {"Kalpha0": 0.5, "Kalpha1": 0.2, "d_t": 0.02, "rain_amount": 0.5},
         "thin_dirt": {"Ks1": 50.0, "Kc": 5.0, "rain_amount": 0.3, "dirt": 1e-4}}


def params_of(e):
    return np.array([e.Kc, e.Kalpha[0], e.Kalpha[1], e.Kconv, e.Ks[0], e.Ks[1], e.Kd[0], e.Kd[1], e.Ke, e.ENERGY_KEPT,
                     e.Kspeed[0], e.Kspeed[1], e.G, e.d_t], np.float32)


def time_of(s):
    return float(np.float32(s) * np.float32(DT_TIME))


def main():
    if not refshaders.available():
        raise SystemExit("the reference shaders are not available")
    out = {}
    for name, ov in CASES.items():
        w = oracle.World(N, seed=SEED)
        w.gen_heightmap()
        e = w.erosion
        for k, v in ov.items():
            if k in ("Kalpha0", "Kalpha1"): e.Kalpha[int(k[-1])] = v
            elif k == "Ks1": e.Ks[1] = v
            elif k in ("Kc", "d_t"): setattr(e, k, v)
        w.rain.period = PERIOD
        if "rain_amount" in ov: w.rain.amount = ov["rain_amount"]
        H = w.get(0).copy()
        if "dirt" in ov:
            H[..., 1] = ov["dirt"]; H[..., 3] = H[..., 0] + H[..., 1] + H[..., 2]
        ref = refshaders.RefWorld(N, oracle.ErosionData.from_buffer_copy(bytes(e)), oracle.RainData.from_buffer_copy(bytes(w.rain)),
                                  oracle.MapSettingsData.from_buffer_copy(bytes(w.map)))
        ref.heightmap.read[...] = H
        out[f"{name}/params"] = params_of(e)
        out[f"{name}/rain"] = np.array([w.rain.amount, w.rain.mountain_thresh, w.rain.mountain_multip, w.rain.period, w.rain.drops], np.float32)
        out[f"{name}/max_height"] = np.float32(w.map.max_height)
        out[f"{name}/in/H"] = H
        for s in range(1, max(CHECKPOINTS) + 1):
            ref.step(s, time_of(s))
            if s in CHECKPOINTS:
                for k, f in (("H", "heightmap"), ("F", "flux"), ("S", "sediment")):
                    out[f"{name}/step{s}/{k}"] = getattr(ref, f).read.copy()
        w.close()
    # heightmap.glsl: the generated terrain itself (default map settings)
    w = oracle.World(64, seed=SEED)
    ref = refshaders.RefWorld(64, oracle.ErosionData.from_buffer_copy(bytes(w.erosion)), oracle.RainData.from_buffer_copy(bytes(w.rain)),
                              oracle.MapSettingsData.from_buffer_copy(bytes(w.map)))
    ref.gen_heightmap()
    out["init64/H"] = ref.heightmap.read.copy()
    w.close()
    # droplet mode, sparse: 64 droplets (one work group) on 128^2, with a spawn time for which no two droplets ever
    # touch the same texel in these steps, so the result does not depend on lock order and the GPU path (atomics,
    # any order) must reproduce it exactly.  The time offset is searched; collisions are detected from the quads.
    n, count = 128, 64
    part_dt = np.dtype([("sc", "<f4"), ("iters", "<i4"), ("position", "<f4", 2), ("velocity", "<f4", 2), ("volume", "<f4"),
                        ("_p0", "<u4"), ("sediment", "<f4", 2), ("to_kill", "<u4"), ("_p1", "<u4")])
    for t0 in range(0, 200):
        w = oracle.World(n, particle_count=count, erosion_type=1, seed=SEED)
        w.gen_heightmap()
        w.map.hmap_dims[0], w.map.hmap_dims[1] = n, n
        ref = refshaders.RefWorld(n, oracle.ErosionData.from_buffer_copy(bytes(w.erosion)), oracle.RainData.from_buffer_copy(bytes(w.rain)),
                                  oracle.MapSettingsData.from_buffer_copy(bytes(w.map)), particle_count=count)
        H0 = w.get(0)
        ref.heightmap.read[...] = H0
        w.close()
        clean = True
        for s in range(1, 6):
            ref.dispatch_particle(time_of(t0 + s), True)
            p = np.frombuffer(ref.particle_buffer[:count * 48].tobytes(), dtype=part_dt)
            live = p["iters"] > 0
            base = np.floor(p["position"][live]).astype(np.int64)
            quads = np.concatenate([base + np.array(o) for o in ((0, 0), (1, 0), (1, 1), (0, 1))])
            if len({(int(a), int(b)) for a, b in quads}) != len(quads):
                clean = False
                break
        if clean:
            break
    else:
        raise SystemExit("no collision-free droplet case found")
    out["drops/t0"] = np.int32(t0)
    out["drops/in/H"] = H0
    out["drops/step5/H"] = ref.heightmap.read.copy()
    out["drops/step5/M"] = ref.velocity.read.copy()
    out["drops/step5/particles"] = ref.particle_buffer[:count * 48].copy()
    print("droplet case: time offset", t0)
    path = os.path.join(HERE, "refshader_runs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
