"""The opt-in CONTRACTED build (HG_FMAD=1 -> libhydrogen_b200_fmad.so, the same sources with -fmad=true): ptxas may form
FMAs as a GLSL compiler may, so results are no longer bit-identical to the oracle; they must stay inside the contract the
north star states -- per-field max relative error <= 1e-5 after one step (SURVEY.md §8d gate (i): elementwise against
max(|ref|, 1e-6 * field max)), and after 1000 steps mass drift <= max(1e-5, 10 x the oracle's own FMA noise floor 8.6e-6)
and mean |d(rock+dirt)| <= max(1e-3, 10 x 1.3e-4) height units (gate (ii), tests/test_drift.py).  The default library
stays bit-exact; this build only exists to measure what bit-exactness costs (bench.py key "contracted_build")."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_contracted_library_is_built_and_says_so(built):
    import ctypes as C
    path = os.path.join(ROOT, "hydro_gen_b200", "libhydrogen_b200_fmad.so")
    assert os.path.exists(path), "make -C hydro_gen_b200/csrc builds both libraries"
    L = C.CDLL(path)
    L.hg_version.restype = C.c_char_p
    assert b"contracted" in L.hg_version()
    assert b"contracted" not in built.hg_version()


@pytest.mark.gpu
def test_contracted_build_within_north_star_tolerance(built):
    env = dict(os.environ, HG_FMAD="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fmad_check.py")], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert "contracted" in out["version"]
    for name, err in out["one_step"].items():
        assert err <= 1e-5, f"{name}: max relative error {err} after one step"
    assert out["drift_1000"]["mass_rel"] <= 8.6e-5
    assert out["drift_1000"]["mean_abs_terrain"] <= 1.3e-3
    print(out)
