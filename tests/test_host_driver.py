"""The C++ host side (host/hydrogen_erosion.hpp + host/hydro_gen_headless.cpp): builds with plain
g++, handles the reference's config.ini keys (src/main.cpp:203-234), fails loudly without a GPU,
and on a GPU gives the same fields as the Python mirror driving the same C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "hydro-gen-headless")


@pytest.fixture(scope="module")
def exe(built):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "host")], check=True)
    return EXE


def test_default_config_is_written_like_the_reference(exe, tmp_path):
    import torch
    r = subprocess.run([exe, "--steps", "1"], cwd=tmp_path, capture_output=True, text=True)
    text = (tmp_path / "config.ini").read_text()
    assert "[map]\nsize=1024" in text and "particle_count = 262144" in text and "type = grid" in text
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_python_config_reader_agrees(tmp_path):
    from hydro_gen_b200 import config
    p = tmp_path / "config.ini"
    p.write_text("[window]\nwidth = 800\nheight=600\n[map]\nsize = 2048 ; comment\n[erosion]\ntype = particle\nparticle_count = 4096\n")
    c = config.load(str(p))
    assert c == {"window_w": 800, "window_h": 600, "map_size": 2048, "erosion_type": "particle", "particle_count": 4096}
    c = config.load(str(tmp_path / "missing.ini"))
    assert c["map_size"] == 1024 and c["erosion_type"] == "grid" and c["particle_count"] == 0


@pytest.mark.gpu
def test_headless_driver_matches_python_mirror(exe, tmp_path):
    from hydro_gen_b200 import Context
    (tmp_path / "config.ini").write_text("[map]\nsize = 256\n[erosion]\ntype = grid\n")
    r = subprocess.run([exe, "--steps", "40", "--seed", "1234.5", "--rain-period", "8", "--dump", "out"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert re.search(r"grid 256x256, 40 steps", r.stdout)
    ctx = Context(256)
    m = ctx.get_map(); m.seed = 1234.5; ctx.set_map(m)
    rn = ctx.get_rain(); rn.period = 8; ctx.set_rain(rn)
    ctx.gen_heightmap()
    for k in range(40):
        ctx.run(1, np.float32(k + 1) * np.float32(0.015), 0.0, True)
    for name, field in (("heightmap", 0), ("sediment", 3)):
        got = np.fromfile(tmp_path / f"out.{name}.rgba32f", dtype=np.float32).reshape(256, 256, 4)
        assert np.array_equal(got.view(np.uint32), ctx.download(field).view(np.uint32)), name
    ctx.close()


@pytest.mark.gpu
def test_headless_driver_resume_from_checkpoint(exe, tmp_path):
    """40 steps in one go == 25 steps + --checkpoint, then --resume + 15 steps (fields bit for bit)."""
    (tmp_path / "config.ini").write_text("[map]\nsize = 128\n[erosion]\ntype = grid\n")
    common = ["--seed", "1234.5", "--rain-period", "8"]
    r = subprocess.run([exe, "--steps", "40", "--dump", "full"] + common, cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "--steps", "25", "--checkpoint", "mid.hgck"] + common, cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "--steps", "15", "--resume", "mid.hgck", "--dump", "resumed"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for name in ("heightmap", "sediment"):
        a = np.fromfile(tmp_path / f"full.{name}.rgba32f", dtype=np.uint32)
        b = np.fromfile(tmp_path / f"resumed.{name}.rgba32f", dtype=np.uint32)
        assert np.array_equal(a, b), name
