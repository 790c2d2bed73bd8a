"""Checkpoint format: hg_checkpoint_save / hg_checkpoint_load in the C ABI and
the numpy reader/writer of the same files (hydro_gen_b200/checkpoint.py).

CPU: the Python writer and reader agree on the layout, malformed files are refused.
GPU: a file written by the C ABI reads back as exactly what hg_download returns; a run resumed from a
checkpoint continues bit for bit (fields, step counter, rain schedule); a checkpoint built from host
arrays loads; a context of another geometry refuses it."""
import ctypes as C

import numpy as np
import pytest

from hydro_gen_b200 import _lib, checkpoint
from tests.util import DT_TIME, SEED, assert_bit_equal


def _settings():
    e, r, m = _lib.ErosionData(), _lib.RainData(), _lib.MapSettingsData()
    e.Kc, e.d_t, e.G = 0.06, 0.005, 9.81
    e.Kalpha[0], e.Kalpha[1] = 0.9, 0.7
    r.amount, r.period, r.drops = 0.01, 16, 0.02
    m.seed, m.max_height, m.octaves = SEED, 256.0, 8
    m.hmap_dims[0], m.hmap_dims[1] = 48, 32
    return e, r, m


def test_python_writer_reader_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    W, H = 48, 32
    fields = {_lib.FIELD_HEIGHTMAP: rng.random((H, W, 4), dtype=np.float32),
              _lib.FIELD_FLUX: rng.random((H, W, 4), dtype=np.float32),
              _lib.FIELD_SEDIMENT: rng.random((H, W, 4), dtype=np.float32)}
    e, r, m = _settings()
    p = tmp_path / "a.hgck"
    checkpoint.write_checkpoint(p, W, H, fields, e, r, m, erosion_steps=77)
    ck = checkpoint.read_checkpoint(p)
    assert (ck.header.map_w, ck.header.map_h, ck.header.row0, ck.header.rows) == (W, H, 0, H)
    assert ck.erosion_steps == 77 and ck.info["erosion_steps"] == 77 and ck.info["fields"] == ["heightmap", "flux", "sediment"]
    assert bytes(ck.header.erosion) == bytes(e) and bytes(ck.header.rain) == bytes(r) and bytes(ck.header.map) == bytes(m)
    for fid, name in ((_lib.FIELD_HEIGHTMAP, "heightmap"), (_lib.FIELD_FLUX, "flux"), (_lib.FIELD_SEDIMENT, "sediment")):
        assert_bit_equal(ck.fields[name], fields[fid], name)
    assert ck.particles is None
    # slab + droplets
    parts = rng.integers(0, 255, (10, 48), dtype=np.uint8)
    q = tmp_path / "b.hgck"
    checkpoint.write_checkpoint(q, W, 64, {_lib.FIELD_HEIGHTMAP: fields[_lib.FIELD_HEIGHTMAP]}, e, r, m, row0=32, rows=H,
                                erosion_type=_lib.HG_PARTICLES, particles=parts)
    ck = checkpoint.read_checkpoint(q)
    assert (ck.header.row0, ck.header.rows, ck.header.particle_count) == (32, H, 10)
    assert np.array_equal(ck.particles, parts)


def test_reader_refuses_malformed(tmp_path):
    e, r, m = _settings()
    W, H = 16, 8
    f = {_lib.FIELD_HEIGHTMAP: np.zeros((H, W, 4), np.float32)}
    p = tmp_path / "c.hgck"
    checkpoint.write_checkpoint(p, W, H, f, e, r, m)
    raw = p.read_bytes()
    (tmp_path / "trunc").write_bytes(raw[:-5])
    (tmp_path / "magic").write_bytes(b"XXXXXXXX" + raw[8:])
    (tmp_path / "short").write_bytes(raw[:100])
    for name in ("trunc", "magic", "short"):
        with pytest.raises(ValueError):
            checkpoint.read_checkpoint(tmp_path / name)
    with pytest.raises(ValueError):
        checkpoint.write_checkpoint(tmp_path / "d", W, H, {_lib.FIELD_HEIGHTMAP: np.zeros((H, W, 3), np.float32)}, e, r, m)


def _wet_ctx(n=128):
    from hydro_gen_b200 import Context
    ctx = Context(n)
    m = ctx.get_map(); m.seed = SEED; ctx.set_map(m)
    rn = ctx.get_rain(); rn.period = 4; ctx.set_rain(rn)
    ctx.gen_heightmap()
    return ctx


@pytest.mark.gpu
def test_save_reads_back_and_resume_is_bit_exact(built, tmp_path):
    from hydro_gen_b200 import Context
    a = _wet_ctx()
    a.run(10, DT_TIME, DT_TIME, True)
    p = tmp_path / "run.hgck"
    a.save_checkpoint(p)
    ck = checkpoint.read_checkpoint(p)
    assert ck.erosion_steps == 10 and ck.info["map"] == [128, 128] and ck.header.rain.period == 4
    for fid, name in ((_lib.FIELD_HEIGHTMAP, "heightmap"), (_lib.FIELD_FLUX, "flux"), (_lib.FIELD_SEDIMENT, "sediment")):
        assert_bit_equal(ck.fields[name], a.download(fid), f"file vs device: {name}")
    a.run(9, 11 * DT_TIME, DT_TIME, True)          # crosses rain steps 12, 16
    b = Context(128)                               # default settings, no heightmap: everything comes from the file
    b.load_checkpoint(p)
    assert b.steps == 10 and b.get_rain().period == 4 and b.get_map().seed == np.float32(SEED)
    b.run(9, 11 * DT_TIME, DT_TIME, True)
    for fid, name in ((_lib.FIELD_HEIGHTMAP, "heightmap"), (_lib.FIELD_FLUX, "flux"), (_lib.FIELD_SEDIMENT, "sediment")):
        assert_bit_equal(b.download(fid), a.download(fid), f"resumed run: {name}")
    # a checkpoint built on the host from the same arrays loads to the same state
    q = tmp_path / "host.hgck"
    checkpoint.write_checkpoint(q, 128, 128, {_lib.FIELD_HEIGHTMAP: ck.fields["heightmap"], _lib.FIELD_FLUX: ck.fields["flux"],
                                              _lib.FIELD_SEDIMENT: ck.fields["sediment"]},
                                ck.header.erosion, ck.header.rain, ck.header.map, erosion_steps=10)
    c = Context(128)
    c.load_checkpoint(q)
    c.run(9, 11 * DT_TIME, DT_TIME, True)
    assert_bit_equal(c.download(_lib.FIELD_HEIGHTMAP), a.download(_lib.FIELD_HEIGHTMAP), "host-built checkpoint")
    # wrong geometry / not a checkpoint
    d = Context(64)
    with pytest.raises(_lib.HydrogenError):
        d.load_checkpoint(p)
    (tmp_path / "junk").write_bytes(b"not a checkpoint at all" * 20)
    with pytest.raises(_lib.HydrogenError):
        b.load_checkpoint(tmp_path / "junk")
    for x in (a, b, c, d):
        x.close()


@pytest.mark.gpu
def test_particle_checkpoint_roundtrip(built, tmp_path):
    from hydro_gen_b200 import Context
    n, count = 64, 256
    a = Context(n, particle_count=count, erosion_type=_lib.HG_PARTICLES)
    m = a.get_map(); m.seed = SEED; a.set_map(m)
    a.gen_heightmap()
    for k in range(5):
        a.dispatch_particle((k + 1) * DT_TIME, True)
    p = tmp_path / "drops.hgck"
    a.save_checkpoint(p)
    ck = checkpoint.read_checkpoint(p)
    assert ck.header.particle_count == count and set(ck.fields) == {"heightmap", "velocity"}
    b = Context(n, particle_count=count, erosion_type=_lib.HG_PARTICLES)
    b.load_checkpoint(p)
    assert_bit_equal(b.download(_lib.FIELD_HEIGHTMAP), a.download(_lib.FIELD_HEIGHTMAP), "heightmap")
    assert_bit_equal(b.download(_lib.FIELD_VELOCITY), a.download(_lib.FIELD_VELOCITY), "momentum map")
    assert np.array_equal(np.asarray(b.download_particles()).view(np.uint8), np.asarray(a.download_particles()).view(np.uint8))
    a.close(); b.close()


@pytest.mark.gpu
def test_slab_checkpoint_resume_matches_whole_map(built, tmp_path):
    """Three connected slabs save their own rows and are resumed into three NEW connected slabs: the load ends with
    the collective halo refresh (hg_slab_refresh_halo), so the first fused step after the resume reads the neighbours'
    edge rows, and the resumed slabs keep reproducing the whole-map run bit for bit, rain steps included.
    A truncated file is refused before it touches the context."""
    from hydro_gen_b200 import Context
    n, cuts = 128, [0, 40, 88, 128]

    def make_slabs():
        ss = [Context(n, n, row0=cuts[i], rows=cuts[i + 1] - cuts[i]) for i in range(3)]
        for i, s in enumerate(ss):
            s.connect_local(ss, i)
        return ss

    def setup(c):
        m = c.get_map(); m.seed = SEED; c.set_map(m)
        r = c.get_rain(); r.period = 4; c.set_rain(r)
        c.gen_heightmap()

    whole = Context(n)
    setup(whole)
    slabs = make_slabs()
    for s in slabs:
        setup(s)
    for k in range(10):
        t = (k + 1) * DT_TIME
        whole.run(1, t, DT_TIME, True)
        for s in slabs:
            s.run(1, t, DT_TIME, True)
    paths = [tmp_path / f"slab{i}.hgck" for i in range(3)]
    for s, p in zip(slabs, paths):
        s.save_checkpoint(p)
        s.close()
    resumed = make_slabs()          # fresh contexts: zero ghost rows, default settings
    for s, p in zip(resumed, paths):
        s.load_checkpoint(p)
    for k in range(10, 19):         # crosses rain steps 12 and 16
        t = (k + 1) * DT_TIME
        whole.run(1, t, DT_TIME, True)
        for s in resumed:
            s.run(1, t, DT_TIME, True)
    for s in resumed:
        s.sync()
        assert s.slab_errors() == 0
    for fid, name in ((_lib.FIELD_HEIGHTMAP, "heightmap"), (_lib.FIELD_FLUX, "flux"), (_lib.FIELD_SEDIMENT, "sediment")):
        got = np.concatenate([s.download(fid) for s in resumed], axis=0)
        assert_bit_equal(got, whole.download(fid), f"resumed slabs vs whole map: {name}")
    # truncated file: refused, and the context keeps its state and settings
    data = paths[0].read_bytes()
    (tmp_path / "cut.hgck").write_bytes(data[: len(data) - 1000])
    before = resumed[0].download(_lib.FIELD_HEIGHTMAP)
    steps = resumed[0].steps
    with pytest.raises(_lib.HydrogenError):
        resumed[0].load_checkpoint(tmp_path / "cut.hgck")
    assert resumed[0].steps == steps
    assert_bit_equal(resumed[0].download(_lib.FIELD_HEIGHTMAP), before, "state after a refused load")
    assert not (tmp_path / "slab0.hgck.tmp").exists()
    for s in resumed:
        s.close()
    whole.close()
