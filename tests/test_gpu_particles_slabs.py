"""GPU tests of the droplet mode and of row-slab sharding (§8e)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from hydro_gen_b200 import Context, _lib
from tests.util import DT_TIME, FIELDS, SEED, assert_bit_equal, copy_state, max_rel_err, wet_world

pytestmark = pytest.mark.gpu


def _particle_pair(n, count, hmap=None):
    ref = oracle.World(n, particle_count=count, erosion_type=1, seed=SEED)
    ref.gen_heightmap()
    if hmap:
        ref.map.hmap_dims[0], ref.map.hmap_dims[1] = hmap, hmap
    ctx = Context(n, particle_count=count, erosion_type=_lib.HG_PARTICLES)
    ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(ref.map)))
    ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(ref.erosion)))
    copy_state(ref, ctx, ("heightmap", "velocity"))
    return ctx, ref


def test_particle_move_bit_exact(built):
    """particle.glsl: spawn hash (defined sin), normals from bilinear samples, velocity update,
    kill tests.  One thread per droplet, no interaction: bit-exact."""
    ctx, ref = _particle_pair(256, 4096, hmap=256)
    for k in range(1, 4):
        ctx.dispatch_particle_pass(0, k * DT_TIME, True)
        ref.particle_pass(0, k * DT_TIME, True)
        got, want = ctx.download_particles(), ref.particles()
        assert got.tobytes() == want.tobytes(), f"droplets differ after move {k}"
    ctx.close(); ref.close()


def test_particle_erode_sparse_bit_exact(built):
    """particle_erosion.glsl with droplets sparse enough that no two share a texel in a step
    is order-free, so the atomics path must match the sequential oracle bit for bit."""
    ctx, ref = _particle_pair(512, 64, hmap=512)
    for k in range(1, 6):
        ctx.dispatch_particle_pass(0, k * DT_TIME, True)
        ctx.dispatch_particle_pass(1, k * DT_TIME, True)
        ref.particle_pass(0, k * DT_TIME, True)
        ref.particle_pass(1, k * DT_TIME, True)
    assert ctx.download_particles().tobytes() == ref.particles().tobytes()
    for name in ("heightmap", "velocity"):
        got, want = ctx.download(FIELDS[name]), ref.get(FIELDS[name])
        if name == "heightmap":   # H.a is left stale by the erode pass until thermal transport rewrites it
            got, want = got[..., :3], want[..., :3]
        assert_bit_equal(got, want, f"sparse erode: {name}")
    ctx.close(); ref.close()


def test_particle_full_step_statistical(built):
    """Erosion::dispatch_particle with contention (16 droplets per texel on average).  The
    reference itself is lock-order dependent: compare against the oracle's
    id-ordered run within float-reassociation noise, and check what the terrain gained is what
    the droplets lost where the algorithm conserves it."""
    n, count = 128, 262144
    ctx, ref = _particle_pair(n, count, hmap=128)
    for k in range(1, 4):
        ctx.dispatch_particle(k * DT_TIME, True)
        ref.dispatch_particle(k * DT_TIME, True)
    got, want = ctx.download(0), ref.get(0)
    for ch, name in enumerate(("rock", "dirt", "water")):
        scale = np.abs(want[..., ch]).max() + 1e-30
        assert np.abs(got[..., ch] - want[..., ch]).max() / scale < 2e-5, name
    gm, wm = ctx.download(2), ref.get(2)
    assert np.abs(gm - wm).max() <= 2e-5 * (np.abs(wm).max() + 1e-30)
    gp, wp = ctx.download_particles(), ref.particles()
    assert np.array_equal(gp["iters"], wp["iters"]) and np.array_equal(gp["to_kill"], wp["to_kill"])
    np.testing.assert_allclose(gp["position"], wp["position"], rtol=0, atol=1e-3)
    # Mass accounting.  The algorithm itself does not conserve mass (the droplet keeps the sediment
    # of its last corner only), so the invariant is against the oracle: the totals of every map channel and of the
    # sediment the droplets carry are sums of the same terms in another order and must agree to reassociation noise.
    for ch, name in enumerate(("rock", "dirt", "water")):
        a, b = got[..., ch].sum(dtype=np.float64), want[..., ch].sum(dtype=np.float64)
        assert abs(a - b) <= 1e-6 * abs(b) + 1e-6, f"map total of {name}: {a} vs {b}"
    for i in range(2):
        a, b = gp["sediment"][:, i].sum(dtype=np.float64), wp["sediment"][:, i].sum(dtype=np.float64)
        assert abs(a - b) <= 1e-4 * abs(b) + 1e-6, f"carried sediment of layer {i}: {a} vs {b}"
    ctx.close(); ref.close()


@pytest.mark.parametrize("fused_push,cuts", [("1", [0, 96, 168, 256]), ("0", [0, 96, 168, 256]), ("1", [0, 8, 24, 256])])
def test_slabs_on_one_gpu_match_whole_map(built, monkeypatch, fused_push, cuts):
    """Three row slabs (one context each, same GPU, peer pointers inside the process) step in
    lock step through the device-side halo push + flag wait and must reproduce the whole-map
    result bit for bit, far fetches across slab boundaries included.  fused_push "1": the step kernel
    and the fix-up store the edge rows into the neighbours' ghost rows themselves and the fix-up's last
    block signals (the default); "0": the separate push kernel.  The third case has slabs as thin as the
    halo (8 and 16 rows): every row of them is an edge row for both neighbours."""
    monkeypatch.setenv("HG_FUSED_PUSH", fused_push)
    n = 256
    w = wet_world(n, 300)
    whole = Context(n)
    copy_state(w, whole)
    slabs = [Context(n, n, row0=cuts[i], rows=cuts[i + 1] - cuts[i]) for i in range(3)]
    for i, s in enumerate(slabs):
        s.connect_local(slabs, i)
    for name in ("heightmap", "flux", "sediment"):
        full = w.get(FIELDS[name])
        for i, s in enumerate(slabs):
            s.upload(FIELDS[name], full[cuts[i]:cuts[i + 1]])
            lo = np.zeros((_lib.HALO_ROWS, n, 4), np.float32)
            hi = np.zeros((_lib.HALO_ROWS, n, 4), np.float32)
            if i > 0:
                lo[:] = full[cuts[i] - _lib.HALO_ROWS:cuts[i]]
            if i < 2:
                hi[:] = full[cuts[i + 1]:cuts[i + 1] + _lib.HALO_ROWS]
            s.set_ghost(FIELDS[name], 0, lo)
            s.set_ghost(FIELDS[name], 1, hi)
    for _ in range(6):
        whole.dispatch_grid()
        for s in slabs:
            s.dispatch_grid()
    for s in slabs:
        s.sync()
        assert s.slab_errors() == 0
    for name in ("heightmap", "flux", "sediment"):
        full = whole.download(FIELDS[name])
        got = np.concatenate([s.download(FIELDS[name]) for s in slabs], axis=0)
        assert_bit_equal(got, full, f"slabs vs whole: {name}")
    for s in slabs:
        s.close()
    whole.close(); w.close()


def test_particle_tail_fused_equals_passes(built):
    """The grid part of Erosion::dispatch_particle (thermal x2 + smoothing with the momentum map) on the fused kernel
    (FUSED schedule) against the five 1:1 pass kernels (PASSES schedule) and against the oracle: bit for bit.
    Droplets sparse enough that the erosion pass is order-free; steep talus so both thermal layers move terrain."""
    n, count = 512, 64
    def make(schedule):
        ref = oracle.World(n, particle_count=count, erosion_type=1, seed=SEED)
        ref.gen_heightmap()
        ref.map.hmap_dims[0], ref.map.hmap_dims[1] = n, n
        ref.erosion.Kalpha[0], ref.erosion.Kalpha[1] = 0.5, 0.2
        ctx = Context(n, particle_count=count, erosion_type=_lib.HG_PARTICLES)
        ctx.set_schedule(schedule)
        ctx.set_map(_lib.MapSettingsData.from_buffer_copy(bytes(ref.map)))
        ctx.set_erosion(_lib.ErosionData.from_buffer_copy(bytes(ref.erosion)))
        copy_state(ref, ctx, ("heightmap", "velocity"))
        return ctx, ref
    a, ref = make(_lib.SCHEDULE_FUSED)
    b, ref_b = make(_lib.SCHEDULE_PASSES)
    ref_b.close()
    for k in range(1, 7):
        t = float(np.float32(k) * np.float32(DT_TIME))
        a.dispatch_particle(t, True); b.dispatch_particle(t, True); ref.dispatch_particle(t, True)
    assert a.download_particles().tobytes() == b.download_particles().tobytes() == ref.particles().tobytes()
    for name in ("heightmap", "velocity"):
        assert_bit_equal(a.download(FIELDS[name]), b.download(FIELDS[name]), f"fused vs passes tail: {name}")
        assert_bit_equal(a.download(FIELDS[name]), ref.get(FIELDS[name]), f"fused tail vs oracle: {name}")
    H0 = ref.get(0)
    assert np.abs(a.download(FIELDS["velocity"])).max() > 0
    a.close(); b.close(); ref.close()


def test_droplet_slabs_match_whole_map(built):
    """Droplet mode on row slabs: three slabs in one process (peer pointers) against the
    whole-map run, in the sparse regime where the result is order-free: droplet array (merged from the owners), the
    heightmap and the momentum map must be identical bit for bit after every dispatch.  The droplets are placed along
    the two slab edges and given velocities across them, so the run exercises what sharding adds: hand-over of drifting
    and respawning droplets, erosion of corner texels in the neighbour's rows (NVLink-style peer atomics), the gather
    halo of the move pass and the two image exchanges per step."""
    from hydro_gen_b200 import slabs as slabmod
    n, count = 256, 64
    cuts = [0, 88, 176, 256]
    table = [(cuts[i], cuts[i + 1] - cuts[i]) for i in range(3)]

    def setup(c):
        m = c.get_map(); m.seed = SEED; m.hmap_dims[0], m.hmap_dims[1] = n, n; c.set_map(m)
        e = c.get_erosion(); e.Kalpha[0], e.Kalpha[1] = 0.5, 0.2; c.set_erosion(e)      # both thermal layers move terrain
        c.gen_heightmap()

    whole = Context(n, particle_count=count, erosion_type=_lib.HG_PARTICLES)
    setup(whole)
    parts = [Context(n, n, particle_count=count, erosion_type=_lib.HG_PARTICLES, row0=r0, rows=rows) for r0, rows in table]
    for i, s in enumerate(parts):
        s.connect_local(parts, i)
    for s in parts:
        setup(s)
    crossings = 0
    owners_before = None
    for k in range(1, 41):
        t = float(np.float32(k) * np.float32(DT_TIME))
        if k == 3:
            # move the droplets onto the slab edges with velocities across them (same state everywhere; the owner of
            # each is the slab that holds its row, which is what the hand-over rule keeps true)
            P = whole.download_particles()
            rng = np.random.default_rng(11)
            edge = np.where(np.arange(count) % 2 == 0, cuts[1], cuts[2]).astype(np.float32)
            P["position"][:, 1] = edge + rng.uniform(-0.6, 0.6, count).astype(np.float32)
            P["position"][:, 0] = np.linspace(8, n - 8, count).astype(np.float32)      # far apart in x: sparse
            P["velocity"][:, 1] = np.where(rng.random(count) < 0.5, 0.9, -0.9).astype(np.float32)
            P["velocity"][:, 0] = 0.0
            whole.upload_particles(P)
            for s in parts:
                s.upload_particles(P)
        whole.dispatch_particle(t, True)
        for s in parts:
            s.dispatch_particle(t, True)
        for s in parts:
            s.sync()
            assert s.slab_errors() == 0
        owners = [s.particle_owners() for s in parts]
        got = slabmod.merge_droplets([s.download_particles() for s in parts], owners)
        want = whole.download_particles()
        assert got.tobytes() == want.tobytes(), f"droplets differ after dispatch {k}"
        who = np.stack(owners).argmax(axis=0)
        rows_of = np.clip(want["position"][:, 1].astype(np.int64), 0, n - 1)
        assert all(who[i] == slabmod.owner_of_row(int(rows_of[i]), table) for i in range(count) if want["iters"][i] != 0), "a droplet is not with the slab that holds its row"
        if owners_before is not None:
            crossings += int((who != owners_before).sum())
        owners_before = who
        for f in (0, 2):
            full = whole.download(f)
            sl = np.concatenate([s.download(f) for s in parts], axis=0)
            assert_bit_equal(sl, full, f"field {f} after dispatch {k}")
    assert crossings >= 10, f"only {crossings} hand-overs: the test does not exercise migration"
    for s in parts:
        s.close()
    whole.close()


def test_refresh_halo_after_plain_uploads(built):
    """hg_upload on a connected slab replaces the owned rows only; hg_slab_refresh_halo (collective) re-fills the ghost
    rows from the neighbours, after which the slabs reproduce the whole-map run bit for bit without any
    hg_slab_set_ghost (plain uploads left stale ghost rows)."""
    n = 256
    w = wet_world(n, 300)
    whole = Context(n)
    copy_state(w, whole)
    cuts = [0, 64, 200, 256]
    slabs = [Context(n, n, row0=cuts[i], rows=cuts[i + 1] - cuts[i]) for i in range(3)]
    for i, s in enumerate(slabs):
        s.connect_local(slabs, i)
    for name in ("heightmap", "flux", "sediment"):
        full = w.get(FIELDS[name])
        for i, s in enumerate(slabs):
            s.upload(FIELDS[name], full[cuts[i]:cuts[i + 1]])
    for s in slabs:
        s.refresh_halo()
    for _ in range(5):
        whole.dispatch_grid()
        for s in slabs:
            s.dispatch_grid()
    for s in slabs:
        s.sync()
        assert s.slab_errors() == 0
    for name in ("heightmap", "flux", "sediment"):
        got = np.concatenate([s.download(FIELDS[name]) for s in slabs], axis=0)
        assert_bit_equal(got, whole.download(FIELDS[name]), f"slabs after refresh_halo: {name}")
    for s in slabs:
        s.close()
    whole.close(); w.close()


def test_halo_timeout_is_sticky(built, monkeypatch):
    """A rank that never arrives: the waiting slab's halo wait times out (HG_HALO_TIMEOUT_S), and from then on the context
    refuses to step or sync with HG_ERR_STATE instead of silently running on stale ghost rows."""
    monkeypatch.setenv("HG_HALO_TIMEOUT_S", "0.3")
    n = 64
    a = Context(n, n, row0=0, rows=32)
    b = Context(n, n, row0=32, rows=32)
    for i, s in enumerate((a, b)):
        s.connect_local([a, b], i)
    for s in (a, b):
        m = s.get_map(); m.seed = SEED; s.set_map(m)
        s.gen_heightmap()
    a.dispatch_grid()            # pushes and signals generation 1; the wait for b's signal sits in front of the next step
    a.dispatch_grid()            # b never steps: this step's wait times out on the device
    with pytest.raises(_lib.HydrogenError, match="halo wait timed out"):
        a.sync()
    with pytest.raises(_lib.HydrogenError, match="halo wait timed out"):
        a.dispatch_grid()
    assert a.slab_errors() >= 1
    a.close(); b.close()
