"""Host-side multi-rank logic on CPU: world_size-2 (and 3) gloo process groups exercise the slab
partition and the export-table exchange that bench.py and scripts/mgpu_check.py use on NCCL."""
import ctypes as C
import os
import socket

import pytest
import torch.multiprocessing as mp

from hydro_gen_b200 import _lib, slabs


def test_slab_rows_tile_the_map():
    for h, world in ((4096, 1), (4096, 2), (16384, 8), (65536, 8), (1000 * 8, 3), (64, 8)):
        nxt = 0
        for r in range(world):
            row0, rows = slabs.slab_rows(h, world, r)
            assert row0 == nxt and rows % 8 == 0 and rows >= _lib.HALO_ROWS
            nxt += rows
        assert nxt == h
    with pytest.raises(ValueError):
        slabs.slab_rows(4090, 2, 0)
    with pytest.raises(ValueError):
        slabs.slab_rows(32, 8, 0)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H = 1024, 8 * 37
    row0, rows = slabs.slab_rows(H, world, rank)
    e = _lib.SlabExport()
    e.row0, e.rows, e.map_w, e.map_h, e.device, e.arena_bytes = row0, rows, W, H, rank, 1234 + rank
    C.memset(e.mem_handle, 0x40 + rank, 64)
    blobs = slabs.gather_exports(bytes(e), dist, world)
    table = [_lib.SlabExport.from_buffer_copy(b) for b in blobs]
    slabs.check_exports(table, W, H)
    ok = all(bytes(t.mem_handle) == bytes([0x40 + k]) * 64 and t.arena_bytes == 1234 + k for k, t in enumerate(table))
    ok = ok and table[rank].row0 == row0 and sum(t.rows for t in table) == H
    # a broken table (two ranks claiming the same rows) must be refused
    bad = list(table); bad[-1] = table[0]
    try:
        slabs.check_exports(bad, W, H)
        ok = False
    except ValueError:
        pass
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_export_exchange_over_gloo(world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]


def _drops_worker(rank, world, port, q):
    """Droplet hand-over bookkeeping between ranks (CPU, gloo): every rank keeps the whole droplet array and one
    ownership byte per droplet; a droplet belongs to the rank whose rows hold its position.  The ranks move a common
    set of droplets (positions drift up or down, some respawn anywhere), apply the hand-over rule of
    hg_particles.cu (slab_store_droplet: the old owner writes the element and sets / clears the two ownership bytes)
    through an all-gather instead of peer pointers, and must end every step with exactly one owner per droplet, the
    owner being the rank that holds the droplet's row, and the merged array equal to the sequential run."""
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, count = 8 * 24, 256
    table = [slabs.slab_rows(H, world, r) for r in range(world)]
    # the exchange of the droplet allocations' handles (hg_slab_export_particles_t blobs)
    e = _lib.SlabExportParticles()
    e.particle_count = count
    C.memset(e.images_handle, 0x10 + rank, 64); C.memset(e.droplets_handle, 0x20 + rank, 64); C.memset(e.owners_handle, 0x30 + rank, 64)
    blobs = slabs.gather_exports(bytes(e), dist, world)
    tab = [_lib.SlabExportParticles.from_buffer_copy(b) for b in blobs]
    ok = all(bytes(t.images_handle) == bytes([0x10 + k]) * 64 and bytes(t.owners_handle) == bytes([0x30 + k]) * 64 and t.particle_count == count for k, t in enumerate(tab))
    rng = np.random.default_rng(5)                      # same stream on every rank
    y = rng.uniform(2, H - 2, count).astype(np.float32)
    seq = y.copy()                                       # the sequential (whole-map) run
    mine = y.copy()                                      # this rank's copy of the array: authoritative where own == 1
    own = np.array([slabs.owner_of_row(int(v), table) == rank for v in y], np.uint8)
    for step in range(30):
        dy = rng.uniform(-0.25, 0.25, count).astype(np.float32)
        respawn = rng.random(count) < 0.05
        newy = rng.uniform(2, H - 2, count).astype(np.float32)
        seq = np.where(respawn, newy, np.clip(seq + dy, 2, H - 2)).astype(np.float32)
        # owner-computes: only the owner advances a droplet
        upd = np.where(respawn, newy, np.clip(mine + dy, 2, H - 2)).astype(np.float32)
        mine = np.where(own == 1, upd, mine)
        dest = np.array([slabs.owner_of_row(int(v), table) for v in mine], np.int64)
        leaving = (own == 1) & (dest != rank)
        # "peer stores": (id, value, destination) of every droplet this rank hands over
        out = [None] * world
        dist.all_gather_object(out, (np.flatnonzero(leaving), mine[leaving], dest[leaving]))
        own[leaving] = 0
        for ids, vals, dst in out:
            sel = dst == rank
            mine[ids[sel]] = vals[sel]
            own[ids[sel]] = 1
        owners = [None] * world
        dist.all_gather_object(owners, own.copy())
        copies = [None] * world
        dist.all_gather_object(copies, mine.copy())
        merged = slabs.merge_droplets([c for c in copies], owners)       # raises unless exactly one owner each
        ok = ok and np.array_equal(merged, seq)
        ok = ok and all(slabs.owner_of_row(int(seq[i]), table) == int(np.stack(owners)[:, i].argmax()) for i in range(count))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_droplet_handover_bookkeeping_over_gloo(world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_drops_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]


def test_merge_droplets_rejects_broken_ownership():
    import numpy as np
    a = np.arange(4, dtype=np.float32)
    with pytest.raises(ValueError):
        slabs.merge_droplets([a, a], [np.array([1, 1, 0, 0], np.uint8), np.array([0, 1, 1, 1], np.uint8)])      # id 1 owned twice
    with pytest.raises(ValueError):
        slabs.merge_droplets([a, a], [np.array([1, 0, 0, 0], np.uint8), np.array([0, 0, 1, 1], np.uint8)])      # id 1 orphaned
    assert slabs.owner_of_row(17, [(0, 16), (16, 16)]) == 1
