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
