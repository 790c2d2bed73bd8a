"""Host-side logic of the multi-GPU row-slab run: how the map rows are divided
between ranks and how the ranks exchange the CUDA IPC handles of their arenas.  One process per
GPU; `dist` is an initialised torch.distributed (NCCL on GPUs, gloo in the CPU tests).  No
collective is on the data path: after connect() every step pushes halo rows over NVLink from
inside the step's own stream."""
from . import _lib


def slab_rows(map_h, world, rank):
    """Rows [row0, row0+rows) of rank `rank`: multiples of 8 (the reference's work-group height),
    as even as possible, earlier ranks take the remainder; every slab at least HALO_ROWS rows."""
    if map_h % 8 or world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad slab request: map_h={map_h} world={world} rank={rank}")
    groups = map_h // 8
    if groups < world:
        raise ValueError(f"{map_h} rows cannot be split into {world} slabs of >= 8 rows")
    base, extra = divmod(groups, world)
    row0 = 8 * (rank * base + min(rank, extra))
    rows = 8 * (base + (1 if rank < extra else 0))
    assert rows >= _lib.HALO_ROWS
    return row0, rows


def gather_exports(blob, dist, world):
    """all-gather of each rank's hg_slab_export blob (bytes), ordered by rank."""
    if world == 1:
        return [blob]
    out = [None] * world
    dist.all_gather_object(out, bytes(blob))
    return out


def check_exports(exports, map_w, map_h):
    """The table every rank passes to hg_slab_connect must tile the map in rank order."""
    nxt = 0
    for k, e in enumerate(exports):
        if (e.map_w, e.map_h) != (map_w, map_h) or e.row0 != nxt:
            raise ValueError(f"slab {k} covers rows [{e.row0},{e.row0 + e.rows}) of a {e.map_w}x{e.map_h} map; expected row0 {nxt}")
        nxt += e.rows
    if nxt != map_h:
        raise ValueError(f"slabs cover {nxt} of {map_h} rows")


def connect_ring(ctx, dist, world, rank):
    """Export this rank's arena, gather everyone's, open the peers (hg_slab_connect)."""
    blobs = gather_exports(bytes(ctx.export_handle()), dist, world)
    exports = [_lib.SlabExport.from_buffer_copy(b) for b in blobs]
    check_exports(exports, ctx.W, ctx.H)
    if world > 1:
        ctx.connect(exports, rank)
        if ctx.particle_count and ctx.erosion_type == _lib.HG_PARTICLES:
            # droplet slabs: the texel images, the droplet array and the ownership bytes travel the same way
            pblobs = gather_exports(bytes(ctx.export_particles()), dist, world)
            ctx.connect_particles([_lib.SlabExportParticles.from_buffer_copy(b) for b in pblobs], rank)
    return exports


def owner_of_row(row, table):
    """index of the slab (row0, rows) that holds map row `row`: the rank that owns a droplet at that row"""
    for k, (row0, rows) in enumerate(table):
        if row0 <= row < row0 + rows:
            return k
    raise ValueError(f"row {row} is outside every slab")


def merge_droplets(parts_per_slab, owners_per_slab):
    """The droplet array of the whole map from the slabs' arrays: element id comes from the slab that owns it.
    Raises if a droplet has no owner or several (the hand-over invariant)."""
    import numpy as np
    own = np.stack(owners_per_slab).astype(np.int32)
    n_owner = own.sum(axis=0)
    if not np.all(n_owner == 1):
        bad = np.flatnonzero(n_owner != 1)
        raise ValueError(f"{bad.size} droplets do not have exactly one owner (first: id {bad[0]} has {n_owner[bad[0]]})")
    who = own.argmax(axis=0)
    out = parts_per_slab[0].copy()
    for k, p in enumerate(parts_per_slab):
        sel = who == k
        out[sel] = p[sel]
    return out
