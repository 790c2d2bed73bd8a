"""Reader / writer of the checkpoint files hg_checkpoint_save produces
(hydro_gen_b200/csrc/hg_checkpoint.cu): host-side tooling for offline rendering, inspection
and for building a checkpoint from arrays.  Pure numpy: it never touches the GPU.

    HgCkptHeader (little endian)  |  json_bytes of informational JSON  |  fields  |  particles

Fields are float32 [rows][map_w][4] images in the reference's texture format (RGBA32F)."""
import ctypes as C
import json

import numpy as np

from ._lib import (ErosionData, FIELD_FLUX, FIELD_HEIGHTMAP, FIELD_SEDIMENT, FIELD_VELOCITY, HG_GRID, HG_PARTICLES,
                   MapSettingsData, RainData)

MAGIC = b"HGCKPT01"
FIELD_NAMES = {FIELD_HEIGHTMAP: "heightmap", FIELD_FLUX: "flux", FIELD_VELOCITY: "velocity", FIELD_SEDIMENT: "sediment"}
PARTICLE_BYTES = 48


class CkptHeader(C.LittleEndianStructure):
    """HgCkptHeader of hg_checkpoint.cu, field for field."""
    _fields_ = [("magic", C.c_char * 8), ("version", C.c_uint32), ("header_bytes", C.c_uint32), ("json_bytes", C.c_uint32),
                ("map_w", C.c_uint32), ("map_h", C.c_uint32), ("row0", C.c_uint32), ("rows", C.c_uint32),
                ("erosion_type", C.c_int32), ("particle_count", C.c_uint32), ("erosion_steps", C.c_uint32),
                ("n_fields", C.c_uint32), ("field_ids", C.c_int32 * 8),
                ("erosion", ErosionData), ("rain", RainData), ("_pad0", C.c_uint32), ("map", MapSettingsData),
                ("_pad1", C.c_uint32), ("payload_bytes", C.c_uint64)]


assert (C.sizeof(CkptHeader), CkptHeader.erosion.offset, CkptHeader.map.offset, CkptHeader.payload_bytes.offset) == (312, 84, 204, 304)


class Checkpoint:
    def __init__(self, header, info, fields, particles=None):
        self.header, self.info, self.fields, self.particles = header, info, fields, particles

    @property
    def erosion_steps(self):
        return self.header.erosion_steps


def read_checkpoint(path):
    """-> Checkpoint: header (CkptHeader), info (the JSON block as a dict), fields {name: float32 [rows, W, 4]},
    particles (raw uint8 [count, 48] or None).  Raises ValueError on a malformed file."""
    with open(path, "rb") as f:
        raw = f.read(C.sizeof(CkptHeader))
        if len(raw) != C.sizeof(CkptHeader):
            raise ValueError(f"{path}: too short for a checkpoint header")
        h = CkptHeader.from_buffer_copy(raw)
        if h.magic != MAGIC or h.version != 1 or h.header_bytes != C.sizeof(CkptHeader) or h.n_fields > 8:
            raise ValueError(f"{path}: not a hydrogen_b200 checkpoint (version 1)")
        info = json.loads(f.read(h.json_bytes).decode("utf-8")) if h.json_bytes else {}
        n = h.rows * h.map_w * 4
        nparts = h.particle_count if h.erosion_type == HG_PARTICLES else 0
        if h.payload_bytes != h.n_fields * n * 4 + nparts * PARTICLE_BYTES:
            raise ValueError(f"{path}: payload size does not match its header")
        fields = {}
        for k in range(h.n_fields):
            a = np.fromfile(f, dtype="<f4", count=n)
            if a.size != n:
                raise ValueError(f"{path}: truncated in field {k}")
            fields[FIELD_NAMES.get(h.field_ids[k], str(h.field_ids[k]))] = a.reshape(h.rows, h.map_w, 4)
        particles = None
        if nparts:
            particles = np.fromfile(f, dtype=np.uint8, count=nparts * PARTICLE_BYTES)
            if particles.size != nparts * PARTICLE_BYTES:
                raise ValueError(f"{path}: truncated in the droplet array")
            particles = particles.reshape(nparts, PARTICLE_BYTES)
    return Checkpoint(h, info, fields, particles)


def write_checkpoint(path, map_w, map_h, fields, erosion, rain, map_settings, erosion_steps=0, row0=0, rows=None,
                     erosion_type=HG_GRID, particles=None):
    """Build a checkpoint from host arrays (same layout hg_checkpoint_save writes, loadable by
    hg_checkpoint_load).  fields: {FIELD_* id: float32 [rows, map_w, 4]} in file order."""
    rows = map_h if rows is None else rows
    h = CkptHeader()
    h.magic, h.version, h.header_bytes = MAGIC, 1, C.sizeof(CkptHeader)
    h.map_w, h.map_h, h.row0, h.rows = map_w, map_h, row0, rows
    h.erosion_type, h.erosion_steps, h.n_fields = erosion_type, erosion_steps, len(fields)
    nparts = 0 if particles is None else len(particles)
    h.particle_count = nparts
    for k, fid in enumerate(fields):
        h.field_ids[k] = fid
    h.erosion, h.rain, h.map = erosion, rain, map_settings
    arrays = []
    for fid, a in fields.items():
        a = np.ascontiguousarray(a, dtype="<f4")
        if a.shape != (rows, map_w, 4):
            raise ValueError(f"field {fid}: shape {a.shape}, expected {(rows, map_w, 4)}")
        arrays.append(a)
    h.payload_bytes = sum(a.nbytes for a in arrays) + nparts * PARTICLE_BYTES
    info = json.dumps({"format": "hydrogen_b200 checkpoint", "version": 1, "map": [map_w, map_h], "row0": row0, "rows": rows,
                       "erosion_type": "grid" if erosion_type == HG_GRID else "particles", "particle_count": nparts,
                       "erosion_steps": erosion_steps, "fields": [FIELD_NAMES.get(f, str(f)) for f in fields],
                       "field_layout": "float32 little endian [row][x][rgba]"}).encode("utf-8")
    h.json_bytes = len(info)
    with open(path, "wb") as f:
        f.write(bytes(h)); f.write(info)
        for a in arrays:
            a.tofile(f)
        if nparts:
            np.ascontiguousarray(particles, dtype=np.uint8).tofile(f)
