"""Thin object wrapper over the C ABI handle (hg_ctx).  Host buffers are numpy arrays in
the reference's texture format: float32, shape (rows, W, 4) (State::World::Textures,
src/state.hpp:59-76)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (ErosionData, MapSettingsData, RainData, SlabExport, SlabExportParticles, check)

PARTICLE_DTYPE = np.dtype([("sc", "<f4"), ("iters", "<i4"), ("position", "<f4", 2), ("velocity", "<f4", 2),
                           ("volume", "<f4"), ("_pad0", "<u4"), ("sediment", "<f4", 2), ("to_kill", "<u4"),
                           ("_pad1", "<u4")])   # hg_particle == Particle (glsl/bindings.glsl:101-111)
assert PARTICLE_DTYPE.itemsize == 48


class PinnedBuffer:
    """Pinned host memory from hg_host_alloc, viewed as a numpy array."""

    def __init__(self, shape, dtype=np.float32):
        self.L = _lib.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.L.hg_host_alloc(self.nbytes)
        if not self.ptr:
            raise _lib.HydrogenError(self.L.hg_last_error().decode())
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.hg_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, map_w, map_h=None, particle_count=0, erosion_type=_lib.HG_GRID, device=0, row0=None, rows=None):
        self.L = _lib.load()
        self.W = int(map_w)
        self.H = int(map_h if map_h is not None else map_w)
        self.row0 = 0 if row0 is None else int(row0)
        self.rows = self.H if rows is None else int(rows)
        self.particle_count = int(particle_count)
        self.erosion_type = int(erosion_type)
        self.device = int(device)
        self.h = self.L.hg_create_slab(self.W, self.H, self.row0, self.rows, self.particle_count, self.erosion_type, self.device)
        if not self.h:
            raise _lib.HydrogenError(self.L.hg_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.hg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- settings (push_data, src/settings.hpp:22,56,64)
    def get_erosion(self):
        d = ErosionData(); check(self.L.hg_get_erosion(self.h, C.byref(d))); return d

    def set_erosion(self, d):
        check(self.L.hg_set_erosion(self.h, C.byref(d)))

    def get_rain(self):
        d = RainData(); check(self.L.hg_get_rain(self.h, C.byref(d))); return d

    def set_rain(self, d):
        check(self.L.hg_set_rain(self.h, C.byref(d)))

    def get_map(self):
        d = MapSettingsData(); check(self.L.hg_get_map(self.h, C.byref(d))); return d

    def set_map(self, d):
        check(self.L.hg_set_map(self.h, C.byref(d)))

    def set_schedule(self, schedule):
        check(self.L.hg_set_schedule(self.h, int(schedule)))

    # ---- dispatch
    def gen_heightmap(self):
        check(self.L.hg_gen_heightmap(self.h))

    def dispatch_grid_rain(self, time):
        check(self.L.hg_dispatch_grid_rain(self.h, float(time)))

    def dispatch_grid(self):
        check(self.L.hg_dispatch_grid(self.h))

    def dispatch_particle(self, time, should_rain=True):
        check(self.L.hg_dispatch_particle(self.h, float(time), int(should_rain)))

    def dispatch_pass(self, which):
        check(self.L.hg_dispatch_pass(self.h, int(which)))

    def dispatch_particle_pass(self, which, time, should_rain=True):
        check(self.L.hg_dispatch_particle_pass(self.h, int(which), float(time), int(should_rain)))

    def run(self, n_steps, time0=0.0, dtime=0.015, should_rain=True):
        check(self.L.hg_run(self.h, int(n_steps), float(time0), float(dtime), int(should_rain)))

    @property
    def steps(self):
        v = C.c_uint32(); check(self.L.hg_get_steps(self.h, C.byref(v))); return v.value

    @steps.setter
    def steps(self, v):
        check(self.L.hg_set_steps(self.h, int(v)))

    # ---- transfer
    def _host(self, arr, rows=None):
        rows = self.rows if rows is None else rows
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if a.size != rows * self.W * 4:
            raise ValueError(f"expected {rows}x{self.W}x4 floats, got {a.shape}")
        return a

    def upload(self, field, arr, asynchronous=False):
        a = self._host(arr)
        fn = self.L.hg_upload_async if asynchronous else self.L.hg_upload
        check(fn(self.h, int(field), a.ctypes.data))
        return a   # keep alive until synced when asynchronous

    def download(self, field, out=None, asynchronous=False):
        if out is None:
            out = np.empty((self.rows, self.W, 4), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and out.size == self.rows * self.W * 4
        fn = self.L.hg_download_async if asynchronous else self.L.hg_download
        check(fn(self.h, int(field), out.ctypes.data))
        return out

    def step_host_async(self, ins, outs):
        """hg_step_host_async: one dispatch_grid from host images (H, F, S) to host images, pipelined
        over copy streams.  `ins`/`outs`: three float32 (rows, W, 4) arrays each (pinned for overlap);
        none of them may be touched before sync()."""
        for a in list(ins) + list(outs):
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.size == self.rows * self.W * 4
        check(self.L.hg_step_host_async(self.h, *[a.ctypes.data for a in ins], *[a.ctypes.data for a in outs]))

    def upload_particles(self, arr):
        a = np.ascontiguousarray(arr, dtype=PARTICLE_DTYPE)
        check(self.L.hg_upload_particles(self.h, a.ctypes.data, a.shape[0]))

    def download_particles(self):
        out = np.zeros(self.particle_count, dtype=PARTICLE_DTYPE)
        check(self.L.hg_download_particles(self.h, out.ctypes.data, self.particle_count))
        return out

    def set_ghost(self, field, side, rows_arr):
        a = self._host(rows_arr, _lib.HALO_ROWS)
        check(self.L.hg_slab_set_ghost(self.h, int(field), int(side), a.ctypes.data))

    def mass(self):
        out = (C.c_double * 5)()
        check(self.L.hg_mass(self.h, out))
        return np.array(out[:], dtype=np.float64)

    # ---- sync / timing / counters
    def sync(self):
        check(self.L.hg_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        check(self.L.hg_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def timer_start(self):
        check(self.L.hg_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(); check(self.L.hg_timer_stop(self.h, C.byref(ms))); return ms.value

    def run_profiled(self, n_steps, time0=0.0, dtime=0.015, should_rain=True):
        """run() timed on the device: returns (ms of the whole run, average ms of the fused step kernel inside it)"""
        k, tot = C.c_float(), C.c_float()
        check(self.L.hg_run_profiled(self.h, int(n_steps), float(time0), float(dtime), int(should_rain), C.byref(k), C.byref(tot)))
        return tot.value, k.value

    def profile_fused(self, n_steps):
        ms = C.c_float(); check(self.L.hg_profile_fused(self.h, int(n_steps), C.byref(ms))); return ms.value

    @property
    def launch_count(self):
        return int(self.L.hg_launch_count(self.h))

    def far_fetch_count(self):
        v = C.c_uint64(); check(self.L.hg_far_fetch_count(self.h, C.byref(v))); return v.value

    def slab_errors(self):
        v = C.c_uint64(); check(self.L.hg_slab_errors(self.h, C.byref(v))); return v.value

    def pack_device(self, field, device_ptr):
        """hg_pack_device: the field as an RGBA32F image in device memory (rows*W*4 floats at device_ptr), asynchronous"""
        check(self.L.hg_pack_device(self.h, int(field), C.c_void_p(int(device_ptr))))

    def refresh_halo(self):
        """hg_slab_refresh_halo: collective re-fill of the ghost rows after uploads on a connected slab"""
        check(self.L.hg_slab_refresh_halo(self.h))

    # ---- checkpoint (hg_checkpoint.cu; hydro_gen_b200.checkpoint reads the files without a GPU)
    def save_checkpoint(self, path):
        check(self.L.hg_checkpoint_save(self.h, str(path).encode()))

    def load_checkpoint(self, path):
        check(self.L.hg_checkpoint_load(self.h, str(path).encode()))

    # ---- slabs
    def export_handle(self):
        e = SlabExport(); check(self.L.hg_slab_export_handle(self.h, C.byref(e))); return e

    def connect(self, exports, my_index):
        arr = (SlabExport * len(exports))(*exports)
        check(self.L.hg_slab_connect(self.h, arr, len(exports), int(my_index)))

    def export_particles(self):
        e = SlabExportParticles(); check(self.L.hg_slab_export_particles(self.h, C.byref(e))); return e

    def connect_particles(self, exports, my_index):
        arr = (SlabExportParticles * len(exports))(*exports)
        check(self.L.hg_slab_connect_particles(self.h, arr, len(exports), int(my_index)))

    def particle_owners(self):
        """1 per droplet this slab owns (holds the current state of); a whole map owns all"""
        out = np.zeros(self.particle_count, dtype=np.uint8)
        check(self.L.hg_slab_particle_owners(self.h, out.ctypes.data, self.particle_count))
        return out

    def connect_local(self, contexts, my_index):
        arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
        check(self.L.hg_slab_connect_local(self.h, arr, len(contexts), int(my_index)))
