"""ctypes binding of the CPU parity oracle (oracle/hg_oracle.c).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
(hydro_gen_b200) never does.  It is pinned against the reference's own
shaders compiled for the CPU (oracle/refshader/, oracle/refshaders.py; see the
hg_oracle.c header and DESIGN.md §Oracle).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FIELD_H, FIELD_F, FIELD_V, FIELD_S, FIELD_TC, FIELD_TD = range(6)
PASS_FLUX, PASS_EROSION, PASS_SEDIMENT, PASS_THERMAL, PASS_SMOOTH = range(5)


class ErosionData(C.Structure):
    """hg_erosion_data == Erosion_data (glsl/bindings.glsl:39-60)."""
    _fields_ = [("particle_count", C.c_uint32), ("Kc", C.c_float), ("Kalpha", C.c_float * 2),
                ("Kconv", C.c_float), ("_pad0", C.c_uint32), ("Ks", C.c_float * 2), ("Kd", C.c_float * 2),
                ("Ke", C.c_float), ("ENERGY_KEPT", C.c_float), ("Kspeed", C.c_float * 2), ("G", C.c_float),
                ("d_t", C.c_float), ("density", C.c_float), ("init_volume", C.c_float), ("friction", C.c_float),
                ("inertia", C.c_float), ("min_volume", C.c_float), ("min_velocity", C.c_float),
                ("ttl", C.c_uint32), ("_pad1", C.c_uint32)]


class RainData(C.Structure):
    """hg_rain_data == Rain_data (glsl/bindings.glsl:62-68)."""
    _fields_ = [("amount", C.c_float), ("mountain_thresh", C.c_float), ("mountain_multip", C.c_float),
                ("period", C.c_int32), ("drops", C.c_float)]


class MapSettingsData(C.Structure):
    """hg_map_settings_data == Map_settings_data (glsl/bindings.glsl:70-99)."""
    _fields_ = [("max_height", C.c_float), ("max_dirt", C.c_float), ("hmap_dims", C.c_int32 * 2),
                ("height_mult", C.c_float), ("water_lvl", C.c_float), ("seed", C.c_float),
                ("persistance", C.c_float), ("lacunarity", C.c_float), ("scale", C.c_float),
                ("redistribution", C.c_float), ("octaves", C.c_int32), ("fake_erosion", C.c_uint32),
                ("mask_round", C.c_uint32), ("mask_exp", C.c_uint32), ("mask_power", C.c_uint32),
                ("mask_slope", C.c_uint32), ("uplift", C.c_uint32), ("uplift_scale", C.c_float),
                ("domain_warp", C.c_int32), ("domain_warp_scale", C.c_float), ("terrace", C.c_int32),
                ("terrace_scale", C.c_float), ("_pad0", C.c_uint32)]


PARTICLE_DTYPE = np.dtype([("sc", "<f4"), ("iters", "<i4"), ("position", "<f4", 2), ("velocity", "<f4", 2),
                           ("volume", "<f4"), ("_pad0", "<u4"), ("sediment", "<f4", 2), ("to_kill", "<u4"),
                           ("_pad1", "<u4")])
assert PARTICLE_DTYPE.itemsize == 48 and C.sizeof(ErosionData) == 96 and C.sizeof(MapSettingsData) == 96


def build(force=False):
    """Compile both oracle builds with the committed recipe (oracle/Makefile)."""
    targets = [os.path.join(_HERE, n) for n in ("libhg_oracle.so", "libhg_oracle_fma.so")]
    srcs = [os.path.join(_HERE, n) for n in ("hg_oracle.c", "hg_oracle.h", "Makefile")]
    stale = force or any(not os.path.exists(t) or os.path.getmtime(t) < max(os.path.getmtime(s) for s in srcs)
                         for t in targets)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return targets


_libs = {}


def lib(fma=False):
    """Load (building if needed) the oracle; fma=True is the contraction-on build
    used only to measure the oracle's own FP noise floor."""
    key = bool(fma)
    if key in _libs:
        return _libs[key]
    build()
    L = C.CDLL(os.path.join(_HERE, "libhg_oracle_fma.so" if fma else "libhg_oracle.so"))
    vp, i, u, f = C.c_void_p, C.c_int, C.c_uint32, C.c_float
    L.orc_create.restype = vp
    L.orc_create.argtypes = [i, i, u, i, f]
    L.orc_destroy.argtypes = [vp]
    for name in ("orc_gen_heightmap", "orc_dispatch_grid_rain", "orc_dispatch_grid"):
        getattr(L, name).argtypes = [vp]
    L.orc_dispatch_particle.argtypes = [vp, i]
    L.orc_step.argtypes = [vp, f, i]
    L.orc_pass.argtypes = [vp, i]
    L.orc_particle_pass.argtypes = [vp, i, i]
    L.orc_field.restype = C.POINTER(C.c_float)
    L.orc_field.argtypes = [vp, i]
    L.orc_particles.restype = vp
    L.orc_particles.argtypes = [vp]
    L.orc_erosion.restype = C.POINTER(ErosionData)
    L.orc_erosion.argtypes = [vp]
    L.orc_rain.restype = C.POINTER(RainData)
    L.orc_rain.argtypes = [vp]
    L.orc_map.restype = C.POINTER(MapSettingsData)
    L.orc_map.argtypes = [vp]
    L.orc_steps.restype = u
    L.orc_steps.argtypes = [vp]
    L.orc_set_steps.argtypes = [vp, u]
    L.orc_set_time.argtypes = [vp, f]
    for name in ("orc_atanf", "orc_expf", "orc_sinf"):
        getattr(L, name).restype = f
        getattr(L, name).argtypes = [f]
    L.orc_simplex.restype = f
    L.orc_simplex.argtypes = [f, f]
    L.orc_noised.argtypes = [f, f, C.POINTER(C.c_float)]
    _libs[key] = L
    return L


class World:
    """State::World::Textures + State::Settings of the oracle (src/state.hpp:10-76)."""

    def __init__(self, width, height=None, particle_count=0, erosion_type=0, seed=1234.5, fma=False):
        self.L = lib(fma)
        self.W = int(width)
        self.H = int(height if height is not None else width)
        self.particle_count = int(particle_count)
        self.erosion_type = int(erosion_type)
        self.h = self.L.orc_create(self.W, self.H, self.particle_count, self.erosion_type, seed)
        if not self.h:
            raise MemoryError("orc_create failed")

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def erosion(self):
        return self.L.orc_erosion(self.h).contents

    @property
    def rain(self):
        return self.L.orc_rain(self.h).contents

    @property
    def map(self):
        return self.L.orc_map(self.h).contents

    @property
    def steps(self):
        return self.L.orc_steps(self.h)

    @steps.setter
    def steps(self, v):
        self.L.orc_set_steps(self.h, int(v))

    def field(self, which):
        """numpy VIEW (H, W, 4) of the current read image of a field; re-fetch after a dispatch."""
        p = self.L.orc_field(self.h, which)
        return np.ctypeslib.as_array(p, shape=(self.H, self.W, 4))

    def get(self, which):
        return self.field(which).copy()

    def set(self, which, arr):
        self.field(which)[...] = np.asarray(arr, dtype=np.float32).reshape(self.H, self.W, 4)

    def particles(self):
        n = max(self.particle_count, 1)
        buf = (C.c_char * (48 * n)).from_address(self.L.orc_particles(self.h))
        return np.frombuffer(buf, dtype=PARTICLE_DTYPE, count=self.particle_count)

    def gen_heightmap(self):
        self.L.orc_gen_heightmap(self.h)

    def dispatch_grid_rain(self, time):
        self.L.orc_set_time(self.h, time)  # Textures::time (state.hpp:60)
        self.L.orc_dispatch_grid_rain(self.h)

    def dispatch_grid(self):
        self.L.orc_dispatch_grid(self.h)

    def dispatch_particle(self, time, should_rain=True):
        self.L.orc_set_time(self.h, time)
        self.L.orc_dispatch_particle(self.h, int(should_rain))

    def step(self, time, should_rain=True):
        self.L.orc_step(self.h, time, int(should_rain))

    def run_pass(self, which):
        self.L.orc_pass(self.h, which)

    def particle_pass(self, which, time, should_rain=True):
        self.L.orc_set_time(self.h, time)
        self.L.orc_particle_pass(self.h, which, int(should_rain))
