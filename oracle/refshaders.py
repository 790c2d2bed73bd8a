"""Driver of oracle/_ref/libhg_refshaders.so — the REFERENCE'S OWN compute shaders compiled for the
CPU (oracle/refshader/build_ref.py).  TEST INFRASTRUCTURE: used by tests/test_refshaders.py to validate
the oracle restatement against the reference's shader text, and by tests/golden/make_golden.py.

This file restates only the reference's HOST side of a step: which texture of which ping-pong pair is
bound to which shader variable, and when a pair swaps (src/erosion.cpp:76-200, Tex_pair:
src/shaderprogram.cpp:51-82).  The arithmetic is the shaders' own."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libhg_refshaders.so")
REFERENCE_GLSL = os.environ.get("HG_REFERENCE_GLSL", "/root/reference/glsl")


def available(build=True):
    """True if the library exists; builds it first when the reference sources are present."""
    if build and os.path.isdir(REFERENCE_GLSL):
        src = os.path.join(_HERE, "refshader")
        newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
            import subprocess, sys
            subprocess.run([sys.executable, os.path.join(src, "build_ref.py")], check=True)
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_bind.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_set_uniform.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int]
        L.ref_run.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.ref_run1d.argtypes = [C.c_char_p, C.c_int]
        _lib = L
    return _lib


class TexPair:
    """gl::Tex_pair (src/shaderprogram.cpp:51-82): two RGBA32F textures, read index = swap count mod 2."""

    def __init__(self, n):
        self.tex = [np.zeros((n, n, 4), np.float32), np.zeros((n, n, 4), np.float32)]
        self.cntr = 0

    def swap(self):
        self.cntr += 1

    @property
    def read(self):
        return self.tex[self.cntr % 2]

    @property
    def write(self):
        return self.tex[(self.cntr + 1) % 2]


class RefWorld:
    """State::World::Textures (src/state.cpp:3-44) + the settings, stepped by the reference's shaders."""
    FIELDS = ("heightmap", "flux", "velocity", "sediment", "thermal_c", "thermal_d")

    def __init__(self, n, erosion, rain, map_settings, particle_count=0):
        self.n, self.L = n, lib()
        for f in self.FIELDS:
            setattr(self, f, TexPair(n))
        # droplet mode (src/state.cpp:23-33): r32ui lock map and the Particle SSBO, zero-initialised
        self.particle_count = particle_count
        self.lockmap = np.zeros((n, n), np.uint32)
        self.particle_buffer = np.zeros(max(particle_count, 1) * 48, np.uint8)
        self.erosion, self.rain, self.map = erosion, rain, map_settings     # ctypes images of the std140 blocks
        self.time = 0.0

    # -- Compute_program::bind_texture / bind_image (by variable name), set_uniform, run (erosion.cpp:91-101)
    def _bind(self, shader, **images):
        for name, arr in images.items():
            assert arr.flags.c_contiguous and arr.dtype == np.float32
            if self.L.ref_bind(shader.encode(), name.encode(), arr.ctypes.data, self.n, self.n) != 0:
                raise RuntimeError(f"{shader} has no image variable {name}")

    def _uniform(self, shader, name, value):
        if self.L.ref_set_uniform(shader.encode(), name.encode(), C.byref(value), C.sizeof(value)) != 0:
            raise RuntimeError(f"{shader}: uniform {name} missing or of another size")

    def _run(self, shader):
        if self.L.ref_run(shader.encode(), self.n, self.n) != 0:
            raise RuntimeError(shader)

    def gen_heightmap(self):                                 # src/state.cpp:116-147 (State::World::gen_heightmap)
        self._uniform("heightmap", "cfg", self.map)
        self._bind("heightmap", dest_heightmap=self.heightmap.write, dest_vel=self.velocity.write,
                   dest_flux=self.flux.write, dest_sediment=self.sediment.write)
        self._run("heightmap")
        self.heightmap.swap(); self.velocity.swap(); self.flux.swap(); self.sediment.swap()

    def dispatch_grid_rain(self, time):                      # src/erosion.cpp:76-89
        self._uniform("rain", "time", C.c_float(time))
        self._uniform("rain", "set", self.rain)
        self._uniform("rain", "map_set", self.map)
        self._bind("rain", heightmap=self.heightmap.read, out_heightmap=self.heightmap.write)
        self._run("rain")
        self.heightmap.swap()

    def run_thermal_erosion(self):                           # src/erosion.cpp:103-122
        for layer in range(2):
            self._uniform("thermal_erosion", "t_layer", C.c_int(layer))      # set once per program in setup_shaders, src/erosion.cpp:56-58
            self._uniform("thermal_erosion", "set", self.erosion)
            self._bind("thermal_erosion", heightmap=self.heightmap.read, out_thflux_c=self.thermal_c.write, out_thflux_d=self.thermal_d.write)
            self._run("thermal_erosion")
            self.thermal_c.swap(); self.thermal_d.swap()
            self._uniform("thermal_transport", "t_layer", C.c_int(layer))
            self._bind("thermal_transport", heightmap=self.heightmap.read, out_heightmap=self.heightmap.write,
                       thflux_c=self.thermal_c.read, thflux_d=self.thermal_d.read)
            self._run("thermal_transport")
            self.heightmap.swap()

    def pass_flux(self):                                     # src/erosion.cpp:158-169
        self._uniform("hydro_flux", "set", self.erosion)
        self._bind("hydro_flux", heightmap=self.heightmap.read, fluxmap=self.flux.read, velocitymap=self.velocity.read,
                   out_heightmap=self.heightmap.write, out_fluxmap=self.flux.write, out_velocitymap=self.velocity.write)
        self._run("hydro_flux")
        self.heightmap.swap(); self.flux.swap(); self.velocity.swap()

    def pass_erosion(self):                                  # src/erosion.cpp:171-179
        self._uniform("hydro_erosion", "set", self.erosion)
        self._bind("hydro_erosion", heightmap=self.heightmap.read, sedimap=self.sediment.read, velocitymap=self.velocity.read,
                   out_heightmap=self.heightmap.write, out_sedimap=self.sediment.write)
        self._run("hydro_erosion")
        self.heightmap.swap(); self.sediment.swap()

    def pass_sediment(self):                                 # src/erosion.cpp:181-190
        self._uniform("sediment_transport", "set", self.erosion)
        self._bind("sediment_transport", heightmap=self.heightmap.read, velocitymap=self.velocity.read, sedimap=self.sediment.read,
                   out_heightmap=self.heightmap.write, out_sedimap=self.sediment.write)
        self._run("sediment_transport")
        self.heightmap.swap(); self.sediment.swap()

    def pass_smooth(self):                                   # src/erosion.cpp:194-200 (momentmap unbound in grid mode)
        self._uniform("smoothing", "set", self.erosion)
        self._bind("smoothing", heightmap=self.heightmap.read, out_heightmap=self.heightmap.write)
        self._run("smoothing")
        self.heightmap.swap()

    def _run1d(self, shader, count):
        count = count // 64 * 64          # run_particles dispatches particle_count / (8 * 8) work groups of 64 (src/erosion.cpp:124-130)
        if self.L.ref_run1d(shader.encode(), count) != 0:
            raise RuntimeError(shader)

    def _bind_buffer(self, shader, block, arr):
        if self.L.ref_bind(shader.encode(), block.encode(), arr.ctypes.data, 0, 0) != 0:
            raise RuntimeError(f"{shader} has no buffer block {block}")

    def particle_move(self, time, should_rain):              # src/erosion.cpp:132-138
        self._uniform("particle", "time", C.c_float(time))
        self._uniform("particle", "should_rain", C.c_uint32(1 if should_rain else 0))
        self._uniform("particle", "set", self.erosion)
        self._uniform("particle", "map_set", self.map)
        self._bind("particle", heightmap=self.heightmap.read, momentmap=self.velocity.read)
        self._bind_buffer("particle", "ParticleBuffer", self.particle_buffer)
        self._run1d("particle", self.particle_count)

    def particle_erode(self):                                # src/erosion.cpp:140-144
        self._uniform("particle_erosion", "set", self.erosion)
        if self.L.ref_bind(b"particle_erosion", b"lockmap", self.lockmap.ctypes.data, self.n, self.n) != 0:
            raise RuntimeError("lockmap")
        self._bind("particle_erosion", heightmap=self.heightmap.read, momentmap=self.velocity.read)
        self._bind_buffer("particle_erosion", "ParticleBuffer", self.particle_buffer)
        self._run1d("particle_erosion", self.particle_count)

    def dispatch_particle(self, time, should_rain=True):     # src/erosion.cpp:132-156
        self.particle_move(time, should_rain)
        self.particle_erode()
        self.run_thermal_erosion()
        self._uniform("smoothing", "set", self.erosion)
        self._bind("smoothing", heightmap=self.heightmap.read, momentmap=self.velocity.read,
                   out_heightmap=self.heightmap.write, out_momentmap=self.velocity.write)
        self._run("smoothing")
        self.heightmap.swap(); self.velocity.swap()

    PASSES = ("pass_flux", "pass_erosion", "pass_sediment", "run_thermal_erosion", "pass_smooth")

    def dispatch_grid(self):                                 # src/erosion.cpp:158-200
        for p in self.PASSES:
            getattr(self, p)()

    def step(self, steps, time):                             # src/main.cpp:310-321
        if steps % self.rain.period == 0:
            self.dispatch_grid_rain(time)
        self.dispatch_grid()
