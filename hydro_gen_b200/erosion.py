"""Mirror of src/erosion.hpp: Programs, setup_shaders and the three per-step dispatch
functions, with the reference's signatures.  `world.time` plays State::World::Textures::time."""
from . import _lib

GRID, PARTICLES = _lib.HG_GRID, _lib.HG_PARTICLES   # Erosion::Programs::Erosion_type


class Programs:
    """Erosion::Programs (src/erosion.hpp:27-36): only the type survives; the GLSL programs
    are kernels inside libhydrogen_b200.so."""

    def __init__(self, type_):
        self.type = type_


def setup_shaders(type_, settings, data, particle_count=0):
    """Erosion::setup_shaders (src/erosion.cpp:21-74): binds the settings blocks to the
    world (the UBO bindings of :53-72) and pushes them once."""
    if type_ != data.ctx.erosion_type:
        raise _lib.HydrogenError(
            "erosion type does not match the world (gen_textures decides it from particle_count, as main.cpp:229-234 does)")
    settings._bind(data)
    return Programs(type_)


def dispatch_grid_rain(prog, data):
    """Erosion::dispatch_grid_rain (src/erosion.cpp:76-89)."""
    data.ctx.dispatch_grid_rain(data.time)


def dispatch_grid(prog, data):
    """Erosion::dispatch_grid (src/erosion.cpp:158-200)."""
    data.ctx.dispatch_grid()


def dispatch_particle(prog, data, should_rain):
    """Erosion::dispatch_particle (src/erosion.cpp:132-156)."""
    data.ctx.dispatch_particle(data.time, should_rain)
