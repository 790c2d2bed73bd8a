"""Mirror of src/state.hpp + src/settings.hpp: the settings structs with their defaults
and push_data(), and State::World::Textures with gen_textures / gen_heightmap /
delete_textures.  Same names, argument meaning and call order as the reference; the
GL objects are replaced by one hg_ctx handle."""
from . import _lib
from ._lib import ErosionData, MapSettingsData, RainData
from .context import Context

MAX_HEIGHT = 256.0      # src/settings.hpp:10
WATER_HEIGHT = 96.0     # src/settings.hpp:11


def default_erosion(is_particle=False, particle_count=0):
    """State::setup_settings defaults, src/state.cpp:61-92."""
    e = ErosionData()
    e.Kc = 0.2
    e.Kalpha[0], e.Kalpha[1] = 1.3, 0.6
    e.Kconv = 0.001
    e.Ks[0], e.Ks[1] = 0.03, 0.09
    e.Kd[0], e.Kd[1] = 0.01, 0.03
    e.Ke = 0.03
    if is_particle:
        e.particle_count = particle_count
        e.Kspeed[0], e.Kspeed[1] = 0.002, 0.008
        e.G, e.d_t, e.density, e.init_volume = 9.81, 0.25, 1.0, 1.0
        e.friction, e.inertia, e.min_volume, e.min_velocity, e.ttl = 0.2, 1.0, 0.0, 0.001, 15000
    else:
        e.ENERGY_KEPT = 1.0
        e.Kspeed[0], e.Kspeed[1] = 0.5, 2.0
        e.G, e.d_t = 1.0, 0.001
    return e


def default_rain():
    """Rain_settings, src/settings.hpp:15-21."""
    return RainData(0.01, 0.55, 0.05, 512, 0.02)


def default_map(seed=0.0):
    """Map_settings, src/settings.hpp:29-55; `seed` is rand()-derived in the reference."""
    m = MapSettingsData()
    m.max_height, m.max_dirt = MAX_HEIGHT, 2.0
    m.hmap_dims[0], m.hmap_dims[1] = 1024, 1024
    m.height_mult, m.water_lvl, m.seed = 1.0, WATER_HEIGHT, seed
    m.persistance, m.lacunarity, m.scale, m.redistribution, m.octaves = 0.44, 2.0, 0.00075, 1.0, 8
    m.mask_round, m.mask_exp, m.mask_power, m.mask_slope = 0, 1, 1, 0
    m.uplift, m.uplift_scale = 0, 1.16
    m.domain_warp, m.domain_warp_scale, m.terrace, m.terrace_scale = 1, 100.0, 0, 0.5
    return m


class _Block:
    """Erosion_settings / Rain_settings / Map_settings: host copy + push_data()."""

    def __init__(self, data, setter):
        self.data = data
        self._setter = setter
        self._world = None

    def push_data(self):
        if self._world is None or self._world.ctx is None:
            return          # nothing bound yet: gen_textures pushes all three
        getattr(self._world.ctx, self._setter)(self.data)


class Settings:
    """State::Settings (src/state.hpp:10-14)."""

    def __init__(self, erosion, rain, map_):
        self.erosion = _Block(erosion, "set_erosion")
        self.rain = _Block(rain, "set_rain")
        self.map = _Block(map_, "set_map")

    def _bind(self, world):
        for b in (self.erosion, self.rain, self.map):
            b._world = world
            b.push_data()


def setup_settings(is_particle=False, particle_count=0, seed=0.0):
    """State::setup_settings (src/state.cpp:57-106)."""
    return Settings(default_erosion(is_particle, particle_count), default_rain(), default_map(seed))


def delete_settings(settings):
    """State::delete_settings (src/state.cpp:108-113): nothing device-side to free."""
    for b in (settings.erosion, settings.rain, settings.map):
        b._world = None


class World:
    class Textures:
        """State::World::Textures (src/state.hpp:59-76): time, map_size, particle_count and
        the fields (here: one device context holding the SoA planes)."""

        def __init__(self, map_size, particle_count, device=0, map_height=None, row0=None, rows=None):
            self.time = 0.0
            self.map_size = int(map_size)
            self.particle_count = int(particle_count)
            etype = _lib.HG_PARTICLES if particle_count else _lib.HG_GRID
            self.ctx = Context(map_size, map_height, particle_count, etype, device, row0, rows)

        # the renderer's inputs, src/rendering.cpp:104-105
        def heightmap(self):
            return self.ctx.download(_lib.FIELD_HEIGHTMAP)

        def sediment(self):
            return self.ctx.download(_lib.FIELD_SEDIMENT)

    @staticmethod
    def gen_textures(size, particle_count, device=0, **kw):
        """State::World::gen_textures (src/state.cpp:3-44)."""
        return World.Textures(size, particle_count, device, **kw)

    @staticmethod
    def delete_textures(data):
        """State::World::delete_textures (src/state.cpp:46-55)."""
        data.ctx.close()
        data.ctx = None

    @staticmethod
    def gen_heightmap(settings, world_data, program=None):
        """State::World::gen_heightmap (src/state.cpp:116-147): pushes the map settings and
        runs the heightmap kernel; `program` (the GL Compute_program) is accepted and ignored."""
        settings._bind(world_data)
        settings.map.push_data()
        world_data.ctx.gen_heightmap()
