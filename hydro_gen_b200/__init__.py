"""hydro_gen_b200 — B200-native erosion step of ger0/hydro-gen behind the reference's own
interface.  `state` mirrors src/state.hpp (Settings, World.Textures, gen_textures,
gen_heightmap), `erosion` mirrors src/erosion.hpp (setup_shaders, dispatch_grid_rain,
dispatch_grid, dispatch_particle); both are thin host code over the C ABI
(include/hydrogen_b200.h -> libhydrogen_b200.so, hand-written sm_100a CUDA).
No CPU fallback exists: importing works anywhere, creating a context needs a GPU."""
from . import _lib
from ._lib import (FIELD_FLUX, FIELD_HEIGHTMAP, FIELD_SEDIMENT, FIELD_THERMAL_C, FIELD_THERMAL_D, FIELD_VELOCITY,
                   HG_GRID, HG_PARTICLES, SCHEDULE_FUSED, SCHEDULE_PASSES, ErosionData, HydrogenError,
                   MapSettingsData, RainData)
from .context import PARTICLE_DTYPE, Context, PinnedBuffer
from . import state, erosion, config, slabs

__all__ = ["state", "erosion", "config", "slabs", "Context", "PinnedBuffer", "ErosionData", "RainData", "MapSettingsData",
           "HydrogenError", "PARTICLE_DTYPE"]
