"""ctypes loader of libhydrogen_b200.so (the C ABI of include/hydrogen_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises, and
hg_create itself fails when no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HG_FMAD=1: the opt-in contracted build (-fmad=true; within the north star's tolerance of the reference instead of
# bit-identical, tests/test_fmad_build.py).  HG_B200_LIB: any other build of the same library (tuning aid).
LIB_PATH = os.environ.get("HG_B200_LIB") or os.path.join(_HERE, "libhydrogen_b200_fmad.so" if os.environ.get("HG_FMAD") == "1" else "libhydrogen_b200.so")

HG_OK, HG_ERR_INVALID, HG_ERR_CUDA, HG_ERR_STATE, HG_ERR_NO_DEVICE = range(5)
HG_GRID, HG_PARTICLES = 0, 1
FIELD_HEIGHTMAP, FIELD_FLUX, FIELD_VELOCITY, FIELD_SEDIMENT, FIELD_THERMAL_C, FIELD_THERMAL_D = range(6)
SCHEDULE_FUSED, SCHEDULE_PASSES = 0, 1
PASS_FLUX, PASS_EROSION, PASS_SEDIMENT, PASS_THERMAL, PASS_SMOOTH = range(5)
HALO_ROWS = 8
MAX_SLABS = 16


class ErosionData(C.Structure):
    """hg_erosion_data == Erosion_data (glsl/bindings.glsl:39-60), 96 bytes."""
    _fields_ = [("particle_count", C.c_uint32), ("Kc", C.c_float), ("Kalpha", C.c_float * 2),
                ("Kconv", C.c_float), ("_pad0", C.c_uint32), ("Ks", C.c_float * 2), ("Kd", C.c_float * 2),
                ("Ke", C.c_float), ("ENERGY_KEPT", C.c_float), ("Kspeed", C.c_float * 2), ("G", C.c_float),
                ("d_t", C.c_float), ("density", C.c_float), ("init_volume", C.c_float), ("friction", C.c_float),
                ("inertia", C.c_float), ("min_volume", C.c_float), ("min_velocity", C.c_float),
                ("ttl", C.c_uint32), ("_pad1", C.c_uint32)]


class RainData(C.Structure):
    """hg_rain_data == Rain_data (glsl/bindings.glsl:62-68), 20 bytes."""
    _fields_ = [("amount", C.c_float), ("mountain_thresh", C.c_float), ("mountain_multip", C.c_float),
                ("period", C.c_int32), ("drops", C.c_float)]


class MapSettingsData(C.Structure):
    """hg_map_settings_data == Map_settings_data (glsl/bindings.glsl:70-99), 96 bytes."""
    _fields_ = [("max_height", C.c_float), ("max_dirt", C.c_float), ("hmap_dims", C.c_int32 * 2),
                ("height_mult", C.c_float), ("water_lvl", C.c_float), ("seed", C.c_float),
                ("persistance", C.c_float), ("lacunarity", C.c_float), ("scale", C.c_float),
                ("redistribution", C.c_float), ("octaves", C.c_int32), ("fake_erosion", C.c_uint32),
                ("mask_round", C.c_uint32), ("mask_exp", C.c_uint32), ("mask_power", C.c_uint32),
                ("mask_slope", C.c_uint32), ("uplift", C.c_uint32), ("uplift_scale", C.c_float),
                ("domain_warp", C.c_int32), ("domain_warp_scale", C.c_float), ("terrace", C.c_int32),
                ("terrace_scale", C.c_float), ("_pad0", C.c_uint32)]


class SlabExport(C.Structure):
    """hg_slab_export: CUDA IPC handle + geometry of one rank's slab."""
    _fields_ = [("mem_handle", C.c_ubyte * 64), ("arena_bytes", C.c_uint64), ("row0", C.c_uint32),
                ("rows", C.c_uint32), ("map_w", C.c_uint32), ("map_h", C.c_uint32), ("device", C.c_int32),
                ("_pad", C.c_uint32)]


class SlabExportParticles(C.Structure):
    """hg_slab_export_particles_t: IPC handles of a droplet slab's texel images, droplet array and ownership bytes."""
    _fields_ = [("images_handle", C.c_ubyte * 64), ("droplets_handle", C.c_ubyte * 64), ("owners_handle", C.c_ubyte * 64),
                ("particle_count", C.c_uint32), ("_pad", C.c_uint32)]


assert C.sizeof(ErosionData) == 96 and C.sizeof(RainData) == 20 and C.sizeof(MapSettingsData) == 96

# every symbol include/hydrogen_b200.h declares: (restype, argtypes)
_vp, _i, _u, _f = C.c_void_p, C.c_int, C.c_uint32, C.c_float
_fp = C.POINTER(C.c_float)
SYMBOLS = {
    "hg_create": (_vp, [_u, _u, _u, _i, _i]),
    "hg_create_slab": (_vp, [_u, _u, _u, _u, _u, _i, _i]),
    "hg_destroy": (None, [_vp]),
    "hg_last_error": (C.c_char_p, []),
    "hg_version": (C.c_char_p, []),
    "hg_set_erosion": (_i, [_vp, C.POINTER(ErosionData)]),
    "hg_set_rain": (_i, [_vp, C.POINTER(RainData)]),
    "hg_set_map": (_i, [_vp, C.POINTER(MapSettingsData)]),
    "hg_get_erosion": (_i, [_vp, C.POINTER(ErosionData)]),
    "hg_get_rain": (_i, [_vp, C.POINTER(RainData)]),
    "hg_get_map": (_i, [_vp, C.POINTER(MapSettingsData)]),
    "hg_set_schedule": (_i, [_vp, _i]),
    "hg_gen_heightmap": (_i, [_vp]),
    "hg_dispatch_grid_rain": (_i, [_vp, _f]),
    "hg_dispatch_grid": (_i, [_vp]),
    "hg_dispatch_particle": (_i, [_vp, _f, _i]),
    "hg_dispatch_pass": (_i, [_vp, _i]),
    "hg_dispatch_particle_pass": (_i, [_vp, _i, _f, _i]),
    "hg_run": (_i, [_vp, _u, _f, _f, _i]),
    "hg_get_steps": (_i, [_vp, C.POINTER(_u)]),
    "hg_set_steps": (_i, [_vp, _u]),
    "hg_upload": (_i, [_vp, _i, _vp]),
    "hg_download": (_i, [_vp, _i, _vp]),
    "hg_upload_async": (_i, [_vp, _i, _vp]),
    "hg_download_async": (_i, [_vp, _i, _vp]),
    "hg_upload_particles": (_i, [_vp, _vp, _u]),
    "hg_download_particles": (_i, [_vp, _vp, _u]),
    "hg_step_host_async": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hg_checkpoint_save": (_i, [_vp, C.c_char_p]),
    "hg_checkpoint_load": (_i, [_vp, C.c_char_p]),
    "hg_host_alloc": (_vp, [C.c_size_t]),
    "hg_host_free": (None, [_vp]),
    "hg_mass": (_i, [_vp, C.POINTER(C.c_double)]),
    "hg_sync": (_i, [_vp]),
    "hg_set_stream": (_i, [_vp, _vp]),
    "hg_get_stream": (_vp, [_vp]),
    "hg_timer_start": (_i, [_vp]),
    "hg_timer_stop": (_i, [_vp, _fp]),
    "hg_profile_fused": (_i, [_vp, _u, _fp]),
    "hg_run_profiled": (_i, [_vp, _u, _f, _f, _i, _fp, _fp]),
    "hg_launch_count": (C.c_uint64, [_vp]),
    "hg_far_fetch_count": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "hg_slab_export_handle": (_i, [_vp, C.POINTER(SlabExport)]),
    "hg_slab_connect": (_i, [_vp, C.POINTER(SlabExport), _i, _i]),
    "hg_slab_connect_local": (_i, [_vp, C.POINTER(_vp), _i, _i]),
    "hg_slab_set_ghost": (_i, [_vp, _i, _i, _vp]),
    "hg_slab_errors": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "hg_slab_refresh_halo": (_i, [_vp]),
    "hg_slab_export_particles": (_i, [_vp, C.POINTER(SlabExportParticles)]),
    "hg_slab_connect_particles": (_i, [_vp, C.POINTER(SlabExportParticles), _i, _i]),
    "hg_slab_particle_owners": (_i, [_vp, _vp, _u]),
    "hg_register_gl": (_i, [_vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "hg_publish_gl": (_i, [_vp, _i]),
    "hg_unregister_gl": (_i, [_vp]),
    "hg_pack_device": (_i, [_vp, _i, _vp]),
}

_lib = None


class HydrogenError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HydrogenError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C hydro_gen_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)        # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != HG_OK:
        raise HydrogenError(f"hydrogen_b200 error {rc}: {load().hg_last_error().decode()}")
