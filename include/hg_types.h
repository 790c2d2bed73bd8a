/* hg_types.h — parameter and droplet structs of the erosion path, plain C.
 *
 * Byte-for-byte the layout the reference shares between C++ and GLSL through
 * glsl/bindings.glsl:39-111 (every member is `alignas(sizeof(member))`, so a
 * vec2 sits on an 8-byte boundary and each struct is padded to 8).  The sizes
 * and offsets are asserted below; a caller that already holds the reference's
 * `Erosion_data` / `Rain_data` / `Map_settings_data` / `Particle` can pass a
 * pointer to it unchanged.
 *
 * Defaults (hg_default_*) restate src/state.cpp:61-92 (erosion, per mode) and
 * src/settings.hpp:15-55 (rain, map).
 */
#ifndef HG_TYPES_H
#define HG_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_SED_LAYERS 2        /* bindings.glsl:5  (0 = rock, 1 = dirt) */
#define HG_WRKGRP 8            /* bindings.glsl:9-10; map sizes must be multiples of it */
#define HG_OOB_HEIGHT 999999999999.0f /* hydro_flux.glsl:37, thermal_erosion.glsl:24 */

/* bindings.glsl:39-60 — UBO binding 1 */
typedef struct hg_erosion_data {
    uint32_t particle_count;   /*  0 */
    float    Kc;               /*  4 sediment capacity constant */
    float    Kalpha[2];        /*  8 talus angle per layer, radians */
    float    Kconv;            /* 16 rock-sediment -> dirt-sediment rate */
    uint32_t _pad0;            /* 20 */
    float    Ks[2];            /* 24 dissolving constant per layer */
    float    Kd[2];            /* 32 deposition constant per layer */
    float    Ke;               /* 40 evaporation */
    float    ENERGY_KEPT;      /* 44 */
    float    Kspeed[2];        /* 48 thermal slippage speed per layer */
    float    G;                /* 56 */
    float    d_t;              /* 60 */
    float    density;          /* 64 (unused by the shaders) */
    float    init_volume;      /* 68 */
    float    friction;         /* 72 */
    float    inertia;          /* 76 */
    float    min_volume;       /* 80 */
    float    min_velocity;     /* 84 */
    uint32_t ttl;              /* 88 */
    uint32_t _pad1;            /* 92 */
} hg_erosion_data;

/* bindings.glsl:62-68 — UBO binding 3 */
typedef struct hg_rain_data {
    float   amount;
    float   mountain_thresh;
    float   mountain_multip;
    int32_t period;
    float   drops;
} hg_rain_data;

/* bindings.glsl:70-99 — UBO binding 2 */
typedef struct hg_map_settings_data {
    float    max_height;       /*  0 */
    float    max_dirt;         /*  4 */
    int32_t  hmap_dims[2];     /*  8 */
    float    height_mult;      /* 16 */
    float    water_lvl;        /* 20 */
    float    seed;             /* 24 */
    float    persistance;      /* 28 */
    float    lacunarity;       /* 32 */
    float    scale;            /* 36 */
    float    redistribution;   /* 40 */
    int32_t  octaves;          /* 44 */
    uint32_t fake_erosion;     /* 48 */
    uint32_t mask_round;       /* 52 */
    uint32_t mask_exp;         /* 56 */
    uint32_t mask_power;       /* 60 */
    uint32_t mask_slope;       /* 64 */
    uint32_t uplift;           /* 68 */
    float    uplift_scale;     /* 72 */
    int32_t  domain_warp;      /* 76 */
    float    domain_warp_scale;/* 80 */
    int32_t  terrace;          /* 84 */
    float    terrace_scale;    /* 88 */
    uint32_t _pad0;            /* 92 */
} hg_map_settings_data;

/* bindings.glsl:101-111 — SSBO binding 4 element (std430 view: 4-byte bool) */
typedef struct hg_particle {
    float    sc;               /*  0 sediment capacity at the droplet */
    int32_t  iters;            /*  4 */
    float    position[2];      /*  8 */
    float    velocity[2];      /* 16 */
    float    volume;           /* 24 */
    uint32_t _pad0;            /* 28 */
    float    sediment[2];      /* 32 */
    uint32_t to_kill;          /* 40 */
    uint32_t _pad1;            /* 44 */
} hg_particle;

#ifdef __cplusplus
#define HG_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define HG_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
HG_STATIC_ASSERT(sizeof(hg_erosion_data) == 96, "Erosion_data is 96 bytes");
HG_STATIC_ASSERT(offsetof(hg_erosion_data, Ks) == 24, "Ks at 24");
HG_STATIC_ASSERT(offsetof(hg_erosion_data, Kspeed) == 48, "Kspeed at 48");
HG_STATIC_ASSERT(offsetof(hg_erosion_data, ttl) == 88, "ttl at 88");
HG_STATIC_ASSERT(sizeof(hg_rain_data) == 20, "Rain_data is 20 bytes");
HG_STATIC_ASSERT(sizeof(hg_map_settings_data) == 96, "Map_settings_data is 96 bytes");
HG_STATIC_ASSERT(offsetof(hg_map_settings_data, domain_warp) == 76, "domain_warp at 76");
HG_STATIC_ASSERT(sizeof(hg_particle) == 48, "Particle is 48 bytes");
HG_STATIC_ASSERT(offsetof(hg_particle, sediment) == 32, "sediment at 32");
HG_STATIC_ASSERT(offsetof(hg_particle, to_kill) == 40, "to_kill at 40");

/* src/state.cpp:81-92 (grid) and :61-79 (particle); unset members are zero
 * exactly as the reference's designated initialisers leave them. */
static inline hg_erosion_data hg_default_erosion(int is_particle, uint32_t particle_count) {
    hg_erosion_data e;
    for (size_t i = 0; i < sizeof(e); ++i) ((unsigned char*)&e)[i] = 0;
    e.Kc = 0.2f;
    e.Kalpha[0] = 1.3f;  e.Kalpha[1] = 0.6f;
    e.Kconv = 0.001f;
    e.Ks[0] = 0.03f;     e.Ks[1] = 0.09f;
    e.Kd[0] = 0.01f;     e.Kd[1] = 0.03f;
    e.Ke = 0.03f;
    if (is_particle) {
        e.particle_count = particle_count;
        e.Kspeed[0] = 0.002f; e.Kspeed[1] = 0.008f;
        e.G = 9.81f;
        e.d_t = 0.25f;
        e.density = 1.0f;
        e.init_volume = 1.0f;
        e.friction = 0.2f;
        e.inertia = 1.0f;
        e.min_volume = 0.0f;
        e.min_velocity = 0.001f;
        e.ttl = 15000u;
    } else {
        e.ENERGY_KEPT = 1.0f;
        e.Kspeed[0] = 0.5f;  e.Kspeed[1] = 2.0f;
        e.G = 1.0f;
        e.d_t = 0.001f;
    }
    return e;
}

/* src/settings.hpp:15-21 */
static inline hg_rain_data hg_default_rain(void) {
    hg_rain_data r;
    r.amount = 0.01f;
    r.mountain_thresh = 0.55f;
    r.mountain_multip = 0.05f;
    r.period = 512;
    r.drops = 0.02f;
    return r;
}

/* src/settings.hpp:29-55.  The reference draws `seed` from rand() seeded with
 * wall-clock time (main.cpp:182); here it is an explicit input.  hmap_dims
 * stays (1024,1024) whatever the map size, exactly as the reference leaves it
 * (only the droplet spawn/kill box reads it, particle.glsl:77-80,109-113). */
static inline hg_map_settings_data hg_default_map(float seed) {
    hg_map_settings_data m;
    for (size_t i = 0; i < sizeof(m); ++i) ((unsigned char*)&m)[i] = 0;
    m.max_height = 256.0f;
    m.max_dirt = 2.0f;
    m.hmap_dims[0] = 1024; m.hmap_dims[1] = 1024;
    m.height_mult = 1.0f;
    m.water_lvl = 96.0f;
    m.seed = seed;
    m.persistance = 0.44f;
    m.lacunarity = 2.0f;
    m.scale = 0.00075f;
    m.redistribution = 1.0f;
    m.octaves = 8;
    m.mask_round = 0; m.mask_exp = 1; m.mask_power = 1; m.mask_slope = 0;
    m.uplift = 0;
    m.uplift_scale = 1.16f;
    m.domain_warp = 1;
    m.domain_warp_scale = 100.0f;
    m.terrace = 0;
    m.terrace_scale = 0.5f;
    return m;
}

#ifdef __cplusplus
}
#endif
#endif /* HG_TYPES_H */
