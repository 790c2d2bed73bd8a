/* hg_oracle.h — CPU parity oracle for the erosion step.  TEST INFRASTRUCTURE:
 * see the header of hg_oracle.c.  Pinned against the reference's own shaders compiled for the CPU (oracle/refshader/). */
#ifndef HG_ORACLE_H
#define HG_ORACLE_H

#include <stdint.h>
#include "../include/hg_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* gl::Tex_pair: two RGBA32F images and a swap counter (src/shaderprogram.hpp:58-70) */
typedef struct orc_pair {
    float* tex[2];
    uint32_t idx_write, idx_read, cntr;
} orc_pair;

/* State::World::Textures + State::Settings (src/state.hpp:10-76) */
typedef struct orc_world {
    int W, H;
    float time;
    uint32_t particle_count;
    int erosion_type;           /* 0 grid, 1 particles (Erosion::Programs::Erosion_type) */
    uint32_t erosion_steps;     /* State::Program_state::erosion_steps */
    orc_pair heightmap, flux, velocity, sediment, thermal_c, thermal_d;
    hg_particle* particles;
    hg_erosion_data erosion;
    hg_rain_data rain;
    hg_map_settings_data map;
} orc_world;

enum { ORC_PASS_FLUX = 0, ORC_PASS_EROSION = 1, ORC_PASS_SEDIMENT = 2, ORC_PASS_THERMAL = 3, ORC_PASS_SMOOTH = 4 };

orc_world* orc_create(int W, int H, uint32_t particle_count, int erosion_type, float seed);
void orc_destroy(orc_world* w);

void orc_gen_heightmap(orc_world* w);
void orc_dispatch_grid_rain(orc_world* w);
void orc_dispatch_grid(orc_world* w);
void orc_dispatch_particle(orc_world* w, int should_rain);
void orc_step(orc_world* w, float time, int should_rain);
void orc_pass(orc_world* w, int pass);
void orc_particle_pass(orc_world* w, int which, int should_rain);

/* field = 0 H, 1 F, 2 V, 3 S, 4 TC, 5 TD: pointer to the current READ image (W*H*4 floats) */
float* orc_field(orc_world* w, int field);
hg_particle* orc_particles(orc_world* w);
hg_erosion_data* orc_erosion(orc_world* w);
hg_rain_data* orc_rain(orc_world* w);
hg_map_settings_data* orc_map(orc_world* w);
uint32_t orc_steps(orc_world* w);
void orc_set_steps(orc_world* w, uint32_t s);
void orc_set_time(orc_world* w, float t);

float orc_atanf(float x);
float orc_expf(float x);
float orc_sinf(float x);
float orc_simplex(float x, float y);
void orc_noised(float x, float y, float* out3);

#ifdef __cplusplus
}
#endif
#endif
