// hydrogen_erosion.hpp — header-only C++ mirror of the reference's erosion interface
// (src/erosion.hpp:6-49, src/state.hpp:8-87, src/settings.hpp:8-69) on top of the C ABI
// of include/hydrogen_b200.h.  Same namespaces, names, argument order and call sequence, so
// the erosion part of src/main.cpp (lines 252-276 set-up, 310-324 per-step dispatch) compiles
// against it unchanged apart from the include; INTEGRATION.md shows the patch.
//
// What differs, by necessity:
//  * Textures holds one hg_ctx* instead of GL texture names; fields are read back with
//    World::download() (the reference never reads them back; its renderer samples them).
//  * Compute_program is an empty tag (the kernels live in libhydrogen_b200.so).
//  * Map_settings.data.seed is not drawn from rand() at static-init time; set it before
//    gen_heightmap (the reference's value is wall-clock dependent, src/main.cpp:182).
//  * Failures print hg_last_error() and exit(1) — the reference's convention for fatal
//    shader/GL errors (src/shaderprogram.cpp:132-158); there is no CPU fallback.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/hydrogen_b200.h"

using u32 = std::uint32_t;
#ifndef SED_LAYERS
#define SED_LAYERS HG_SED_LAYERS          // glsl/bindings.glsl:5
#endif
// glsl/bindings.glsl:39-111 — identical layouts (static-asserted in hg_types.h)
using Erosion_data = hg_erosion_data;
using Rain_data = hg_rain_data;
using Map_settings_data = hg_map_settings_data;
using Particle = hg_particle;

struct Compute_program {                  // src/shaderprogram.hpp:84-116: nothing to compile here
    explicit Compute_program(const char* = "") {}
};

namespace hydrogen_detail {
[[noreturn]] inline void fail(const char* what) {
    std::fprintf(stderr, "[hydrogen_b200] %s: %s\n", what, hg_last_error());
    std::exit(1);
}
inline void check(int rc, const char* what) { if (rc != HG_OK) fail(what); }
}  // namespace hydrogen_detail

namespace State {

constexpr float MAX_HEIGHT = 256.f;       // src/settings.hpp:10
constexpr float WATER_HEIGHT = 96.f;      // src/settings.hpp:11

// A settings block = host copy + push_data(), as src/settings.hpp:13-66.  `ctx` is bound by
// Erosion::setup_shaders / World::gen_heightmap (the reference binds its UBOs there,
// src/erosion.cpp:62-72); push_data() before that only updates the host copy.
struct Rain_settings {
    hg_ctx* ctx = nullptr;
    Rain_data data = hg_default_rain();
    void push_data() { if (ctx) hydrogen_detail::check(hg_set_rain(ctx, &data), "Rain_settings::push_data"); }
};
struct Map_settings {
    hg_ctx* ctx = nullptr;
    Map_settings_data data = hg_default_map(0.0f);
    void push_data() { if (ctx) hydrogen_detail::check(hg_set_map(ctx, &data), "Map_settings::push_data"); }
};
struct Erosion_settings {
    hg_ctx* ctx = nullptr;
    Erosion_data data = hg_default_erosion(0, 0);
    void push_data() { if (ctx) hydrogen_detail::check(hg_set_erosion(ctx, &data), "Erosion_settings::push_data"); }
};

struct Settings {                         // src/state.hpp:10-14
    Erosion_settings erosion;
    Rain_settings rain;
    Map_settings map;
};

// src/state.cpp:57-106
inline Settings setup_settings(bool is_particle = false, u32 particle_count = 0) {
    Settings s;
    s.erosion.data = hg_default_erosion(is_particle ? 1 : 0, particle_count);
    return s;
}
inline void delete_settings(Settings& s) { s.erosion.ctx = nullptr; s.rain.ctx = nullptr; s.map.ctx = nullptr; }

struct Program_state {                    // the members of src/state.hpp:19-39 the erosion loop uses
    bool should_rain = true;
    bool should_erode = true;
    u32 erosion_steps = 0;
};

namespace World {

struct Textures {                         // src/state.hpp:59-76
    float time = 0.f;
    u32 map_size = 0;
    u32 particle_count = 0;
    hg_ctx* ctx = nullptr;                // the SoA planes behind heightmap/flux/velocity/sediment/thermal_*
    int device = 0;
};

// src/state.cpp:3-44.  particle_count != 0 selects droplet mode exactly as main.cpp:229-234
// ties the count to the erosion type.
inline Textures gen_textures(const u32 size, const u32 particle_count, int device = 0) {
    Textures t;
    t.map_size = size;
    t.particle_count = particle_count;
    t.device = device;
    t.ctx = hg_create(size, size, particle_count, particle_count ? HG_PARTICLES : HG_GRID, device);
    if (!t.ctx) hydrogen_detail::fail("gen_textures");
    return t;
}
inline void delete_textures(Textures& data) { hg_destroy(data.ctx); data.ctx = nullptr; }   // src/state.cpp:46-55

inline void bind(Settings& s, Textures& w) {
    s.erosion.ctx = s.rain.ctx = s.map.ctx = w.ctx;
    s.erosion.push_data(); s.rain.push_data(); s.map.push_data();
}

// src/state.cpp:116-147
inline void gen_heightmap(Settings& settings, Textures& world_data, Compute_program&) {
    bind(settings, world_data);
    hydrogen_detail::check(hg_gen_heightmap(world_data.ctx), "gen_heightmap");
}

// Host copy of a field in the reference's texture format (RGBA32F, [y][x][4]); the read
// textures the renderer binds are HG_FIELD_HEIGHTMAP and HG_FIELD_SEDIMENT (rendering.cpp:104-105).
inline std::vector<float> download(Textures& w, int field) {
    std::vector<float> out((size_t)w.map_size * w.map_size * 4);
    hydrogen_detail::check(hg_download(w.ctx, field, out.data()), "download");
    return out;
}

// CUDA-GL interop with the reference's renderer (library built with -DHG_WITH_GL): the renderer keeps its two
// gl::Tex_pair objects `heightmap` and `sediment` (src/state.hpp:62-67) and samples their read textures
// (src/rendering.cpp:103-104).  register_gl once after gen_textures, with the GL names of both textures of each pair
// (Tex_pair::t1.texture, t2.texture); publish_gl after the frame's erosion steps, with the pairs' read indices
// (Tex_pair::cntr % 2, src/shaderprogram.cpp:51-60): the fields land in the textures the next draw samples.
inline void register_gl(Textures& w, const unsigned heightmap_tex[2], const unsigned sediment_tex[2]) {
    hydrogen_detail::check(hg_register_gl(w.ctx, heightmap_tex, sediment_tex), "register_gl");
}
inline void publish_gl(Textures& w, int heightmap_read_idx, int sediment_read_idx) {
    hydrogen_detail::check(hg_publish_gl(w.ctx, (heightmap_read_idx & 1) | ((sediment_read_idx & 1) << 1)), "publish_gl");
    hydrogen_detail::check(hg_sync(w.ctx), "publish_gl: sync");      // GL samples after the copies have landed
}

}  // namespace World
}  // namespace State

namespace Erosion {

struct Programs {                         // src/erosion.hpp:27-36
    enum Erosion_type { GRID, PARTICLES } type;
};

// src/erosion.cpp:21-74
inline Programs* setup_shaders(Programs::Erosion_type type, State::Settings& set, State::World::Textures& data, u32 /*particle_count*/) {
    if ((type == Programs::PARTICLES) != (data.particle_count != 0)) {
        std::fprintf(stderr, "[hydrogen_b200] setup_shaders: erosion type does not match the textures' particle_count\n");
        std::exit(1);
    }
    State::World::bind(set, data);
    return new Programs{type};
}
// src/erosion.cpp:76-89
inline void dispatch_grid_rain(Programs&, State::World::Textures& data) {
    hydrogen_detail::check(hg_dispatch_grid_rain(data.ctx, data.time), "dispatch_grid_rain");
}
// src/erosion.cpp:158-200
inline void dispatch_grid(Programs&, State::World::Textures& data) {
    hydrogen_detail::check(hg_dispatch_grid(data.ctx), "dispatch_grid");
}
// src/erosion.cpp:132-156
inline void dispatch_particle(Programs&, State::World::Textures& data, bool should_rain) {
    hydrogen_detail::check(hg_dispatch_particle(data.ctx, data.time, should_rain ? 1 : 0), "dispatch_particle");
}

}  // namespace Erosion
