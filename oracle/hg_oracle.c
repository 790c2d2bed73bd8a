/* hg_oracle.c — CPU restatement of hydro-gen's erosion step.  TEST INFRASTRUCTURE.
 *
 * This file is the parity oracle for the CUDA path: a scalar fp32 C restatement
 * of the reference's GLSL compute shaders, one function per shader main(), on
 * the reference's own data layout (interleaved RGBA32F "textures", ping-pong
 * pairs, index [y][x]).  It is never linked into, imported by or called from
 * the product library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * PINNED AGAINST THE REFERENCE'S OWN SHADERS: the reference ships no tests or
 * fixtures and there is no GL here, but its ten compute shaders
 * (six step shaders, rain, heightmap, particle, particle_erosion) compile for
 * the CPU through a C++ shim of the GLSL vocabulary (oracle/refshader/, output
 * oracle/_ref/).  This restatement reproduces them bit for bit after every
 * dispatch, over multi-step grid and droplet runs and for the generated terrain
 * (tests/test_refshaders.py), and so do the committed fixtures (tests/golden/).
 * Every function cites the file:line it follows.
 *
 * Arithmetic rules: every GLSL expression is written with the same operand
 * order and association; build with -ffp-contract=off so no FMA is formed.
 * min/max/clamp/mix/fract/mod/smoothstep/atan/exp/sin come from
 * include/hg_defined_math.h (the one shared piece: it defines the built-ins GL
 * leaves implementation-defined).  Out-of-bounds imageLoad/texelFetch return 0
 * (SURVEY.md §8a hazard 3); uninitialised textures/SSBO are zero (hazard 6).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hg_types.h"
#include "../include/hg_defined_math.h"
#include "hg_oracle.h"

typedef struct { float x, y; } vec2;
typedef struct { float x, y, z; } vec3;
typedef struct { float x, y, z, w; } vec4;

/* ------------------------------------------------------------------ helpers */

static inline vec4 ld4(const float* t, int W, int x, int y) {
    const float* p = t + ((size_t)y * W + x) * 4;
    vec4 v = {p[0], p[1], p[2], p[3]};
    return v;
}
static inline void st4(float* t, int W, int x, int y, vec4 v) {
    float* p = t + ((size_t)y * W + x) * 4;
    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
}
static inline int oob(int W, int H, int x, int y) {
    return x < 0 || x > W - 1 || y < 0 || y > H - 1;
}
/* imageLoad / texelFetch with the out-of-bounds result defined as 0 */
static inline vec4 ld4_zero(const float* t, int W, int H, int x, int y) {
    vec4 z = {0.0f, 0.0f, 0.0f, 0.0f};
    return oob(W, H, x, y) ? z : ld4(t, W, x, y);
}
static inline float comp(vec4 v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
static inline float length2(float x, float y) { return sqrtf(x * x + y * y); }

/* ------------------------------------------------------------- ping-pong --- */
/* gl::Tex_pair (src/shaderprogram.hpp:58-70, src/shaderprogram.cpp:51-82) */
static void pair_swap(orc_pair* p) {
    p->cntr++;
    p->idx_read = p->cntr % 2;
    p->idx_write = (p->cntr + 1) % 2;
}
static inline float* rd(orc_pair* p) { return p->tex[p->idx_read]; }
static inline float* wr(orc_pair* p) { return p->tex[p->idx_write]; }

/* ----------------------------------------------------------------- noise --- */
/* glsl/simplex_noise.glsl:247-256 */
typedef struct {
    float seed, persistance, lacunarity, scale, redistribution;
    int octaves, terbulance, ridge;
} fbm_opts;

/* gln_rand3, simplex_noise.glsl:320 */
static inline float rand3(float p) { return hg_mod(((p * 34.0f) + 1.0f) * p, 289.0f); }

/* gln_simplex, simplex_noise.glsl:377-402 */
static float gln_simplex(vec2 v) {
    const float Cx = 0.211324865405187f, Cy = 0.366025403784439f,
                Cz = -0.577350269189626f, Cw = 0.024390243902439f;
    float dvc = v.x * Cy + v.y * Cy;
    vec2 i = {floorf(v.x + dvc), floorf(v.y + dvc)};
    float dic = i.x * Cx + i.y * Cx;
    vec2 x0 = {v.x - i.x + dic, v.y - i.y + dic};
    vec2 i1;
    if (x0.x > x0.y) { i1.x = 1.0f; i1.y = 0.0f; } else { i1.x = 0.0f; i1.y = 1.0f; }
    vec4 x12 = {x0.x + Cx, x0.y + Cx, x0.x + Cz, x0.y + Cz};
    x12.x -= i1.x;
    x12.y -= i1.y;
    i.x = hg_mod(i.x, 289.0f);
    i.y = hg_mod(i.y, 289.0f);
    vec3 p;
    p.x = rand3(rand3(i.y + 0.0f) + i.x + 0.0f);
    p.y = rand3(rand3(i.y + i1.y) + i.x + i1.x);
    p.z = rand3(rand3(i.y + 1.0f) + i.x + 1.0f);
    vec3 m;
    m.x = hg_max(0.5f - (x0.x * x0.x + x0.y * x0.y), 0.0f);
    m.y = hg_max(0.5f - (x12.x * x12.x + x12.y * x12.y), 0.0f);
    m.z = hg_max(0.5f - (x12.z * x12.z + x12.w * x12.w), 0.0f);
    m.x = m.x * m.x; m.y = m.y * m.y; m.z = m.z * m.z;
    m.x = m.x * m.x; m.y = m.y * m.y; m.z = m.z * m.z;
    vec3 x = {2.0f * hg_fract(p.x * Cw) - 1.0f, 2.0f * hg_fract(p.y * Cw) - 1.0f,
              2.0f * hg_fract(p.z * Cw) - 1.0f};
    vec3 h = {fabsf(x.x) - 0.5f, fabsf(x.y) - 0.5f, fabsf(x.z) - 0.5f};
    vec3 ox = {floorf(x.x + 0.5f), floorf(x.y + 0.5f), floorf(x.z + 0.5f)};
    vec3 a0 = {x.x - ox.x, x.y - ox.y, x.z - ox.z};
    m.x *= 1.79284291400159f - 0.85373472095314f * (a0.x * a0.x + h.x * h.x);
    m.y *= 1.79284291400159f - 0.85373472095314f * (a0.y * a0.y + h.y * h.y);
    m.z *= 1.79284291400159f - 0.85373472095314f * (a0.z * a0.z + h.z * h.z);
    vec3 g;
    g.x = a0.x * x0.x + h.x * x0.y;
    g.y = a0.y * x12.x + h.y * x12.y;
    g.z = a0.z * x12.z + h.z * x12.w;
    return 130.0f * (m.x * g.x + m.y * g.y + m.z * g.z);
}

/* pow(result, redistribution): every call site passes the literal 1.0
 * (heightmap.glsl:60,92; rain.glsl:41), for which pow is the identity on
 * result >= 0; result < 0 is undefined in GLSL and is defined here as the
 * identity too (rain then clamps it with max(0, .), SURVEY.md §8a hazard 2). */
static inline float pow_redistribution(float x, float y) { (void)y; return x; }

/* gln_sfbm, simplex_noise.glsl:419-456 */
static float gln_sfbm(vec2 v, fbm_opts o) {
    v.x += (o.seed * 100.0f);
    v.y += (o.seed * 100.0f);
    int ridge = o.terbulance && o.ridge;
    float result = 0.0f, amplitude = 1.0f, frequency = 1.0f, maximum = amplitude;
    for (int i = 0; i < 30; i++) {
        if (i >= o.octaves) break;
        vec2 p = {v.x * frequency * o.scale, v.y * frequency * o.scale};
        float n = gln_simplex(p);
        if (o.terbulance) n = fabsf(n);
        if (ridge) n = 1.0f - n;
        result += n * amplitude;
        frequency *= o.lacunarity;
        amplitude *= o.persistance;
        maximum += amplitude;
    }
    return pow_redistribution(result, o.redistribution) / maximum;
}

/* hash(ivec2), simplex_noise.glsl:580-588 — GLSL int arithmetic wraps mod 2^32 */
static inline vec2 ihash(int32_t px, int32_t py) {
    uint32_t ux = (uint32_t)px, uy = (uint32_t)py;
    uint32_t nx = ux * 3u + uy * 311u;
    uint32_t ny = ux * 37u + uy * 113u;
    nx = (nx << 13) ^ nx;
    ny = (ny << 13) ^ ny;
    nx = nx * (nx * nx * 15731u + 789221u) + 1376312589u;
    ny = ny * (ny * ny * 15731u + 789221u) + 1376312589u;
    const float d = (float)0x0fffffff;
    vec2 r = {-1.0f + 2.0f * (float)(int32_t)(nx & 0x0fffffffu) / d,
              -1.0f + 2.0f * (float)(int32_t)(ny & 0x0fffffffu) / d};
    return r;
}

/* noised, simplex_noise.glsl:591-612: value + analytic derivatives */
static vec3 noised(vec2 p) {
    float flx = floorf(p.x), fly = floorf(p.y);
    int32_t ix = (int32_t)flx, iy = (int32_t)fly;
    vec2 f = {p.x - flx, p.y - fly};
    vec2 u = {f.x * f.x * f.x * (f.x * (f.x * 6.0f - 15.0f) + 10.0f),
              f.y * f.y * f.y * (f.y * (f.y * 6.0f - 15.0f) + 10.0f)};
    vec2 du = {30.0f * f.x * f.x * (f.x * (f.x - 2.0f) + 1.0f),
               30.0f * f.y * f.y * (f.y * (f.y - 2.0f) + 1.0f)};
    vec2 ga = ihash(ix, iy);
    vec2 gb = ihash((int32_t)((uint32_t)ix + 1u), iy);
    vec2 gc = ihash(ix, (int32_t)((uint32_t)iy + 1u));
    vec2 gd = ihash((int32_t)((uint32_t)ix + 1u), (int32_t)((uint32_t)iy + 1u));
    float va = ga.x * (f.x - 0.0f) + ga.y * (f.y - 0.0f);
    float vb = gb.x * (f.x - 1.0f) + gb.y * (f.y - 0.0f);
    float vc = gc.x * (f.x - 0.0f) + gc.y * (f.y - 1.0f);
    float vd = gd.x * (f.x - 1.0f) + gd.y * (f.y - 1.0f);
    float k = va - vb - vc + vd;
    vec3 r;
    r.x = va + u.x * (vb - va) + u.y * (vc - va) + u.x * u.y * k;
    r.y = ga.x + u.x * (gb.x - ga.x) + u.y * (gc.x - ga.x) + u.x * u.y * (ga.x - gb.x - gc.x + gd.x)
          + du.x * (u.y * k + vb - va);
    r.z = ga.y + u.x * (gb.y - ga.y) + u.y * (gc.y - ga.y) + u.x * u.y * (ga.y - gb.y - gc.y + gd.y)
          + du.y * (u.x * k + vc - va);
    return r;
}

/* perlfbm, simplex_noise.glsl:614-652 */
static float perlfbm(vec2 v, fbm_opts o) {
    v.x += (o.seed * 100.0f);
    v.y += (o.seed * 100.0f);
    int ridge = o.terbulance && o.ridge;
    float result = 0.0f, amplitude = 1.0f, frequency = 1.0f, maximum = amplitude;
    for (int i = 0; i < 30; i++) {
        if (i >= o.octaves) break;
        vec2 p = {v.x * frequency * o.scale, v.y * frequency * o.scale};
        vec3 res = noised(p);
        float n = (res.x + 1.0f) / 2.0f;
        if (o.terbulance) n = fabsf(n);
        if (ridge) n = 1.0f - n;
        result += n * amplitude;
        frequency *= o.lacunarity;
        amplitude *= o.persistance;
        maximum += amplitude;
    }
    return pow_redistribution(result, o.redistribution) / maximum;
}

/* erosion_perlfbm, simplex_noise.glsl:654-693 */
static float erosion_perlfbm(vec2 v, fbm_opts o) {
    v.x += (o.seed * 100.0f);
    v.y += (o.seed * 100.0f);
    int ridge = o.terbulance && o.ridge;
    float result = 0.0f, amplitude = 1.0f, frequency = 1.5f, maximum = amplitude;
    vec2 dsum = {0.0f, 0.0f};
    for (int i = 0; i < 30; i++) {
        if (i >= o.octaves) break;
        vec2 p = {v.x * frequency * o.scale, v.y * frequency * o.scale};
        vec3 res = noised(p);
        if (o.terbulance) { res.x = fabsf(res.x); res.y = fabsf(res.y); res.z = fabsf(res.z); }
        if (ridge) { res.x = 1.0f - res.x; res.y = 1.0f - res.y; res.z = 1.0f - res.z; }
        dsum.x += res.y;
        dsum.y += res.z;
        float n = (res.x + 1.0f) / 2.0f;
        result += n * amplitude / (1.0f + (dsum.x * dsum.x + dsum.y * dsum.y));
        frequency *= o.lacunarity;
        amplitude *= o.persistance;
        maximum += amplitude;
    }
    return pow_redistribution(result, o.redistribution) / maximum;
}

/* ------------------------------------------------------ heightmap.glsl ----- */
/* heightmap.glsl:29-45 */
static float round_mask(float val, vec2 uv) {
    /* pow(x, 2.0) with x possibly negative is undefined in GLSL; defined as x*x */
    float a = uv.x - 0.5f, b = uv.y - 0.5f;
    float v = 1.0f - (a * a + b * b + 0.75f);
    return val * hg_max(0.0f, v * 1.0f);
}
static float slope_mask(float v, vec2 uv) { return 0.25f * v + 0.75f * (v * uv.x * uv.y); }
static float power_mask(float val) {
    float b = val + 0.5f;
    float p3 = b * b * b; /* pow(b, 3) */
    return val * (((p3 - 0.125f) / 3.25f) * 0.55f + 0.45f);
}
static float exp_mask(float val) { return val * (hg_expf(val) - 1.0f) / 1.718f; }

/* heightmap.glsl:52-158, dispatched by State::World::gen_heightmap
 * (src/state.cpp:116-147): writes the WRITE textures, then swaps H,V,F,S. */
void orc_gen_heightmap(orc_world* w) {
    const hg_map_settings_data cfg = w->map;
    const int W = w->W, H = w->H;
    float* dst_h = wr(&w->heightmap);
    float* dst_v = wr(&w->velocity);
    float* dst_f = wr(&w->flux);
    float* dst_s = wr(&w->sediment);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            vec2 uv = {(float)x / (float)W, (float)y / (float)H};
            fbm_opts opts = {cfg.seed, cfg.persistance, cfg.lacunarity, cfg.scale, 1.0f, cfg.octaves, 0, 0};
            vec2 dist = {1.0f, 1.0f};
            if (cfg.domain_warp != 0) {
                vec2 a = {(float)x + 2.3f, (float)y + 2.9f};
                vec2 b = {(float)x - 3.1f, (float)y - 4.3f};
                dist.x = perlfbm(a, opts);
                dist.y = perlfbm(b, opts);
                if (cfg.domain_warp == 2) {
                    vec2 c = {(float)x + cfg.domain_warp_scale * dist.x - 5.7f,
                              (float)y + cfg.domain_warp_scale * dist.y + 27.9f};
                    vec2 d = {(float)x + cfg.domain_warp_scale * dist.x + 11.5f,
                              (float)y + cfg.domain_warp_scale * dist.y - 23.7f};
                    vec2 nd = {perlfbm(c, opts), perlfbm(d, opts)};
                    dist = nd;
                }
            }
            vec2 wp = {(float)x + cfg.domain_warp_scale * dist.x, (float)y + cfg.domain_warp_scale * dist.y};
            float val = erosion_perlfbm(wp, opts);

            float height_multiplier = cfg.height_mult;
            if (cfg.uplift != 0) {
                fbm_opts up = {cfg.seed, cfg.persistance, cfg.lacunarity, cfg.scale / cfg.uplift_scale,
                               1.0f, cfg.octaves, 1, 1};
                vec2 upos = {(float)x - 7.3f, (float)y + 19.9f};
                float upv = gln_sfbm(upos, up);
                if (cfg.mask_exp != 0) height_multiplier += 2.5f;
                height_multiplier += 1.0f;
                val *= upv;
            }
            if (cfg.mask_round != 0) {
                if (cfg.mask_exp != 0) height_multiplier += 16.0f;
                val = round_mask(val, uv);
            }
            if (cfg.mask_exp != 0) {
                height_multiplier += 2.0f;
                val = exp_mask(val);
            }
            if (cfg.mask_power != 0) {
                height_multiplier += 1.25f;
                if (cfg.mask_exp != 0) height_multiplier += 2.0f;
                val = power_mask(val);
            }
            if (cfg.mask_slope != 0) val = slope_mask(val, uv);

            val = val + (float)(4u * cfg.mask_round) * val;
            if (cfg.terrace > 0) {
                int levels = cfg.terrace;
                float lol = floorf(val / (1.0f / ((float)levels * height_multiplier)));
                float t_scl = cfg.terrace_scale;
                val = (t_scl * lol * (1.0f / ((float)levels * height_multiplier))) + val * (1.0f - t_scl);
            }
            float rock_val = hg_min(cfg.max_height, val * cfg.max_height * height_multiplier);
            vec2 dp = {(float)x + 13.7f, (float)y + 27.1f};
            float dirt_val = perlfbm(dp, opts) + 1.5f;
            dirt_val *= cfg.max_dirt;
            vec4 terrain = {rock_val, dirt_val, 0.0f, 0.0f};
            terrain.w = terrain.x + terrain.y + terrain.z;
            vec4 zero = {0.0f, 0.0f, 0.0f, 0.0f};
            st4(dst_h, W, x, y, terrain);
            st4(dst_v, W, x, y, zero);
            st4(dst_f, W, x, y, zero);
            st4(dst_s, W, x, y, zero);
        }
    }
    pair_swap(&w->heightmap);
    pair_swap(&w->velocity);
    pair_swap(&w->flux);
    pair_swap(&w->sediment);
}

/* ------------------------------------------------------------ rain.glsl ---- */
/* rain.glsl:32-56, dispatched by Erosion::dispatch_grid_rain (src/erosion.cpp:76-89) */
void orc_dispatch_grid_rain(orc_world* w) {
    const int W = w->W, H = w->H;
    const hg_rain_data set = w->rain;
    const hg_map_settings_data map_set = w->map;
    const float* src = rd(&w->heightmap);
    float* dst = wr(&w->heightmap);
    const float time = w->time;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            vec4 terr = ld4(src, W, x, y);
            fbm_opts opts = {hg_fract(time * 1.372914227e3f) * 1000.f, 0.5f, 2.0f, set.drops, 1.0f, 8, 0, 0};
            vec2 p = {(float)x, (float)y};
            float r = hg_max(0.0f, gln_sfbm(p, opts));
            float incr = set.amount * r;
            float mountain = terr.w - map_set.max_height * set.mountain_thresh;
            if (mountain > 0.0f) {
                incr += mountain * set.mountain_multip * r / ((1.0f - set.mountain_thresh) * map_set.max_height);
            }
            terr.z += incr;
            terr.w = terr.x + terr.y + terr.z;
            st4(dst, W, x, y, terr);
        }
    }
    pair_swap(&w->heightmap);
}

/* ------------------------------------------------------ hydro_flux.glsl ---- */
/* hydro_flux.glsl:33-47 */
static inline float get_wheight(const float* h, int W, int H, int x, int y) {
    if (oob(W, H, x, y)) return HG_OOB_HEIGHT;
    return ld4(h, W, x, y).w;
}
/* hydro_flux.glsl:77-166 */
static void flux_pass(orc_world* w) {
    const int W = w->W, H = w->H;
    const hg_erosion_data set = w->erosion;
    const float* hm = rd(&w->heightmap);
    const float* fm = rd(&w->flux);
    const float* vm = rd(&w->velocity);
    float* oh = wr(&w->heightmap);
    float* of = wr(&w->flux);
    float* ov = wr(&w->velocity);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            vec4 out_flux = ld4_zero(fm, W, H, x, y);
            vec4 vel = ld4(vm, W, x, y);
            vec4 terrain = ld4(hm, W, x, y);
            float d1 = terrain.z;
            vec4 d_height;
            d_height.x = terrain.w - get_wheight(hm, W, H, x - 1, y);
            d_height.y = terrain.w - get_wheight(hm, W, H, x + 1, y);
            d_height.z = terrain.w - get_wheight(hm, W, H, x, y + 1);
            d_height.w = terrain.w - get_wheight(hm, W, H, x, y - 1);
            vec4 fl = ld4_zero(fm, W, H, x - 1, y);
            vec4 fr = ld4_zero(fm, W, H, x + 1, y);
            vec4 ft = ld4_zero(fm, W, H, x, y + 1);
            vec4 fb = ld4_zero(fm, W, H, x, y - 1);
            vec4 own = out_flux; /* get_flux(pos) before the update, used by the velocity below */
            vec4 in_flux = {fl.y, fr.x, ft.w, fb.z};
            /* A = 1, L = 1 (hydro_flux.glsl:31, bindings.glsl:18): d_t*A*(G*dh)/L */
            out_flux.x = hg_max(0.0f, set.ENERGY_KEPT * out_flux.x + set.d_t * 1.0f * (set.G * d_height.x) / 1.0f);
            out_flux.y = hg_max(0.0f, set.ENERGY_KEPT * out_flux.y + set.d_t * 1.0f * (set.G * d_height.y) / 1.0f);
            out_flux.z = hg_max(0.0f, set.ENERGY_KEPT * out_flux.z + set.d_t * 1.0f * (set.G * d_height.z) / 1.0f);
            out_flux.w = hg_max(0.0f, set.ENERGY_KEPT * out_flux.w + set.d_t * 1.0f * (set.G * d_height.w) / 1.0f);
            if (x <= 0) out_flux.x = 0.0f;
            else if (x >= W - 1) out_flux.y = 0.0f;
            if (y <= 0) out_flux.w = 0.0f;
            else if (y >= H - 1) out_flux.z = 0.0f;
            float sum_in_flux = in_flux.x + in_flux.y + in_flux.z + in_flux.w;
            float sum_out_flux = out_flux.x + out_flux.y + out_flux.z + out_flux.w;
            float K = hg_min(1.0f, (terrain.z * 1.0f * 1.0f) / (sum_out_flux * set.d_t));
            out_flux.x *= K; out_flux.y *= K; out_flux.z *= K; out_flux.w *= K;
            sum_out_flux *= K;
            float d_volume = set.d_t * (sum_in_flux - sum_out_flux);
            float d2 = hg_max(0.0f, d1 + (d_volume / (1.0f * 1.0f)));
            terrain.z = d2;
            terrain.w = terrain.x + d2 + terrain.y;
            vel.z = (d1 + d2);
            if (vel.z > 0.0f) {
                vel.x = (fl.y - own.x + own.y - fr.x) / (1.0f * vel.z);
                vel.y = (fb.z - own.w + own.z - ft.w) / (1.0f * vel.z);
            } else {
                vel.x = 0.0f;
                vel.y = 0.0f;
            }
            st4(of, W, x, y, out_flux);
            st4(ov, W, x, y, vel);
            st4(oh, W, x, y, terrain);
        }
    }
}

/* --------------------------------------------------- hydro_erosion.glsl ---- */
/* hydro_erosion.glsl:23-35 */
static vec3 terr_normal_img(const float* hm, int W, int H, int x, int y) {
    vec4 r = ld4_zero(hm, W, H, x + 1, y);
    vec4 l = ld4_zero(hm, W, H, x - 1, y);
    vec4 b = ld4_zero(hm, W, H, x, y - 1);
    vec4 t = ld4_zero(hm, W, H, x, y + 1);
    float dx = (r.x + r.y - l.x - l.y);
    float dz = (t.x + t.y - b.x - b.y);
    /* normalize(cross(vec3(2L, dx, 0), vec3(0, dz, 2L))) */
    vec3 a = {2.0f * 1.0f, dx, 0.0f}, c = {0.0f, dz, 2.0f * 1.0f};
    vec3 n = {a.y * c.z - a.z * c.y, a.z * c.x - a.x * c.z, a.x * c.y - a.y * c.x};
    float inv = 1.0f / sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    n.x *= inv; n.y *= inv; n.z *= inv;
    return n;
}
/* hydro_erosion.glsl:37-92 */
static void erosion_pass(orc_world* w) {
    const int W = w->W, H = w->H;
    const hg_erosion_data set = w->erosion;
    const float* hm = rd(&w->heightmap);
    const float* sm = rd(&w->sediment);
    const float* vm = rd(&w->velocity);
    float* oh = wr(&w->heightmap);
    float* os = wr(&w->sediment);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            vec4 vel = ld4(vm, W, x, y);
            vec4 terrain4 = ld4(hm, W, x, y);
            vec4 sediment4 = ld4(sm, W, x, y);
            float terrain[2] = {terrain4.x, terrain4.y};
            float sediment[2] = {sediment4.x, sediment4.y};
            float dd = vel.z;
            float ero_vel = length2(vel.x, vel.y);
            if (dd < 1e-3f) {
                dd = hg_max(5e-4f, dd);
                ero_vel = hg_mix(length2(vel.x, vel.y), 0.0f, hg_smoothstep(1e-3f, 5e-4f, dd));
            } else {
                ero_vel = length2(vel.x, vel.y);
            }
            float cap = 0.0f;
            vec3 norm = terr_normal_img(hm, W, H, x, y);
            float sin_a = fabsf(fabsf(sqrtf(1.0f - norm.y * norm.y)));
            for (int i = HG_SED_LAYERS - 1; i >= 0; i--) {
                float Kls = set.d_t * set.Ks[i];
                float Kld = set.d_t * set.Kd[i];
                float c = hg_max(0.0f, set.Kc * hg_max(0.02f, sin_a) * ero_vel - cap);
                if (c > sediment[i]) {
                    float old_terr = terrain[i];
                    float delta = Kls * (c - sediment[i]);
                    terrain[i] -= delta;
                    sediment[i] += delta;
                    if (terrain[i] < 0.0f) {
                        sediment[i] += terrain[i];
                        terrain[i] = 0.0f;
                        cap += old_terr;
                    } else {
                        break;
                    }
                } else {
                    float delta = Kld * (sediment[i] - c);
                    terrain[i] += delta;
                    sediment[i] -= delta;
                }
            }
            for (int i = 0; i < HG_SED_LAYERS - 1; i++) {
                float conv = sediment[i] * set.Kconv * set.d_t;
                sediment[i + 1] += conv;
                sediment[i] -= conv;
            }
            terrain4.x = terrain[0]; terrain4.y = terrain[1];
            sediment4.x = sediment[0]; sediment4.y = sediment[1];
            terrain4.w = terrain4.x + terrain4.y + terrain4.z;
            st4(os, W, x, y, sediment4);
            st4(oh, W, x, y, terrain4);
        }
    }
}

/* ---------------------------------------------- sediment_transport.glsl ---- */
/* img_interpolation.glsl:3-22 (WORLD_SCALE = 1): 4 texelFetch + 3 mix */
static vec4 img_bilinear(const float* img, int W, int H, vec2 sp) {
    /* a NaN coordinate has no defined ivec2(); defined here as 0 */
    if (!(sp.x == sp.x)) sp.x = 0.0f;
    if (!(sp.y == sp.y)) sp.y = 0.0f;
    int px = (int)(sp.x * 1.0f), py = (int)(sp.y * 1.0f);
    float sx = hg_fract(sp.x * 1.0f), sy = hg_fract(sp.y * 1.0f);
    vec4 a = ld4_zero(img, W, H, px, py), b = ld4_zero(img, W, H, px + 1, py);
    vec4 c = ld4_zero(img, W, H, px, py + 1), d = ld4_zero(img, W, H, px + 1, py + 1);
    vec4 v1 = {hg_mix(a.x, b.x, sx), hg_mix(a.y, b.y, sx), hg_mix(a.z, b.z, sx), hg_mix(a.w, b.w, sx)};
    vec4 v2 = {hg_mix(c.x, d.x, sx), hg_mix(c.y, d.y, sx), hg_mix(c.z, d.z, sx), hg_mix(c.w, d.w, sx)};
    vec4 v = {hg_mix(v1.x, v2.x, sy), hg_mix(v1.y, v2.y, sy), hg_mix(v1.z, v2.z, sy), hg_mix(v1.w, v2.w, sy)};
    return v;
}
/* sediment_transport.glsl:26-30, 66-93 */
static void sediment_pass(orc_world* w) {
    const int W = w->W, H = w->H;
    const hg_erosion_data set = w->erosion;
    const float* hm = rd(&w->heightmap);
    const float* vm = rd(&w->velocity);
    const float* sm = rd(&w->sediment);
    float* oh = wr(&w->heightmap);
    float* os = wr(&w->sediment);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            vec4 vel = ld4(vm, W, x, y);
            vec2 back = {(float)x - vel.x * set.d_t, (float)y - vel.y * set.d_t};
            back.x = hg_clamp(back.x, 0.0f, (float)(W - 1));
            back.y = hg_clamp(back.y, 0.0f, (float)(H - 1));
            vec4 st = img_bilinear(sm, W, H, back);
            vec4 terrain = ld4(hm, W, x, y);
            terrain.z *= (1.0f - set.Ke * set.d_t);
            terrain.w = terrain.x + terrain.y + terrain.z;
            st4(os, W, x, y, st);
            st4(oh, W, x, y, terrain);
        }
    }
}

/* --------------------------------------------------- thermal_erosion.glsl -- */
/* thermal_erosion.glsl:20-26 */
static inline vec4 th_get_height(const float* hm, int W, int H, int x, int y) {
    vec4 far = {HG_OOB_HEIGHT, HG_OOB_HEIGHT, HG_OOB_HEIGHT, HG_OOB_HEIGHT};
    return oob(W, H, x, y) ? far : ld4(hm, W, x, y);
}
/* thermal_erosion.glsl:28-115 with t_layer = layer (src/erosion.cpp:53-60) */
static void thermal_flux_pass(orc_world* w, int layer) {
    const int W = w->W, H = w->H;
    const hg_erosion_data set = w->erosion;
    const float* hm = rd(&w->heightmap);
    float* oc = wr(&w->thermal_c);
    float* od = wr(&w->thermal_d);
    static const int off[2][4][2] = {
        {{-1, 0}, {1, 0}, {0, 1}, {0, -1}},      /* L R T B */
        {{-1, 1}, {1, 1}, {-1, -1}, {1, -1}}};   /* LT RT LB RB */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            vec4 terrain = ld4(hm, W, x, y);
            float d_h[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
            for (int i = 0; i <= layer; i++) {
                for (int j = 0; j < 2; j++)
                    for (int k = 0; k < 4; k++)
                        d_h[j][k] += comp(terrain, i) - comp(th_get_height(hm, W, H, x + off[j][k][0], y + off[j][k][1]), i);
            }
            float Hm = 0.0f;
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 4; k++)
                    if (d_h[j][k] > Hm) Hm = d_h[j][k];
            Hm = hg_min(comp(terrain, layer), Hm);
            float out[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
            float bk = 0.0f;
            float sharpness = 1.0f;
            for (int j = 0; j < 2; j++) {
                for (int k = 0; k < 4; k++) {
                    float b = d_h[j][k];
                    if (b <= 0.0f) { out[j][k] = 0.0f; continue; }
                    float d = 1.0f;
                    if (j == 1) d *= 1.41421356237309504880f; /* sqrt(2.0) */
                    float alph = hg_atanf(b / (d / 1.0f));
                    float Kl_alph = set.Kalpha[layer];
                    if (alph > Kl_alph) {
                        float newsh = 1.0f + alph - Kl_alph;
                        if (newsh > sharpness) sharpness = newsh;
                        bk += b;
                        out[j][k] = 1.0f;
                        continue;
                    }
                    out[j][k] = 0.0f;
                }
            }
            sharpness *= sharpness * sharpness;
            float Klspeed = set.Kspeed[layer];
            float S = set.d_t * Klspeed * sharpness * 1.0f * Hm / 2.0f;
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 4; k++)
                    out[j][k] = (out[j][k] == 1.0f) ? S * d_h[j][k] / bk : 0.0f;
            vec4 c = {out[0][0], out[0][1], out[0][2], out[0][3]};
            vec4 d = {out[1][0], out[1][1], out[1][2], out[1][3]};
            st4(oc, W, x, y, c);
            st4(od, W, x, y, d);
        }
    }
}

/* ------------------------------------------------ thermal_transport.glsl --- */
/* thermal_transport.glsl:31-65 */
static void thermal_transport_pass(orc_world* w, int layer) {
    const int W = w->W, H = w->H;
    const float* hm = rd(&w->heightmap);
    const float* tc = rd(&w->thermal_c);
    const float* td = rd(&w->thermal_d);
    float* oh = wr(&w->heightmap);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            vec4 terrain = ld4(hm, W, x, y);
            float in_flux = 0.0f;
            in_flux += ld4_zero(tc, W, H, x - 1, y).y;
            in_flux += ld4_zero(tc, W, H, x + 1, y).x;
            in_flux += ld4_zero(tc, W, H, x, y + 1).w;
            in_flux += ld4_zero(tc, W, H, x, y - 1).z;
            in_flux += ld4_zero(td, W, H, x - 1, y + 1).w;
            in_flux += ld4_zero(td, W, H, x + 1, y + 1).z;
            in_flux += ld4_zero(td, W, H, x - 1, y - 1).y;
            in_flux += ld4_zero(td, W, H, x + 1, y - 1).x;
            float sum_flux = 0.0f;
            vec4 oc = ld4(tc, W, x, y), od = ld4(td, W, x, y);
            sum_flux -= oc.x; sum_flux -= oc.y; sum_flux -= oc.z; sum_flux -= oc.w;
            sum_flux -= od.x; sum_flux -= od.y; sum_flux -= od.z; sum_flux -= od.w;
            sum_flux += in_flux;
            if (layer == 0) terrain.x += sum_flux; else terrain.y += sum_flux;
            terrain.w = terrain.x + terrain.y + terrain.z;
            st4(oh, W, x, y, terrain);
        }
    }
}

/* src/erosion.cpp:103-121 */
static void run_thermal_erosion(orc_world* w) {
    for (int i = 0; i < HG_SED_LAYERS; i++) {
        thermal_flux_pass(w, i);
        pair_swap(&w->thermal_c);
        pair_swap(&w->thermal_d);
        thermal_transport_pass(w, i);
        pair_swap(&w->heightmap);
    }
}

/* ------------------------------------------------------- smoothing.glsl ---- */
/* smoothing.glsl:22-103; `momentum` selects whether momentmap/out_momentmap are
 * bound (particle mode, src/erosion.cpp:148-155) or not (grid, :193-199). */
static void smooth_pass(orc_world* w, int momentum_bound) {
    const int W = w->W, H = w->H;
    const hg_erosion_data set = w->erosion;
    const float* hm = rd(&w->heightmap);
    float* oh = wr(&w->heightmap);
    const float* mm = rd(&w->velocity);
    float* om = wr(&w->velocity);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            float d_time = set.d_t;
            vec4 terrain = ld4(hm, W, x, y);
            vec2 terr = {terrain.x, terrain.y};
            if (x == 0 || y == 0 || x == W - 1 || y == H - 1) {
                st4(oh, W, x, y, terrain);
                /* out_momentmap is not written on the border in the reference
                 * (stale texel, never read by a droplet: SURVEY.md App. A); 0 here */
                if (momentum_bound && set.particle_count != 0) {
                    vec4 z = {0.0f, 0.0f, 0.0f, 0.0f};
                    st4(om, W, x, y, z);
                }
                continue;
            }
            vec4 l4 = ld4(hm, W, x - 1, y), r4 = ld4(hm, W, x + 1, y);
            vec4 t4 = ld4(hm, W, x, y + 1), b4 = ld4(hm, W, x, y - 1);
            vec2 l = {l4.x, l4.y}, r = {r4.x, r4.y}, t = {t4.x, t4.y}, b = {b4.x, b4.y};
            vec2 d_l = {terr.x - l.x, terr.y - l.y}; d_l.y += d_l.x;
            vec2 d_r = {terr.x - r.x, terr.y - r.y}; d_r.y += d_r.x;
            vec2 d_t = {terr.x - t.x, terr.y - t.y}; d_t.y += d_t.x;
            vec2 d_b = {terr.x - b.x, terr.y - b.y}; d_b.y += d_b.x;
            float g_hdiff = (d_l.y + d_r.y + d_t.y + d_b.y) / (4.0f);
            float r_hdiff = (d_l.x + d_r.x + d_t.x + d_b.x) / (4.0f);
            g_hdiff = fabsf(g_hdiff);
            r_hdiff = fabsf(r_hdiff);
            vec2 x_crv = {d_l.x * d_r.x, d_l.y * d_r.y};
            vec2 y_crv = {d_t.x * d_b.x, d_t.y * d_b.y};
            if ((((-d_l.x) > r_hdiff || (-d_r.x) > r_hdiff) && x_crv.x > 0.0f)
                || (((-d_t.x) > r_hdiff || (-d_b.x) > r_hdiff) && y_crv.x > 0.0f)) {
                terr.x = (terr.x + l.x + r.x + t.x + b.x) / 5.0f;
            }
            if ((((-d_l.y) > g_hdiff || (-d_r.y) > g_hdiff) && x_crv.y > 0.0f)
                || (((-d_t.y) > g_hdiff || (-d_b.y) > g_hdiff) && y_crv.y > 0.0f)) {
                terr.y = (terr.y + l.y + r.y + t.y + b.y) / 5.0f;
            }
            float multip = hg_clamp(set.Kspeed[1] * d_time, 0.0f, 1.0f);
            if (set.particle_count != 0) {
                float pc = (float)set.particle_count;
                vec4 momentum = momentum_bound ? ld4(mm, W, x, y) : (vec4){0.0f, 0.0f, 0.0f, 0.0f};
                float keep = hg_clamp(1.0f - (1e-12f * pc), 0.0f, 1.0f);
                momentum.x *= keep;
                momentum.y *= keep;
                momentum.x += (1e-12f * pc) * momentum.z;
                momentum.y += (1e-12f * pc) * momentum.w;
                momentum.z = 0.0f;
                momentum.w = 0.0f;
                terrain.z *= hg_clamp(1.0f - (8e-8f * pc), 0.0f, 1.0f);
                if (terrain.z < 1e-6f) terrain.z = 0.0f; /* the border tests are dead here (:83-86) */
                if (length2(momentum.x, momentum.y) < 1e-12f) { momentum.x = 0.0f; momentum.y = 0.0f; }
                if (momentum_bound) st4(om, W, x, y, momentum);
                multip = hg_clamp(set.Kspeed[1] * d_time, 0.0f, 1.0f);
            }
            terrain.x = multip * terr.x + (1.0f - multip) * terrain.x;
            terrain.y = multip * terr.y + (1.0f - multip) * terrain.y;
            terrain.w = terrain.x + terrain.y + terrain.z;
            st4(oh, W, x, y, terrain);
        }
    }
}

/* Erosion::dispatch_grid, src/erosion.cpp:158-200 */
void orc_dispatch_grid(orc_world* w) {
    flux_pass(w);
    pair_swap(&w->heightmap); pair_swap(&w->flux); pair_swap(&w->velocity);
    erosion_pass(w);
    pair_swap(&w->heightmap); pair_swap(&w->sediment);
    sediment_pass(w);
    pair_swap(&w->heightmap); pair_swap(&w->sediment);
    run_thermal_erosion(w);
    smooth_pass(w, 0);
    pair_swap(&w->heightmap);
}

/* individual passes, for per-pass parity tests (same swaps as dispatch_grid) */
void orc_pass(orc_world* w, int pass) {
    switch (pass) {
    case ORC_PASS_FLUX: flux_pass(w); pair_swap(&w->heightmap); pair_swap(&w->flux); pair_swap(&w->velocity); break;
    case ORC_PASS_EROSION: erosion_pass(w); pair_swap(&w->heightmap); pair_swap(&w->sediment); break;
    case ORC_PASS_SEDIMENT: sediment_pass(w); pair_swap(&w->heightmap); pair_swap(&w->sediment); break;
    case ORC_PASS_THERMAL: run_thermal_erosion(w); break;
    case ORC_PASS_SMOOTH: smooth_pass(w, w->erosion_type == 1); pair_swap(&w->heightmap);
        if (w->erosion_type == 1) pair_swap(&w->velocity);
        break;
    default: break;
    }
}

/* ---------------------------------------------------------- particle.glsl -- */
/* particle.glsl:41-44 */
static inline float prand(vec2 p) {
    return hg_fract(1e4f * hg_sinf(17.0f * p.x + p.y * 0.1f) * (0.1f + fabsf(hg_sinf(p.y * 13.0f + p.x))));
}
/* particle.glsl:50-62 (bilinear samples, not texel loads) */
static vec3 terr_normal_bil(const float* hm, int W, int H, vec2 pos) {
    vec2 pr = {pos.x + 1.0f, pos.y + 0.0f}, pl = {pos.x + -1.0f, pos.y + 0.0f};
    vec2 pb = {pos.x + 0.0f, pos.y + -1.0f}, pt = {pos.x + 0.0f, pos.y + 1.0f};
    vec4 r = img_bilinear(hm, W, H, pr), l = img_bilinear(hm, W, H, pl);
    vec4 b = img_bilinear(hm, W, H, pb), t = img_bilinear(hm, W, H, pt);
    float dx = (r.x + r.y - l.x - l.y);
    float dz = (t.x + t.y - b.x - b.y);
    vec3 a = {2.0f, dx, 0.0f}, c = {0.0f, dz, 2.0f};
    vec3 n = {a.y * c.z - a.z * c.y, a.z * c.x - a.x * c.z, a.x * c.y - a.y * c.x};
    float inv = 1.0f / sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    n.x *= inv; n.y *= inv; n.z *= inv;
    return n;
}
/* particle.glsl:64-136 */
static void particle_move_pass(orc_world* w, int should_rain) {
    const int W = w->W, H = w->H;
    const hg_erosion_data set = w->erosion;
    const hg_map_settings_data map_set = w->map;
    const float* hm = rd(&w->heightmap);
    const float* mm = rd(&w->velocity);
    const float time = w->time;
    const int64_t count = (int64_t)(w->particle_count / 64u) * 64; /* erosion.cpp:127 integer division */
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < count; id++) {
        hg_particle p = w->particles[id];
        if (p.iters == 0 && !should_rain) continue;
        for (int i = 0; i < HG_SED_LAYERS; i++)
            if (p.sediment[i] < 0.0f || p.iters == 0) p.sediment[i] = 0.0f;
        if (p.iters == 0 || p.to_kill) {
            vec2 a = {hg_fract(time * 1.37f) * 1000.0f, (float)(uint32_t)id};
            vec2 b = {hg_fract(time * 7.21f) * 1000.0f, (float)(uint32_t)id + 3.14f};
            vec2 pos = {prand(a) * (float)((float)map_set.hmap_dims[0] - 4.0f) / 1.0f + 2.0f,
                        prand(b) * (float)((float)map_set.hmap_dims[1] - 4.0f) / 1.0f + 2.0f};
            p.to_kill = 0;
            p.position[0] = pos.x; p.position[1] = pos.y;
            p.velocity[0] = 0.0f; p.velocity[1] = 0.0f;
            p.volume = set.init_volume;
            if (should_rain) {
                p.iters = 1;
            } else {
                p.iters = 0;
                continue; /* `return` before the store: nothing is written (:84-87) */
            }
        }
        vec2 ppos = {p.position[0], p.position[1]};
        vec3 norm = terr_normal_bil(hm, W, H, ppos);
        vec4 mom4 = img_bilinear(mm, W, H, ppos);
        vec2 momentum = {mom4.x, mom4.y};
        float water = img_bilinear(hm, W, H, ppos).z;
        p.velocity[0] -= (set.d_t * norm.x) / (p.volume) * set.G;
        p.velocity[1] -= (set.d_t * norm.z) / (p.volume) * set.G;
        float lm = length2(momentum.x, momentum.y), lv = length2(p.velocity[0], p.velocity[1]);
        if (lm > 0.0f && lv > 0.0f) {
            float im = 1.0f / sqrtf(momentum.x * momentum.x + momentum.y * momentum.y);
            float iv = 1.0f / sqrtf(p.velocity[0] * p.velocity[0] + p.velocity[1] * p.velocity[1]);
            float dt = (momentum.x * im) * (p.velocity[0] * iv) + (momentum.y * im) * (p.velocity[1] * iv);
            float f = set.inertia * dt / (p.volume + 1e5f * water);
            p.velocity[0] += f * momentum.x;
            p.velocity[1] += f * momentum.y;
        }
        if (length2(p.velocity[0], p.velocity[1]) > 1.0f) {
            float iv = 1.0f / sqrtf(p.velocity[0] * p.velocity[0] + p.velocity[1] * p.velocity[1]);
            p.velocity[0] *= iv;
            p.velocity[1] *= iv;
        }
        vec2 old_pos = {p.position[0], p.position[1]};
        p.position[0] += set.d_t * p.velocity[0];
        p.position[1] += set.d_t * p.velocity[1];
        if (p.position[0] <= 1.0f || p.position[1] <= 1.0f
            || p.position[0] * 1.0f >= (float)(map_set.hmap_dims[0] - 2)
            || p.position[1] * 1.0f >= (float)(map_set.hmap_dims[1] - 2)) {
            p.position[0] = old_pos.x; p.position[1] = old_pos.y;
            p.velocity[0] = 0.0f; p.velocity[1] = 0.0f;
            p.to_kill = 1;
        }
        float fr = (1.0f - set.d_t * set.friction * norm.y);
        p.velocity[0] *= fr;
        p.velocity[1] *= fr;
        p.volume -= set.d_t * set.Ke;
        float sin_a = fabsf(fabsf(sqrtf(1.0f - norm.y * norm.y)));
        p.sc = hg_max(0.0f, set.Kc * p.volume * length2(p.velocity[0], p.velocity[1]) * hg_max(0.02f, sin_a));
        p.iters++;
        if (p.volume <= set.min_volume || length2(p.velocity[0], p.velocity[1]) < set.min_velocity
            || (uint32_t)p.iters >= set.ttl) {
            p.to_kill = 1;
        }
        w->particles[id] = p;
    }
}

/* particle_erosion.glsl:22-85, one corner of the droplet's quad, under the lock */
static void erode_layers(orc_world* w, int64_t id, int px, int py, vec2 offset, vec2 old_sediment) {
    const int W = w->W, H = w->H;
    const hg_erosion_data set = w->erosion;
    float* hm = rd(&w->heightmap);   /* read texture, modified in place (erosion.cpp:141-143) */
    float* mm = rd(&w->velocity);
    if (oob(W, H, px, py)) return;   /* image store out of bounds is a no-op */
    hg_particle part = w->particles[id];
    vec4 terr4 = ld4(hm, W, px, py);
    vec4 momentm = ld4(mm, W, px, py);
    float terr[2] = {terr4.x, terr4.y};
    float old_sed[2] = {old_sediment.x, old_sediment.y};
    float multipl = offset.x * offset.y;
    float cap = 0.0f;
    for (int i = HG_SED_LAYERS - 1; i >= 0; i--) {
        if (part.to_kill) {
            float sed = old_sed[i] * multipl;
            terr[i] += sed;
            part.sediment[i] -= sed;
            continue;
        }
        float Kls = set.d_t * set.Ks[i];
        float Kld = set.d_t * set.Kd[i];
        float c = hg_max(0.0f, part.sc - cap);
        float s1 = old_sed[i];
        float old_terr = terr[i];
        if (c > s1) {
            float eroded = multipl * Kls * (c - s1);
            s1 += eroded;
            terr[i] -= eroded;
            if (terr[i] < 0.0f) {
                s1 += terr[i];
                terr[i] = 0.0f;
                cap += old_terr;
            } else {
                part.sediment[i] = s1;
                break;
            }
        } else {
            float deposit = multipl * Kld * (s1 - c);
            s1 -= deposit;
            terr[i] += deposit;
        }
        part.sediment[i] = s1;
    }
    for (int i = 0; i < HG_SED_LAYERS - 1; i++) {
        float conv = part.sediment[i] * set.Kconv * set.d_t;
        part.sediment[i + 1] += conv;
        part.sediment[i] -= conv;
    }
    w->particles[id] = part;
    terr4.x = terr[0]; terr4.y = terr[1];
    terr4.z += 1e-5f * part.volume * multipl;
    terr4.w = terr4.x + terr4.z + terr4.y;
    momentm.z += part.volume * part.velocity[0] * multipl;
    momentm.w += part.volume * part.velocity[1] * multipl;
    st4(hm, W, px, py, terr4);
    st4(mm, W, px, py, momentm);
}

/* particle_erosion.glsl:101-128.  The reference serialises droplets that share
 * a texel with a spin lock, in whatever order the GPU grants it; the oracle
 * fixes that order to droplet id, corner 0..3. */
static void particle_erode_pass(orc_world* w) {
    const int64_t count = (int64_t)(w->particle_count / 64u) * 64;
    for (int64_t id = 0; id < count; id++) {
        hg_particle part = w->particles[id];
        if (part.iters == 0) continue;
        int bx = (int)(part.position[0] * 1.0f), by = (int)(part.position[1] * 1.0f);
        int pos[4][2] = {{bx, by}, {bx + 1, by}, {bx + 1, by + 1}, {bx, by + 1}};
        vec2 off = {hg_fract(part.position[0] * 1.0f), hg_fract(part.position[1] * 1.0f)};
        vec2 offset[4] = {{1.0f - off.x, 1.0f - off.y}, {off.x, 1.0f - off.y}, {off.x, off.y}, {1.0f - off.x, off.y}};
        vec2 sediment = {part.sediment[0], part.sediment[1]};
        for (int i = 0; i < 4; i++) erode_layers(w, id, pos[i][0], pos[i][1], offset[i], sediment);
    }
}

/* Erosion::dispatch_particle, src/erosion.cpp:132-156 */
void orc_dispatch_particle(orc_world* w, int should_rain) {
    particle_move_pass(w, should_rain);
    particle_erode_pass(w);
    run_thermal_erosion(w);
    smooth_pass(w, 1);
    pair_swap(&w->heightmap);
    pair_swap(&w->velocity);
}

void orc_particle_pass(orc_world* w, int which, int should_rain) {
    if (which == 0) particle_move_pass(w, should_rain);
    else particle_erode_pass(w);
}

/* the erosion part of the main loop, src/main.cpp:310-324; `time` is an input
 * (the reference reads the wall clock, main.cpp:290) */
void orc_step(orc_world* w, float time, int should_rain) {
    w->erosion_steps++;
    w->time = time;
    if (w->erosion_type == 0) {
        if (should_rain && w->rain.period != 0 && !(w->erosion_steps % (uint32_t)w->rain.period))
            orc_dispatch_grid_rain(w);
        orc_dispatch_grid(w);
    } else {
        orc_dispatch_particle(w, should_rain);
    }
}

/* ---------------------------------------------------------- world object --- */
/* State::World::gen_textures (src/state.cpp:3-44) + State::setup_settings
 * (src/state.cpp:57-106); textures start zeroed. */
static int pair_alloc(orc_pair* p, size_t n) {
    p->tex[0] = (float*)calloc(n, sizeof(float));
    p->tex[1] = (float*)calloc(n, sizeof(float));
    p->idx_read = 0; p->idx_write = 1; p->cntr = 0;
    return p->tex[0] && p->tex[1];
}
static void pair_free(orc_pair* p) { free(p->tex[0]); free(p->tex[1]); }

orc_world* orc_create(int W, int H, uint32_t particle_count, int erosion_type, float seed) {
    orc_world* w = (orc_world*)calloc(1, sizeof(orc_world));
    if (!w) return NULL;
    w->W = W; w->H = H;
    w->particle_count = particle_count;
    w->erosion_type = erosion_type;
    size_t n = (size_t)W * H * 4;
    int ok = pair_alloc(&w->heightmap, n) & pair_alloc(&w->flux, n) & pair_alloc(&w->velocity, n)
           & pair_alloc(&w->sediment, n) & pair_alloc(&w->thermal_c, n) & pair_alloc(&w->thermal_d, n);
    w->particles = (hg_particle*)calloc(particle_count ? particle_count : 1, sizeof(hg_particle));
    w->erosion = hg_default_erosion(erosion_type == 1, particle_count);
    w->rain = hg_default_rain();
    w->map = hg_default_map(seed);
    if (!ok || !w->particles) { orc_destroy(w); return NULL; }
    return w;
}
void orc_destroy(orc_world* w) {
    if (!w) return;
    pair_free(&w->heightmap); pair_free(&w->flux); pair_free(&w->velocity);
    pair_free(&w->sediment); pair_free(&w->thermal_c); pair_free(&w->thermal_d);
    free(w->particles);
    free(w);
}
static orc_pair* field_pair(orc_world* w, int field) {
    switch (field) {
    case 0: return &w->heightmap;
    case 1: return &w->flux;
    case 2: return &w->velocity;
    case 3: return &w->sediment;
    case 4: return &w->thermal_c;
    case 5: return &w->thermal_d;
    default: return NULL;
    }
}
float* orc_field(orc_world* w, int field) { orc_pair* p = field_pair(w, field); return p ? rd(p) : NULL; }
hg_particle* orc_particles(orc_world* w) { return w->particles; }
hg_erosion_data* orc_erosion(orc_world* w) { return &w->erosion; }
hg_rain_data* orc_rain(orc_world* w) { return &w->rain; }
hg_map_settings_data* orc_map(orc_world* w) { return &w->map; }
uint32_t orc_steps(orc_world* w) { return w->erosion_steps; }
void orc_set_steps(orc_world* w, uint32_t s) { w->erosion_steps = s; }
void orc_set_time(orc_world* w, float t) { w->time = t; }

/* defined-math probes for tests/test_defined_math.py */
float orc_atanf(float x) { return hg_atanf(x); }
float orc_expf(float x) { return hg_expf(x); }
float orc_sinf(float x) { return hg_sinf(x); }
float orc_simplex(float x, float y) { vec2 v = {x, y}; return gln_simplex(v); }
void orc_noised(float x, float y, float* out3) { vec2 v = {x, y}; vec3 r = noised(v); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; }
