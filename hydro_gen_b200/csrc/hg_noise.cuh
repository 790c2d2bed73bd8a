// hg_noise.cuh — the noise functions reachable from rain.glsl and heightmap.glsl
// (glsl/simplex_noise.glsl:247-256, 320, 377-456, 580-693), on scalars.
// Integer hashing wraps mod 2^32 like GLSL int; all float work is +,-,*,/,floor in
// the shader's association, so the result is bit-reproducible (no FMA, see hg_cell.cuh).
#pragma once
#include "hg_cell.cuh"

struct HgFbm {
    float seed, persistance, lacunarity, scale;
    int octaves;
    bool turbulence, ridge;   // gln_tFBMOpts.terbulance / .ridge
};

// mod(x, 289) = x - 289*floor(x/289).  On the device the IEEE quotient comes from one multiply
// and two FMAs (q = x*c; r = fma(-289, q, x); q + r*c with c = RN(1/289)), bit-identical to x/289
// for every finite float (exhaustive proof: scripts/check_div_const.c); the host build keeps the
// defining formula and the GPU parity tests compare the two.
HG_FN float hg_mod289(float x) {
#if HG_DEVICE_FAST
    const float c = 1.0f / 289.0f;
    const float q = __fmul_rn(x, c);
    const float q2 = __fmaf_rn(__fmaf_rn(-289.0f, q, x), c, q);
    return x - 289.0f * floorf(q2);
#else
    return hg_mod(x, 289.0f);
#endif
}

// gln_rand3 == _permute: mod(((p*34)+1)*p, 289)      simplex_noise.glsl:320
HG_FN float hg_permute(float p) { return hg_mod289(((p * 34.0f) + 1.0f) * p); }

// gln_simplex                                          simplex_noise.glsl:377-402
HG_FN float hg_simplex(float vx, float vy) {
    const float Cx = 0.211324865405187f, Cy = 0.366025403784439f;
    const float Cz = -0.577350269189626f, Cw = 0.024390243902439f;
    float s = vx * Cy + vy * Cy;
    float ix = floorf(vx + s), iy = floorf(vy + s);
    float t = ix * Cx + iy * Cx;
    float x0x = vx - ix + t, x0y = vy - iy + t;
    float i1x = (x0x > x0y) ? 1.0f : 0.0f;
    float i1y = (x0x > x0y) ? 0.0f : 1.0f;
    float x1x = x0x + Cx, x1y = x0y + Cx, x2x = x0x + Cz, x2y = x0y + Cz;
    x1x -= i1x;
    x1y -= i1y;
    ix = hg_mod289(ix);
    iy = hg_mod289(iy);
    // i1y is 0 or 1, so the inner permute of the middle corner is one of the other two (same expression, same bits)
    const float in0 = hg_permute(iy + 0.0f), in2 = hg_permute(iy + 1.0f);
    float p0 = hg_permute(in0 + ix + 0.0f);
    float p1 = hg_permute(((x0x > x0y) ? in0 : in2) + ix + i1x);
    float p2 = hg_permute(in2 + ix + 1.0f);
    float m0 = hg_max(0.5f - (x0x * x0x + x0y * x0y), 0.0f);
    float m1 = hg_max(0.5f - (x1x * x1x + x1y * x1y), 0.0f);
    float m2 = hg_max(0.5f - (x2x * x2x + x2y * x2y), 0.0f);
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    float q0 = 2.0f * hg_fract(p0 * Cw) - 1.0f;
    float q1 = 2.0f * hg_fract(p1 * Cw) - 1.0f;
    float q2 = 2.0f * hg_fract(p2 * Cw) - 1.0f;
    float h0 = fabsf(q0) - 0.5f, h1 = fabsf(q1) - 0.5f, h2 = fabsf(q2) - 0.5f;
    float a0 = q0 - floorf(q0 + 0.5f), a1 = q1 - floorf(q1 + 0.5f), a2 = q2 - floorf(q2 + 0.5f);
    m0 *= 1.79284291400159f - 0.85373472095314f * (a0 * a0 + h0 * h0);
    m1 *= 1.79284291400159f - 0.85373472095314f * (a1 * a1 + h1 * h1);
    m2 *= 1.79284291400159f - 0.85373472095314f * (a2 * a2 + h2 * h2);
    float g0 = a0 * x0x + h0 * x0y;
    float g1 = a1 * x1x + h1 * x1y;
    float g2 = a2 * x2x + h2 * x2y;
    return 130.0f * (m0 * g0 + m1 * g1 + m2 * g2);
}

// ---- table form of gln_simplex for the rain kernel --------------------------------------------------------------
// Every argument of _permute in gln_simplex is an integer-valued float in [0, 579] (lattice coordinates mod 289, plus
// a permuted value < 289, plus 0 or 1), so the five permutes of an evaluation are lookups in a table of 580 entries
// that each CTA fills in shared memory with hg_permute itself: the same bits, 2 conversions + 2 integer mod 289 + 5
// loads instead of 7 x (3 multiply-adds + a 6-instruction mod 289).  The permute chain was 40 % of k_rain's
// instructions (profiles/r01i_all_kernels.txt: 92 % of issue slots busy).
constexpr int HG_PERM_N = 580;
// The gradient of a corner depends on its hash alone: (a, h) = the two components gln_simplex derives from
// p = permute(k) and the normalisation factor 1.79284291400159 - 0.85373472095314 (a a + h h), formed by the same
// operations in the same order (hg_simplex_grad), so the 12 operations per corner become one 16-byte load.
struct HgGrad { float a, h, nrm, pad; };
HG_FN HgGrad hg_simplex_grad(float p) {
    const float Cw = 0.024390243902439f;
    const float q = 2.0f * hg_fract(p * Cw) - 1.0f;
    HgGrad g;
    g.h = fabsf(q) - 0.5f;
    g.a = q - floorf(q + 0.5f);
    g.nrm = 1.79284291400159f - 0.85373472095314f * (g.a * g.a + g.h * g.h);
    g.pad = 0.0f;
    return g;
}
struct HgPermTab { const int* ti; const float* tf; const HgGrad* tg; };      // permute(k) as int and as float; gradient of permute(k) (null: computed)
HG_FN int hg_imod289(int v) { int r = v % 289; return r < 0 ? r + 289 : r; }
HG_FN float hg_simplex_tab(float vx, float vy, const HgPermTab& T) {
    const float Cx = 0.211324865405187f, Cy = 0.366025403784439f;
    const float Cz = -0.577350269189626f, Cw = 0.024390243902439f;
    float s = vx * Cy + vy * Cy;
    float ix = floorf(vx + s), iy = floorf(vy + s);
    float t = ix * Cx + iy * Cx;
    float x0x = vx - ix + t, x0y = vy - iy + t;
    const bool hi = x0x > x0y;
    float i1x = hi ? 1.0f : 0.0f;
    float i1y = hi ? 0.0f : 1.0f;
    float x1x = x0x + Cx, x1y = x0y + Cx, x2x = x0x + Cz, x2y = x0y + Cz;
    x1x -= i1x;
    x1y -= i1y;
    // lattice coordinates beyond +-2^30 (never reached by map coordinates times a noise scale) take the float path
    if (!(fabsf(ix) < 1.0e9f && fabsf(iy) < 1.0e9f)) return hg_simplex(vx, vy);
    const int jx = hg_imod289((int)ix), jy = hg_imod289((int)iy);      // == mod289(ix), mod289(iy): exact integers
    const int in0 = T.ti[jy], in2 = T.ti[jy + 1];
    const int k0 = in0 + jx, k1 = (hi ? in0 : in2) + jx + (hi ? 1 : 0), k2 = in2 + jx + 1;
    float m0 = hg_max(0.5f - (x0x * x0x + x0y * x0y), 0.0f);
    float m1 = hg_max(0.5f - (x1x * x1x + x1y * x1y), 0.0f);
    float m2 = hg_max(0.5f - (x2x * x2x + x2y * x2y), 0.0f);
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    if (T.tg) {
        const HgGrad G0 = T.tg[k0], G1 = T.tg[k1], G2 = T.tg[k2];
        m0 *= G0.nrm; m1 *= G1.nrm; m2 *= G2.nrm;
        const float g0 = G0.a * x0x + G0.h * x0y;
        const float g1 = G1.a * x1x + G1.h * x1y;
        const float g2 = G2.a * x2x + G2.h * x2y;
        return 130.0f * (m0 * g0 + m1 * g1 + m2 * g2);
    }
    float p0 = T.tf[k0], p1 = T.tf[k1], p2 = T.tf[k2];
    float q0 = 2.0f * hg_fract(p0 * Cw) - 1.0f;
    float q1 = 2.0f * hg_fract(p1 * Cw) - 1.0f;
    float q2 = 2.0f * hg_fract(p2 * Cw) - 1.0f;
    float h0 = fabsf(q0) - 0.5f, h1 = fabsf(q1) - 0.5f, h2 = fabsf(q2) - 0.5f;
    float a0 = q0 - floorf(q0 + 0.5f), a1 = q1 - floorf(q1 + 0.5f), a2 = q2 - floorf(q2 + 0.5f);
    m0 *= 1.79284291400159f - 0.85373472095314f * (a0 * a0 + h0 * h0);
    m1 *= 1.79284291400159f - 0.85373472095314f * (a1 * a1 + h1 * h1);
    m2 *= 1.79284291400159f - 0.85373472095314f * (a2 * a2 + h2 * h2);
    float g0 = a0 * x0x + h0 * x0y;
    float g1 = a1 * x1x + h1 * x1y;
    float g2 = a2 * x2x + h2 * x2y;
    return 130.0f * (m0 * g0 + m1 * g1 + m2 * g2);
}
HG_FN float hg_sfbm_tab(float vx, float vy, const HgFbm& o, const HgPermTab& T) {
    vx += (o.seed * 100.0f);
    vy += (o.seed * 100.0f);
    bool ridge = o.turbulence && o.ridge;
    float result = 0.0f, amplitude = 1.0f, frequency = 1.0f, maximum = amplitude;
    for (int i = 0; i < 30; i++) {
        if (i >= o.octaves) break;
        float n = hg_simplex_tab(vx * frequency * o.scale, vy * frequency * o.scale, T);
        if (o.turbulence) n = fabsf(n);
        if (ridge) n = 1.0f - n;
        result += n * amplitude;
        frequency *= o.lacunarity;
        amplitude *= o.persistance;
        maximum += amplitude;
    }
    return result / maximum;
}

// gln_sfbm; pow(result, 1.0) is the identity      simplex_noise.glsl:419-456
HG_FN float hg_sfbm(float vx, float vy, const HgFbm& o) {
    vx += (o.seed * 100.0f);
    vy += (o.seed * 100.0f);
    bool ridge = o.turbulence && o.ridge;
    float result = 0.0f, amplitude = 1.0f, frequency = 1.0f, maximum = amplitude;
    for (int i = 0; i < 30; i++) {
        if (i >= o.octaves) break;
        float n = hg_simplex(vx * frequency * o.scale, vy * frequency * o.scale);
        if (o.turbulence) n = fabsf(n);
        if (ridge) n = 1.0f - n;
        result += n * amplitude;
        frequency *= o.lacunarity;
        amplitude *= o.persistance;
        maximum += amplitude;
    }
    return result / maximum;
}

// hash(ivec2), one component                       simplex_noise.glsl:580-588
HG_FN float hg_ihash1(uint32_t n) {
    n = (n << 13) ^ n;
    n = n * (n * n * 15731u + 789221u) + 1376312589u;
    return -1.0f + 2.0f * (float)(int32_t)(n & 0x0fffffffu) / (float)0x0fffffff;
}

// noised: value and analytic derivatives          simplex_noise.glsl:591-612
HG_FN void hg_noised(float px, float py, float& val, float& ddx, float& ddy) {
    float flx = floorf(px), fly = floorf(py);
    uint32_t ix = (uint32_t)(int32_t)flx, iy = (uint32_t)(int32_t)fly;
    float fx = px - flx, fy = py - fly;
    float ux = fx * fx * fx * (fx * (fx * 6.0f - 15.0f) + 10.0f);
    float uy = fy * fy * fy * (fy * (fy * 6.0f - 15.0f) + 10.0f);
    float dux = 30.0f * fx * fx * (fx * (fx - 2.0f) + 1.0f);
    float duy = 30.0f * fy * fy * (fy * (fy - 2.0f) + 1.0f);
    uint32_t ix1 = ix + 1u, iy1 = iy + 1u;
    float gax = hg_ihash1(ix * 3u + iy * 311u),   gay = hg_ihash1(ix * 37u + iy * 113u);
    float gbx = hg_ihash1(ix1 * 3u + iy * 311u),  gby = hg_ihash1(ix1 * 37u + iy * 113u);
    float gcx = hg_ihash1(ix * 3u + iy1 * 311u),  gcy = hg_ihash1(ix * 37u + iy1 * 113u);
    float gdx = hg_ihash1(ix1 * 3u + iy1 * 311u), gdy = hg_ihash1(ix1 * 37u + iy1 * 113u);
    float va = gax * (fx - 0.0f) + gay * (fy - 0.0f);
    float vb = gbx * (fx - 1.0f) + gby * (fy - 0.0f);
    float vc = gcx * (fx - 0.0f) + gcy * (fy - 1.0f);
    float vd = gdx * (fx - 1.0f) + gdy * (fy - 1.0f);
    float k = va - vb - vc + vd;
    val = va + ux * (vb - va) + uy * (vc - va) + ux * uy * k;
    ddx = gax + ux * (gbx - gax) + uy * (gcx - gax) + ux * uy * (gax - gbx - gcx + gdx) + dux * (uy * k + vb - va);
    ddy = gay + ux * (gby - gay) + uy * (gcy - gay) + ux * uy * (gay - gby - gcy + gdy) + duy * (ux * k + vc - va);
}

// perlfbm                                          simplex_noise.glsl:614-652
HG_FN float hg_perlfbm(float vx, float vy, const HgFbm& o) {
    vx += (o.seed * 100.0f);
    vy += (o.seed * 100.0f);
    bool ridge = o.turbulence && o.ridge;
    float result = 0.0f, amplitude = 1.0f, frequency = 1.0f, maximum = amplitude;
    for (int i = 0; i < 30; i++) {
        if (i >= o.octaves) break;
        float v, dx, dy;
        hg_noised(vx * frequency * o.scale, vy * frequency * o.scale, v, dx, dy);
        float n = (v + 1.0f) / 2.0f;
        if (o.turbulence) n = fabsf(n);
        if (ridge) n = 1.0f - n;
        result += n * amplitude;
        frequency *= o.lacunarity;
        amplitude *= o.persistance;
        maximum += amplitude;
    }
    return result / maximum;
}

// erosion_perlfbm: amplitude damped by the running gradient   simplex_noise.glsl:654-693
HG_FN float hg_erosion_perlfbm(float vx, float vy, const HgFbm& o) {
    vx += (o.seed * 100.0f);
    vy += (o.seed * 100.0f);
    bool ridge = o.turbulence && o.ridge;
    float result = 0.0f, amplitude = 1.0f, frequency = 1.5f, maximum = amplitude;
    float dsx = 0.0f, dsy = 0.0f;
    for (int i = 0; i < 30; i++) {
        if (i >= o.octaves) break;
        float v, dx, dy;
        hg_noised(vx * frequency * o.scale, vy * frequency * o.scale, v, dx, dy);
        if (o.turbulence) { v = fabsf(v); dx = fabsf(dx); dy = fabsf(dy); }
        if (ridge) { v = 1.0f - v; dx = 1.0f - dx; dy = 1.0f - dy; }
        dsx += dx;
        dsy += dy;
        float n = (v + 1.0f) / 2.0f;
        result += n * amplitude / (1.0f + (dsx * dsx + dsy * dsy));
        frequency *= o.lacunarity;
        amplitude *= o.persistance;
        maximum += amplitude;
    }
    return result / maximum;
}

// One cell of heightmap.glsl:52-158.  x,y global cell; W,H = imageSize (the map).
HG_FN void hg_heightmap_cell(const hg_map_settings_data& cfg, int x, int y, int W, int H, float& rock, float& dirt) {
    float uvx = (float)x / (float)W, uvy = (float)y / (float)H;
    HgFbm opts{cfg.seed, cfg.persistance, cfg.lacunarity, cfg.scale, cfg.octaves, false, false};
    float distx = 1.0f, disty = 1.0f;
    if (cfg.domain_warp != 0) {
        distx = hg_perlfbm((float)x + 2.3f, (float)y + 2.9f, opts);
        disty = hg_perlfbm((float)x - 3.1f, (float)y - 4.3f, opts);
        if (cfg.domain_warp == 2) {
            float ax = (float)x + cfg.domain_warp_scale * distx - 5.7f, ay = (float)y + cfg.domain_warp_scale * disty + 27.9f;
            float bx = (float)x + cfg.domain_warp_scale * distx + 11.5f, by = (float)y + cfg.domain_warp_scale * disty - 23.7f;
            float nx = hg_perlfbm(ax, ay, opts), ny = hg_perlfbm(bx, by, opts);
            distx = nx; disty = ny;
        }
    }
    float val = hg_erosion_perlfbm((float)x + cfg.domain_warp_scale * distx, (float)y + cfg.domain_warp_scale * disty, opts);
    float hm = cfg.height_mult;
    if (cfg.uplift != 0) {
        HgFbm up{cfg.seed, cfg.persistance, cfg.lacunarity, cfg.scale / cfg.uplift_scale, cfg.octaves, true, true};
        float upv = hg_sfbm((float)x - 7.3f, (float)y + 19.9f, up);
        if (cfg.mask_exp != 0) hm += 2.5f;
        hm += 1.0f;
        val *= upv;
    }
    if (cfg.mask_round != 0) {                       // round_mask, heightmap.glsl:29-33
        if (cfg.mask_exp != 0) hm += 16.0f;
        float a = uvx - 0.5f, b = uvy - 0.5f;
        float v = 1.0f - (a * a + b * b + 0.75f);
        val = val * hg_max(0.0f, v * 1.0f);
    }
    if (cfg.mask_exp != 0) {                         // exp_mask, heightmap.glsl:43-45
        hm += 2.0f;
        val = val * (hg_expf(val) - 1.0f) / 1.718f;
    }
    if (cfg.mask_power != 0) {                       // power_mask, heightmap.glsl:39-41
        hm += 1.25f;
        if (cfg.mask_exp != 0) hm += 2.0f;
        float b = val + 0.5f;
        float p3 = b * b * b;
        val = val * (((p3 - 0.125f) / 3.25f) * 0.55f + 0.45f);
    }
    if (cfg.mask_slope != 0) val = 0.25f * val + 0.75f * (val * uvx * uvy);   // slope_mask, :35-37
    val = val + (float)(4u * cfg.mask_round) * val;
    if (cfg.terrace > 0) {
        float lv = (float)cfg.terrace * hm;
        float lol = floorf(val / (1.0f / lv));
        float ts = cfg.terrace_scale;
        val = (ts * lol * (1.0f / lv)) + val * (1.0f - ts);
    }
    rock = hg_min(cfg.max_height, val * cfg.max_height * hm);
    dirt = hg_perlfbm((float)x + 13.7f, (float)y + 27.1f, opts) + 1.5f;
    dirt *= cfg.max_dirt;
}

// One cell of rain.glsl:32-56.  `total` is H.a as stored; returns the water increment.  T: null, or the permute tables.
HG_FN float hg_rain_cell(const hg_rain_data& set, const hg_map_settings_data& map_set, float time, int x, int y, float total, const HgPermTab* T = nullptr) {
    HgFbm opts{hg_fract(time * 1.372914227e3f) * 1000.f, 0.5f, 2.0f, set.drops, 8, false, false};
    float r = hg_max(0.0f, T ? hg_sfbm_tab((float)x, (float)y, opts, *T) : hg_sfbm((float)x, (float)y, opts));
    float incr = set.amount * r;
    float mountain = total - map_set.max_height * set.mountain_thresh;
    if (mountain > 0.0f) {
        incr += mountain * set.mountain_multip * r / ((1.0f - set.mountain_thresh) * map_set.max_height);
    }
    return incr;
}
