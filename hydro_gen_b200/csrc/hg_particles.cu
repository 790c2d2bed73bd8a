// hg_particles.cu — droplet (particle) erosion: Erosion::dispatch_particle's two droplet
// dispatches (src/erosion.cpp:133-144 -> glsl/particle.glsl, glsl/particle_erosion.glsl).
//
// The reference serialises droplets that touch the same texel with a per-pixel CAS spin
// lock around a read-modify-write of H and M (particle_erosion.glsl:89-99).  Here there is
// no lock.  Everything a corner update adds to a texel is independent of the texel's value
// except the "layer exhausted" clamp (particle_erosion.glsl:55-59), so:
//   * deposits, display water and the momentum accumulator are fire-and-forget float
//     reductions (red.global.add.f32), first combined across the lanes of a warp that hit
//     the same texel (__match_any_sync) so contended texels see one atomic per warp;
//   * erosion is a compare-and-swap loop on the one layer value, which makes the clamp
//     exact and linearisable without ever blocking another droplet.
// The droplet's own state (sediment carried from corner to corner, the Kconv conversion
// applied once per corner, the last-corner-wins quirk of SURVEY.md §8a P2) is carried in
// registers in corner order 0..3, exactly as the shader re-reads and re-writes its SSBO
// element.  Order between droplets is free, as it is in the reference.
#include "hg_internal.cuh"

namespace {

struct PDom { int W, H, pitch; };
__device__ __forceinline__ size_t pidx(const PDom& d, int x, int y) { return (size_t)(y + HG_HALO_ROWS) * d.pitch + x; }
__device__ __forceinline__ float fetch(const float* __restrict__ p, const PDom& d, int x, int y) {
    return (x < 0 || x > d.W - 1 || y < 0 || y > d.H - 1) ? 0.0f : __ldg(p + pidx(d, x, y));
}
// img_bilinear of one channel at a float position (img_interpolation.glsl:3-22)
__device__ __forceinline__ float bilinear(const float* __restrict__ p, const PDom& d, float sx, float sy) {
    if (!(sx == sx)) sx = 0.0f;
    if (!(sy == sy)) sy = 0.0f;
    int px = (int)sx, py = (int)sy;
    float fx = hg_fract(sx), fy = hg_fract(sy);
    return hg_bilerp(fetch(p, d, px, py), fetch(p, d, px + 1, py), fetch(p, d, px, py + 1), fetch(p, d, px + 1, py + 1), fx, fy);
}

// rand(vec2), particle.glsl:41-44
__device__ __forceinline__ float prand(float px, float py) {
    return hg_fract(1e4f * hg_sinf(17.0f * px + py * 0.1f) * (0.1f + fabsf(hg_sinf(py * 13.0f + px))));
}

struct MoveArgs { const float *rock, *dirt, *water, *mx, *my; hg_particle* particles; };

// particle.glsl:64-136
__global__ void __launch_bounds__(128) k_particle_move(PDom d, HgStepParams P, hg_erosion_data set, hg_map_settings_data map_set,
                                                       MoveArgs A, uint32_t count, float time, int should_rain) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    hg_particle p = A.particles[id];
    if (p.iters == 0 && !should_rain) return;
#pragma unroll
    for (int i = 0; i < HG_SED_LAYERS; i++)
        if (p.sediment[i] < 0.0f || p.iters == 0) p.sediment[i] = 0.0f;
    if (p.iters == 0 || p.to_kill) {
        float posx = prand(hg_fract(time * 1.37f) * 1000.0f, (float)id) * (float)((float)map_set.hmap_dims[0] - 4.0f) / 1.0f + 2.0f;
        float posy = prand(hg_fract(time * 7.21f) * 1000.0f, (float)id + 3.14f) * (float)((float)map_set.hmap_dims[1] - 4.0f) / 1.0f + 2.0f;
        p.to_kill = 0;
        p.position[0] = posx; p.position[1] = posy;
        p.velocity[0] = 0.0f; p.velocity[1] = 0.0f;
        p.volume = set.init_volume;
        if (should_rain) {
            p.iters = 1;
        } else {
            p.iters = 0;
            return;
        }
    }
    float px = p.position[0], py = p.position[1];
    // get_terr_normal, particle.glsl:50-62 (bilinear samples)
    float rr = bilinear(A.rock, d, px + 1.0f, py + 0.0f), rg = bilinear(A.dirt, d, px + 1.0f, py + 0.0f);
    float lr = bilinear(A.rock, d, px + -1.0f, py + 0.0f), lg = bilinear(A.dirt, d, px + -1.0f, py + 0.0f);
    float br = bilinear(A.rock, d, px + 0.0f, py + -1.0f), bg = bilinear(A.dirt, d, px + 0.0f, py + -1.0f);
    float tr = bilinear(A.rock, d, px + 0.0f, py + 1.0f), tg = bilinear(A.dirt, d, px + 0.0f, py + 1.0f);
    float dx = (rr + rg - lr - lg);
    float dz = (tr + tg - br - bg);
    float nx = dx * 2.0f - 0.0f * dz, ny = 0.0f * 0.0f - 2.0f * 2.0f, nz = 2.0f * dz - dx * 0.0f;
    float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    nx *= inv; ny *= inv; nz *= inv;
    float momx = bilinear(A.mx, d, px, py), momy = bilinear(A.my, d, px, py);
    float water = bilinear(A.water, d, px, py);

    p.velocity[0] -= (set.d_t * nx) / (p.volume) * set.G;
    p.velocity[1] -= (set.d_t * nz) / (p.volume) * set.G;
    float lm = sqrtf(momx * momx + momy * momy), lv = sqrtf(p.velocity[0] * p.velocity[0] + p.velocity[1] * p.velocity[1]);
    if (lm > 0.0f && lv > 0.0f) {
        float im = 1.0f / sqrtf(momx * momx + momy * momy);
        float iv = 1.0f / sqrtf(p.velocity[0] * p.velocity[0] + p.velocity[1] * p.velocity[1]);
        float dt = (momx * im) * (p.velocity[0] * iv) + (momy * im) * (p.velocity[1] * iv);
        float f = set.inertia * dt / (p.volume + 1e5f * water);
        p.velocity[0] += f * momx;
        p.velocity[1] += f * momy;
    }
    if (sqrtf(p.velocity[0] * p.velocity[0] + p.velocity[1] * p.velocity[1]) > 1.0f) {
        float iv = 1.0f / sqrtf(p.velocity[0] * p.velocity[0] + p.velocity[1] * p.velocity[1]);
        p.velocity[0] *= iv;
        p.velocity[1] *= iv;
    }
    float oldx = p.position[0], oldy = p.position[1];
    p.position[0] += set.d_t * p.velocity[0];
    p.position[1] += set.d_t * p.velocity[1];
    if (p.position[0] <= 1.0f || p.position[1] <= 1.0f
        || p.position[0] * 1.0f >= (float)(map_set.hmap_dims[0] - 2)
        || p.position[1] * 1.0f >= (float)(map_set.hmap_dims[1] - 2)) {
        p.position[0] = oldx; p.position[1] = oldy;
        p.velocity[0] = 0.0f; p.velocity[1] = 0.0f;
        p.to_kill = 1;
    }
    float fr = (1.0f - set.d_t * set.friction * ny);
    p.velocity[0] *= fr;
    p.velocity[1] *= fr;
    p.volume -= set.d_t * set.Ke;
    float sin_a = fabsf(fabsf(sqrtf(1.0f - ny * ny)));
    float speed = sqrtf(p.velocity[0] * p.velocity[0] + p.velocity[1] * p.velocity[1]);
    p.sc = hg_max(0.0f, set.Kc * p.volume * speed * hg_max(0.02f, sin_a));
    p.iters++;
    if (p.volume <= set.min_volume || speed < set.min_velocity || (uint32_t)p.iters >= set.ttl) p.to_kill = 1;
    A.particles[id] = p;
}

struct ErodePlanes { float *rock, *dirt, *water, *mz, *mw; };

// What one corner of one droplet adds to its texel (particle_erosion.glsl:61-75,94-98): deposits per layer
// (only where the layer deposited), display water and the momentum accumulator.
struct CornerAdd { float rock, dirt, water, mz, mw; bool has_rock, has_dirt; };

// Apply the corner additions of all 32 lanes.  Lanes of the warp that hit the same texel are found with ONE
// match per corner; the lowest such lane sums its peers' five values in lane order (a lane that has no deposit
// for a layer contributes its 0.0f, which changes no sum) and issues one red per plane.  All 32 lanes must call.
__device__ __forceinline__ void warp_corner_add(const ErodePlanes& A, size_t ti, const CornerAdd& v, bool act) {
    const int lane = threadIdx.x & 31;
    unsigned long long key = act ? (unsigned long long)ti : ~0ull - lane;
    unsigned peers = __match_any_sync(0xffffffffu, key);
    unsigned rock_lanes = __ballot_sync(0xffffffffu, act && v.has_rock), dirt_lanes = __ballot_sync(0xffffffffu, act && v.has_dirt);
    if (!act) return;
    if (__popc(peers) == 1) {              // the common case on a sparse map: nobody else in the warp hits this texel
        if (v.has_dirt) atomicAdd(A.dirt + ti, v.dirt);
        if (v.has_rock) atomicAdd(A.rock + ti, v.rock);
        atomicAdd(A.water + ti, v.water);
        atomicAdd(A.mz + ti, v.mz);
        atomicAdd(A.mw + ti, v.mw);
        return;
    }
    // peers > 1 only occurs among active lanes (inactive keys are unique)
    float s_rock = 0.0f, s_dirt = 0.0f, s_water = 0.0f, s_mz = 0.0f, s_mw = 0.0f;
    unsigned rest = peers;
    while (rest) {
        int src = __ffs(rest) - 1;
        rest &= rest - 1;
        s_rock += __shfl_sync(peers, v.rock, src);
        s_dirt += __shfl_sync(peers, v.dirt, src);
        s_water += __shfl_sync(peers, v.water, src);
        s_mz += __shfl_sync(peers, v.mz, src);
        s_mw += __shfl_sync(peers, v.mw, src);
    }
    if (lane == __ffs(peers) - 1) {
        if (dirt_lanes & peers) atomicAdd(A.dirt + ti, s_dirt);
        if (rock_lanes & peers) atomicAdd(A.rock + ti, s_rock);
        atomicAdd(A.water + ti, s_water);
        atomicAdd(A.mz + ti, s_mz);
        atomicAdd(A.mw + ti, s_mw);
    }
}

// terr -= eroded with the exhausted-layer clamp of particle_erosion.glsl:53-59, as one
// linearisable update.  Returns the value before the update; *unclamped = old - eroded.
__device__ __forceinline__ float erode_clamped(float* addr, float eroded, float* unclamped) {
    unsigned* ua = reinterpret_cast<unsigned*>(addr);
    unsigned assumed, old = *reinterpret_cast<volatile unsigned*>(ua);
    float nv;
    do {
        assumed = old;
        float cur = __uint_as_float(assumed);
        nv = cur - eroded;
        float store = (nv < 0.0f) ? 0.0f : nv;
        old = atomicCAS(ua, assumed, __float_as_uint(store));
    } while (old != assumed);
    *unclamped = nv;
    return __uint_as_float(assumed);
}

struct ErodeArgs : ErodePlanes { hg_particle* particles; };

// particle_erosion.glsl:101-128 + erode_layers :22-85
__global__ void __launch_bounds__(128) k_particle_erode(PDom d, HgStepParams P, ErodeArgs A, uint32_t count) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = id < count;
    hg_particle part = {};
    if (live) part = A.particles[id];
    live = live && part.iters != 0;
    // all lanes stay in the loop: warp_corner_add is warp-collective
    int bx = 0, by = 0;
    float offx = 0.0f, offy = 0.0f, old_sed[2] = {0.0f, 0.0f};
    if (live) {
        bx = (int)(part.position[0] * 1.0f); by = (int)(part.position[1] * 1.0f);
        offx = hg_fract(part.position[0] * 1.0f); offy = hg_fract(part.position[1] * 1.0f);
        old_sed[0] = part.sediment[0]; old_sed[1] = part.sediment[1];
    }
    //  3---2
    //  |   |
    //  0---1
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int cx = bx + ((k == 1 || k == 2) ? 1 : 0), cy = by + ((k >= 2) ? 1 : 0);
        float wx = (k == 1 || k == 2) ? offx : 1.0f - offx;
        float wy = (k >= 2) ? offy : 1.0f - offy;
        bool act = live && !(cx < 0 || cx > d.W - 1 || cy < 0 || cy > d.H - 1);
        size_t ti = act ? pidx(d, cx, cy) : 0;
        float multipl = wx * wy;
        float dep[2] = {0.0f, 0.0f};      // value-independent additions to terrain, per layer
        bool has_dep[2] = {false, false};
        if (act) {
            float cap = 0.0f;
#pragma unroll
            for (int i = HG_SED_LAYERS - 1; i >= 0; i--) {
                if (part.to_kill) {
                    float sed = old_sed[i] * multipl;
                    dep[i] = sed; has_dep[i] = true;
                    part.sediment[i] -= sed;
                    continue;
                }
                float c = hg_max(0.0f, part.sc - cap);
                float s1 = old_sed[i];
                if (c > s1) {
                    float eroded = multipl * P.Kls[i] * (c - s1);
                    s1 += eroded;
                    float after;
                    float old_terr = erode_clamped((i == 0 ? A.rock : A.dirt) + ti, eroded, &after);
                    if (after < 0.0f) {
                        s1 += after;
                        cap += old_terr;
                    } else {
                        part.sediment[i] = s1;
                        break;
                    }
                } else {
                    float deposit = multipl * P.Kld[i] * (s1 - c);
                    s1 -= deposit;
                    dep[i] = deposit; has_dep[i] = true;
                }
                part.sediment[i] = s1;
            }
            float conv = part.sediment[0] * P.Kconv * P.d_t;
            part.sediment[1] += conv;
            part.sediment[0] -= conv;
        }
        CornerAdd v;
        v.rock = dep[0]; v.dirt = dep[1]; v.has_rock = has_dep[0]; v.has_dirt = has_dep[1];
        v.water = 1e-5f * part.volume * multipl;
        v.mz = part.volume * part.velocity[0] * multipl;
        v.mw = part.volume * part.velocity[1] * multipl;
        warp_corner_add(A, ti, v, act);
    }
    if (live) A.particles[id] = part;
}

}  // namespace

int hg_launch_particle_move(hg_ctx* c, float time, int should_rain) {
    PDom d{c->g.W, c->g.H, c->g.pitch};
    uint32_t count = (c->particle_count / 64u) * 64u;    // glDispatchCompute(particle_count/64), erosion.cpp:127
    if (!count) return HG_OK;
    MoveArgs A{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_vel(c, 0, 1), hg_vel(c, 1, 1), c->particles};
    k_particle_move<<<(count + 127) / 128, 128, 0, c->stream>>>(d, c->sp, c->erosion, c->map, A, count, time, should_rain);
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}

int hg_launch_particle_erode(hg_ctx* c) {
    PDom d{c->g.W, c->g.H, c->g.pitch};
    uint32_t count = (c->particle_count / 64u) * 64u;
    if (!count) return HG_OK;
    // in place on the READ images of heightmap and momentum map (erosion.cpp:141-143)
    ErodeArgs A;
    A.rock = hg_cur(c, PL_ROCK, 1); A.dirt = hg_cur(c, PL_DIRT, 1); A.water = hg_cur(c, PL_WATER, 1);
    A.mz = hg_vel(c, 2, 1); A.mw = hg_vel(c, 3, 1); A.particles = c->particles;
    k_particle_erode<<<(count + 127) / 128, 128, 0, c->stream>>>(d, c->sp, A, count);
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}
