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
//
// Layout.  Both kernels touch single texels at data-dependent positions, so they work on the heightmap and the
// momentum map in the reference's own TEXTURE layout, one 16-byte texel per cell ((rock, dirt, water, total) and
// (mx, my, acc_x, acc_y); hg_particle_layout): the move pass reads (rock, dirt) of a texel with one 8-byte load and a
// droplet's whole footprint lies in ~11 DRAM sectors instead of 81 with one plane per channel (2.1 KB of DRAM traffic
// per droplet, profiles/r01i_all_kernels.txt); the erode pass issues one red.global.add.v4.f32 on the heightmap texel
// and one .v2 on the accumulator half of the momentum texel per corner instead of five scalar reductions.
#include "hg_internal.cuh"

namespace {

struct PDom { int W, H, pitch, row0; };      // global map size; first owned row of this context's images (ghost rows below it)
__device__ __forceinline__ size_t pidx(const PDom& d, int x, int y) { return (size_t)(y - d.row0 + HG_HALO_ROWS) * d.pitch + x; }

// Droplets on row slabs: every rank keeps the whole droplet array (the id is part of a
// droplet's identity: spawn hash) but owns -- moves and erodes -- only the droplets whose position lies in its rows;
// own[id] says which.  A droplet that respawns or drifts into another slab is handed over by writing its element and
// the two ownership bytes through the peer pointers; a corner texel in a neighbour's rows is eroded in the
// neighbour's image directly (atomics work over NVLink), so the texel stays ONE location with the lock's semantics.
struct PSlabs {
    int n, me;
    int row0[HG_MAX_SLABS], rows[HG_MAX_SLABS];
    float4* ha[HG_MAX_SLABS];            // read images of every slab (same set index on all ranks)
    float4* ma[HG_MAX_SLABS];
    hg_particle* parts[HG_MAX_SLABS];
    unsigned char* own[HG_MAX_SLABS];
};
__device__ __forceinline__ int slab_of_row(const PSlabs& S, int y) {
    int k = S.me;
    if (y < S.row0[k] || y >= S.row0[k] + S.rows[k]) {
        k = y < S.row0[0] ? 0 : S.n - 1;
        for (int j = 0; j < S.n; j++)
            if (y >= S.row0[j] && y < S.row0[j] + S.rows[j]) { k = j; break; }
    }
    return k;
}
// Store droplet `id` (state p) with the slab that holds its position.  Ownership byte: 0 = another slab's, 1 = mine,
// 2 = handed to me during an erode pass: mine from the NEXT dispatch on (its spawn kernel promotes 2 to 1).  Without the
// pending state a slab whose erode kernel starts after its neighbour's has finished would erode the droplets it has just
// been handed a second time (found by running the slab test under compute-sanitizer, which delays kernels).  Hand-overs
// of the spawn pass take effect at once: the receiving spawn kernel has nothing to do for a freshly spawned droplet,
// and the move pass runs after an all-rank generation.
__device__ __forceinline__ void slab_store_droplet(const PSlabs& S, uint32_t id, const hg_particle& p, int H, unsigned char marker) {
    int y = (int)p.position[1];
    y = min(max(y, 0), H - 1);
    const int o = slab_of_row(S, y);
    S.parts[o][id] = p;
    if (o != S.me) {
        __threadfence_system();          // the element before the ownership byte
        S.own[o][id] = marker;
        S.own[S.me][id] = 0;
    }
}
__device__ __forceinline__ bool poob(const PDom& d, int x, int y) { return x < 0 || x > d.W - 1 || y < 0 || y > d.H - 1; }
// texelFetch outside the image is 0 (hazard 3)
__device__ __forceinline__ float2 fetch_xy(const float4* __restrict__ img, const PDom& d, int x, int y) {
    return poob(d, x, y) ? make_float2(0.0f, 0.0f) : __ldg(reinterpret_cast<const float2*>(img + pidx(d, x, y)));
}
__device__ __forceinline__ float fetch_z(const float4* __restrict__ img, const PDom& d, int x, int y) {
    return poob(d, x, y) ? 0.0f : __ldg(reinterpret_cast<const float*>(img + pidx(d, x, y)) + 2);
}
// img_bilinear (img_interpolation.glsl:3-22) of the first two channels of a texel image at a float position
__device__ __forceinline__ float2 bilinear_xy(const float4* __restrict__ img, const PDom& d, float sx, float sy) {
    if (!(sx == sx)) sx = 0.0f;
    if (!(sy == sy)) sy = 0.0f;
    const int px = (int)sx, py = (int)sy;
    const float fx = hg_fract(sx), fy = hg_fract(sy);
    const float2 t00 = fetch_xy(img, d, px, py), t10 = fetch_xy(img, d, px + 1, py), t01 = fetch_xy(img, d, px, py + 1), t11 = fetch_xy(img, d, px + 1, py + 1);
    return make_float2(hg_bilerp(t00.x, t10.x, t01.x, t11.x, fx, fy), hg_bilerp(t00.y, t10.y, t01.y, t11.y, fx, fy));
}
__device__ __forceinline__ float bilinear_z(const float4* __restrict__ img, const PDom& d, float sx, float sy) {
    if (!(sx == sx)) sx = 0.0f;
    if (!(sy == sy)) sy = 0.0f;
    const int px = (int)sx, py = (int)sy;
    const float fx = hg_fract(sx), fy = hg_fract(sy);
    return hg_bilerp(fetch_z(img, d, px, py), fetch_z(img, d, px + 1, py), fetch_z(img, d, px, py + 1), fetch_z(img, d, px + 1, py + 1), fx, fy);
}

// rand(vec2), particle.glsl:41-44
__device__ __forceinline__ float prand(float px, float py) {
    return hg_fract(1e4f * hg_sinf(17.0f * px + py * 0.1f) * (0.1f + fabsf(hg_sinf(py * 13.0f + px))));
}

struct MoveArgs { const float4 *ha, *ma; hg_particle* particles; const uint32_t* order; const unsigned char* own; const uint32_t* n_valid; };

// The spawn half of particle.glsl:64-90 on its own, for slabs: a droplet that (re)spawns gets its hashed position
// anywhere on the map and must be with its new owner BEFORE it moves (the move samples the terrain around it).
__global__ void __launch_bounds__(128) k_particle_spawn(PSlabs S, int H, hg_erosion_data set, hg_map_settings_data map_set, uint32_t count, float time, int should_rain) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    unsigned char mine = S.own[S.me][id];
    if (mine == 2) { S.own[S.me][id] = 1; mine = 1; }                  // handed over during the last erode pass
    if (mine != 1 || !should_rain) return;                             // without rain the shader returns before it stores anything
    hg_particle p = S.parts[S.me][id];
    if (!(p.iters == 0 || p.to_kill)) return;
#pragma unroll
    for (int i = 0; i < HG_SED_LAYERS; i++)
        if (p.sediment[i] < 0.0f || p.iters == 0) p.sediment[i] = 0.0f;
    float posx = prand(hg_fract(time * 1.37f) * 1000.0f, (float)id) * (float)((float)map_set.hmap_dims[0] - 4.0f) / 1.0f + 2.0f;
    float posy = prand(hg_fract(time * 7.21f) * 1000.0f, (float)id + 3.14f) * (float)((float)map_set.hmap_dims[1] - 4.0f) / 1.0f + 2.0f;
    p.to_kill = 0;
    p.position[0] = posx; p.position[1] = posy;
    p.velocity[0] = 0.0f; p.velocity[1] = 0.0f;
    p.volume = set.init_volume;
    p.iters = 1;
    slab_store_droplet(S, id, p, H, 1);
}

// particle.glsl:64-136
__global__ void __launch_bounds__(128) k_particle_move(PDom d, HgStepParams P, hg_erosion_data set, hg_map_settings_data map_set,
                                                       MoveArgs A, uint32_t count, float time, int should_rain) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    if (A.order && id >= *A.n_valid) return;      // a slab's order holds only the droplets it owns
    if (A.order) id = A.order[id];      // thread k takes the k-th droplet in tile order; the droplet keeps its id (spawn hash)
    if (A.own && A.own[id] != 1) return;    // slabs: another rank's droplet
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
    // get_terr_normal, particle.glsl:50-62 (bilinear samples of rock and dirt)
    const float2 r_ = bilinear_xy(A.ha, d, px + 1.0f, py + 0.0f), l_ = bilinear_xy(A.ha, d, px + -1.0f, py + 0.0f);
    const float2 b_ = bilinear_xy(A.ha, d, px + 0.0f, py + -1.0f), t_ = bilinear_xy(A.ha, d, px + 0.0f, py + 1.0f);
    float dx = (r_.x + r_.y - l_.x - l_.y);
    float dz = (t_.x + t_.y - b_.x - b_.y);
    float nx = dx * 2.0f - 0.0f * dz, ny = 0.0f * 0.0f - 2.0f * 2.0f, nz = 2.0f * dz - dx * 0.0f;
    float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    nx *= inv; ny *= inv; nz *= inv;
    const float2 mom = bilinear_xy(A.ma, d, px, py);
    const float momx = mom.x, momy = mom.y;
    float water = bilinear_z(A.ha, d, px, py);
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

struct ErodeImages { float4 *ha, *ma; };

// What one corner of one droplet adds to its texel (particle_erosion.glsl:61-75,94-98): deposits per layer
// (only where the layer deposited), display water and the momentum accumulator.
struct CornerAdd { float rock, dirt, water, mz, mw; };

// (rock, dirt, water, -) += on the heightmap texel and (acc_x, acc_y) += on the momentum texel: two vector reductions.
// A layer without a deposit adds +0.0f, which leaves every value as it is.
__device__ __forceinline__ void texel_add(float4* hp, float4* mp, float rock, float dirt, float water, float mz, float mw) {
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(hp), "f"(rock), "f"(dirt), "f"(water), "f"(0.0f) : "memory");
    asm volatile("red.global.v2.f32.add [%0], {%1, %2};" ::"l"(reinterpret_cast<float*>(mp) + 2), "f"(mz), "f"(mw) : "memory");
}

// Apply the corner additions of all 32 lanes.  Lanes of the warp that hit the same texel are found with ONE
// match per corner; the lowest such lane sums its peers' five values in lane order and issues the two reductions.
// All 32 lanes must call.
__device__ __forceinline__ void warp_corner_add(float4* hp, float4* mp, const CornerAdd& v, bool act) {
    const int lane = threadIdx.x & 31;
    unsigned long long key = act ? (unsigned long long)hp : (unsigned long long)lane;      // a texel address is never < 32
    unsigned peers = __match_any_sync(0xffffffffu, key);
    if (!act) return;
    if (__popc(peers) == 1) {              // the common case on a sparse map: nobody else in the warp hits this texel
        texel_add(hp, mp, v.rock, v.dirt, v.water, v.mz, v.mw);
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
    if (lane == __ffs(peers) - 1) texel_add(hp, mp, s_rock, s_dirt, s_water, s_mz, s_mw);
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

// erode_clamped for all 32 lanes at once (all must call; `on`: the lane erodes).  Lanes of the warp that erode the same
// layer of the same texel -- the normal case once the droplets are processed in tile order -- are applied in lane order
// by ONE compare-and-swap of their leader instead of up to 32 competing ones: every lane replays the sequence
// (running = max(running - eroded_j, 0), what erode_clamped stores) from the value the leader read, which gives each
// its own value before and unclamped value after, exactly as if the lanes had taken the reference's lock one after
// the other; the leader publishes the final value and the group repeats if another warp got in between.
__device__ __forceinline__ void warp_erode(float* addr, float eroded, bool on, float* old_terr, float* after) {
    const int lane = threadIdx.x & 31;
    const unsigned long long key = on ? (unsigned long long)addr : (unsigned long long)lane;      // an address is never < 32
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (!on) return;
    if (__popc(peers) == 1) {
        *old_terr = erode_clamped(addr, eroded, after);
        return;
    }
    const int leader = __ffs(peers) - 1;
    unsigned* const ua = reinterpret_cast<unsigned*>(addr);
    unsigned cur = 0;
    if (lane == leader) cur = *reinterpret_cast<volatile unsigned*>(ua);
    for (;;) {
        cur = __shfl_sync(peers, cur, leader);
        float running = __uint_as_float(cur), mine_old = 0.0f, mine_after = 0.0f;
        unsigned rest = peers;
        while (rest) {
            const int src = __ffs(rest) - 1;
            rest &= rest - 1;
            const float e = __shfl_sync(peers, eroded, src);
            const float nv = running - e;
            if (src == lane) { mine_old = running; mine_after = nv; }
            running = (nv < 0.0f) ? 0.0f : nv;
        }
        unsigned seen = cur;
        if (lane == leader) seen = atomicCAS(ua, cur, __float_as_uint(running));
        seen = __shfl_sync(peers, seen, leader);
        if (seen == cur) { *old_terr = mine_old; *after = mine_after; return; }
        cur = seen;
    }
}

struct ErodeArgs : ErodeImages { hg_particle* particles; const uint32_t* order; const uint32_t* n_valid; };

// the heightmap / momentum texel (x, y) wherever it lives: this slab's images, or the owning rank's through its peer pointer
template <bool SLAB>
__device__ __forceinline__ void texel_ptrs(const ErodeArgs& A, const PSlabs& S, const PDom& d, int x, int y, float4** hp, float4** mp) {
    if (!SLAB) { const size_t ti = pidx(d, x, y); *hp = A.ha + ti; *mp = A.ma + ti; return; }
    const int k = slab_of_row(S, y);
    const size_t ti = (size_t)(y - S.row0[k] + HG_HALO_ROWS) * d.pitch + x;
    *hp = S.ha[k] + ti; *mp = S.ma[k] + ti;
}

// particle_erosion.glsl:101-128 + erode_layers :22-85
template <bool SLAB>
__global__ void __launch_bounds__(128) k_particle_erode(PDom d, HgStepParams P, ErodeArgs A, uint32_t count, const __grid_constant__ PSlabs S) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = id < count;
    if (live && A.order) live = id < *A.n_valid;
    if (live && A.order) id = A.order[id];
    if (SLAB) live = live && S.own[S.me][id] == 1;      // not 2: a droplet handed over during this very pass has been eroded by its old owner
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
        float4 *hp = nullptr, *mp = nullptr;
        if (act) texel_ptrs<SLAB>(A, S, d, cx, cy, &hp, &mp);
        float multipl = wx * wy;
        float dep[2] = {0.0f, 0.0f};      // value-independent additions to terrain, per layer
        // erode_layers (particle_erosion.glsl:22-85) layer by layer with every lane of the warp in step, because the
        // erosion of a layer is a warp-collective update (warp_erode); `open`: the lane is still inside the shader's loop
        float cap = 0.0f;
        bool open = act;
#pragma unroll
        for (int i = HG_SED_LAYERS - 1; i >= 0; i--) {
            const bool kill = open && part.to_kill;
            if (kill) {
                float sed = old_sed[i] * multipl;
                dep[i] = sed;
                part.sediment[i] -= sed;
            }
            const float c = hg_max(0.0f, part.sc - cap);
            float s1 = old_sed[i];
            const bool erode = open && !kill && c > s1;
            const float eroded = erode ? multipl * P.Kls[i] * (c - s1) : 0.0f;
            float old_terr = 0.0f, after = 0.0f;
            warp_erode(reinterpret_cast<float*>(hp) + i, eroded, erode, &old_terr, &after);      // .x rock, .y dirt
            if (erode) {
                s1 += eroded;
                if (after < 0.0f) {
                    s1 += after;
                    cap += old_terr;
                    part.sediment[i] = s1;
                } else {
                    part.sediment[i] = s1;
                    open = false;      // break
                }
            } else if (open && !kill) {
                float deposit = multipl * P.Kld[i] * (s1 - c);
                s1 -= deposit;
                dep[i] = deposit;
                part.sediment[i] = s1;
            }
        }
        if (act) {
            float conv = part.sediment[0] * P.Kconv * P.d_t;
            part.sediment[1] += conv;
            part.sediment[0] -= conv;
        }
        CornerAdd v;
        v.rock = dep[0]; v.dirt = dep[1];
        v.water = 1e-5f * part.volume * multipl;
        v.mz = part.volume * part.velocity[0] * multipl;
        v.mw = part.volume * part.velocity[1] * multipl;
        warp_corner_add(hp, mp, v, act);
    }
    if (live) {
        if (SLAB) slab_store_droplet(S, id, part, d.H, 2);      // stays, or goes to the slab it drifted into (from the next dispatch on)
        else A.particles[id] = part;
    }
}


// ------------------------------------------------------------------ processing order
// Droplets are spawned at hashed positions, so droplet k and droplet k+1 sit in unrelated parts of the map and every
// thread of a warp gathers from and scatters to its own DRAM sectors (66 % of the DRAM peak at 7 % issue utilisation,
// profiles/r01i_all_kernels.txt).  The order in which droplets are processed is free (the reference leaves it to the
// GPU's scheduling and to its lock), so both kernels take their droplet from an index array sorted by map tile: a
// counting sort over the tiles of the droplets' positions (tile = 2^shift cells, at most 2^21 bins: single cells
// when hmap_dims is small, so that the droplets of one texel sit next to each other and the warp-level combining of
// the erode pass meets them).  Droplets move at most 0.25 cell per step and about 1 % respawn per step, so the order
// is rebuilt only every p_rebin_period dispatches.
constexpr int kScanBlock = 2048;      // bins per scan block (512 threads x 4)
struct BinDom { int shift, bx, by; };
// own: null, or the slab's ownership bytes: another rank's droplets are left out of the order
__global__ void __launch_bounds__(256) k_bin_keys(const hg_particle* parts, uint32_t count, BinDom b, uint32_t* keys, uint32_t* hist, const unsigned char* own) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    if (own && own[id] != 1) { keys[id] = 0xffffffffu; return; }      // not in the order at all
    const float2 pos = *reinterpret_cast<const float2*>(parts[id].position);
    int cx = (int)fminf(fmaxf(pos.x, 0.0f), 1e9f) >> b.shift, cy = (int)fminf(fmaxf(pos.y, 0.0f), 1e9f) >> b.shift;
    cx = min(cx, b.bx - 1); cy = min(cy, b.by - 1);
    const uint32_t key = (uint32_t)cy * (uint32_t)b.bx + (uint32_t)cx;
    keys[id] = key;
    atomicAdd(hist + key, 1u);
}
// exclusive scan of every block of kScanBlock bins in place; block totals to sums[]
__global__ void __launch_bounds__(512) k_bin_scan1(uint32_t* hist, int nbins, uint32_t* sums) {
    __shared__ uint32_t wsum[16];
    const int base = blockIdx.x * kScanBlock + threadIdx.x * 4;
    uint32_t v[4], t = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = base + k < nbins ? hist[base + k] : 0u; t += v[k]; }
    uint32_t incl = t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < 16 ? wsum[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += n; }
        if (lane < 16) wsum[lane] = wi - w;
        if (lane == 15) sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = wsum[warp] + incl - t;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < nbins) hist[base + k] = run; run += v[k]; }
}
// exclusive scan of the block totals (at most 1024), one block
__global__ void __launch_bounds__(1024) k_bin_scan2(uint32_t* sums, int n) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v = (int)threadIdx.x < n ? sums[threadIdx.x] : 0u, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t m = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += m; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t m = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += m; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    if ((int)threadIdx.x < n) sums[threadIdx.x] = wsum[warp] + incl - v;
}
// last_bin: an empty bin behind all the others; its offset after the scan is the number of droplets in the order
__global__ void __launch_bounds__(256) k_bin_scatter(const uint32_t* keys, uint32_t count, uint32_t* hist, const uint32_t* sums, uint32_t* order,
                                                     uint32_t last_bin, uint32_t* n_valid) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id == 0) *n_valid = hist[last_bin] + sums[last_bin / kScanBlock];
    if (id >= count) return;
    const uint32_t key = keys[id];
    if (key == 0xffffffffu) return;
    const uint32_t pos = atomicAdd(hist + key, 1u) + sums[key / kScanBlock];
    order[pos] = id;
}

}  // namespace

// scratch of the processing order; a droplet SLAB allocates it for the largest bin count at create time (no allocation
// may happen between two exchange generations)
int hg_particle_order_alloc(hg_ctx* c, int nbins) {
    if (!c->p_order) {
        HG_CUDA(cudaMalloc(&c->p_order, (size_t)c->particle_count * sizeof(uint32_t)));
        HG_CUDA(cudaMalloc(&c->p_keys, (size_t)c->particle_count * sizeof(uint32_t)));
    }
    if (c->p_bins_cap < nbins) {
        if (c->p_hist) HG_CUDA(cudaFree(c->p_hist));
        HG_CUDA(cudaMalloc(&c->p_hist, ((size_t)nbins + 1100) * sizeof(uint32_t)));
        c->p_bins_cap = nbins;
    }
    return HG_OK;
}

// (Re)build the processing order from the droplets' current positions.
static int rebuild_order(hg_ctx* c, uint32_t count) {
    int hx = c->map.hmap_dims[0] > 0 ? c->map.hmap_dims[0] : c->g.W, hy = c->map.hmap_dims[1] > 0 ? c->map.hmap_dims[1] : c->g.H;
    if (hx > c->g.W) hx = c->g.W;
    if (hy > c->g.H) hy = c->g.H;
    BinDom b{0, hx, hy};
    while ((long long)b.bx * b.by > (1LL << 21)) { b.shift++; b.bx = (hx + (1 << b.shift) - 1) >> b.shift; b.by = (hy + (1 << b.shift) - 1) >> b.shift; }
    const int nbins = b.bx * b.by + 1;      // + an empty last bin whose offset is the number of droplets in the order (sums[1050])
    const int nblocks = (nbins + kScanBlock - 1) / kScanBlock;
    int rca = hg_particle_order_alloc(c, nbins);
    if (rca) return rca;
    uint32_t* sums = c->p_hist + c->p_bins_cap;
    HG_CUDA(cudaMemsetAsync(c->p_hist, 0, (size_t)nbins * sizeof(uint32_t), c->stream));
    k_bin_keys<<<(count + 255) / 256, 256, 0, c->stream>>>(c->particles, count, b, c->p_keys, c->p_hist, (c->p_own && c->peers_connected) ? c->p_own : nullptr);
    HG_LAUNCH_CHECK(c);
    k_bin_scan1<<<nblocks, 512, 0, c->stream>>>(c->p_hist, nbins, sums);
    HG_LAUNCH_CHECK(c);
    k_bin_scan2<<<1, 1024, 0, c->stream>>>(sums, nblocks);
    HG_LAUNCH_CHECK(c);
    k_bin_scatter<<<(count + 255) / 256, 256, 0, c->stream>>>(c->p_keys, count, c->p_hist, sums, c->p_order, (uint32_t)(nbins - 1), sums + 1050);
    HG_LAUNCH_CHECK(c);
    c->p_order_valid = true;
    c->p_rebin_age = 0;
    return HG_OK;
}

// the droplet view of the slab table: every rank's read images, droplet array and ownership bytes
static PSlabs make_pslabs(hg_ctx* c) {
    PSlabs S;
    memset(&S, 0, sizeof(S));
    S.n = c->slabs.n; S.me = c->slabs.me;
    for (int k = 0; k < S.n; k++) {
        S.row0[k] = c->slabs.row0[k]; S.rows[k] = c->slabs.rows[k];
        const size_t pe = (size_t)(c->slabs.rows[k] + 2 * HG_HALO_ROWS) * c->g.pitch;      // texels per image of slab k
        S.ha[k] = c->peer_pa[k] + (size_t)c->ri[0] * pe;              // every rank flips its read indices in step
        S.ma[k] = c->peer_pa[k] + (size_t)(2 + c->ri[2]) * pe;
        S.parts[k] = c->peer_parts[k];
        S.own[k] = c->peer_own[k];
    }
    return S;
}
static bool droplet_slabs(const hg_ctx* c) { return c->erosion_type == HG_PARTICLES && c->peers_connected; }

// slabs only: respawn + hand-over of the droplets this rank owns (the first half of particle.glsl's main)
int hg_launch_particle_spawn(hg_ctx* c, float time, int should_rain) {
    uint32_t count = (c->particle_count / 64u) * 64u;
    if (!count || !droplet_slabs(c)) return HG_OK;
    k_particle_spawn<<<(count + 127) / 128, 128, 0, c->stream>>>(make_pslabs(c), c->g.H, c->erosion, c->map, count, time, should_rain);
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}

int hg_launch_particle_move(hg_ctx* c, float time, int should_rain) {
    PDom d{c->g.W, c->g.H, c->g.pitch, c->g.row0};
    uint32_t count = (c->particle_count / 64u) * 64u;    // glDispatchCompute(particle_count/64), erosion.cpp:127
    if (!count) return HG_OK;
    int rc = hg_particle_layout(c, true);
    if (rc) return rc;
    const bool slabs = droplet_slabs(c);
    // processing order: rebuilt when it has aged; a fresh context (nothing spawned yet) moves in id order and sorts
    // after the move, when the droplets have their spawn positions (hg_launch_particle_erode).  A slab's set of
    // droplets changes every step (hand-overs), so it sorts the droplets it owns before every move; the others
    // land in a bin of their own at the end and are skipped by their ownership byte.
    if (c->p_rebin_period > 0 && (slabs || (c->p_order_valid && ++c->p_rebin_age >= c->p_rebin_period))) {
        rc = rebuild_order(c, count);
        if (rc) return rc;
    }
    MoveArgs A{hg_pa_h(c, 1), hg_pa_m(c, 1), c->particles, c->p_order_valid ? c->p_order : nullptr, slabs ? c->p_own : nullptr, c->p_order_valid ? c->p_hist + c->p_bins_cap + 1050 : nullptr};
    k_particle_move<<<(count + 127) / 128, 128, 0, c->stream>>>(d, c->sp, c->erosion, c->map, A, count, time, should_rain);
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}

int hg_launch_particle_erode(hg_ctx* c) {
    PDom d{c->g.W, c->g.H, c->g.pitch, c->g.row0};
    uint32_t count = (c->particle_count / 64u) * 64u;
    if (!count) return HG_OK;
    // in place on the READ images of heightmap and momentum map (erosion.cpp:141-143)
    int rc = hg_particle_layout(c, true);
    if (rc) return rc;
    const bool slabs = droplet_slabs(c);
    if (!slabs && c->p_rebin_period > 0 && !c->p_order_valid) {
        rc = rebuild_order(c, count);
        if (rc) return rc;
    }
    ErodeArgs A;
    A.ha = hg_pa_h(c, 1); A.ma = hg_pa_m(c, 1); A.particles = c->particles; A.order = c->p_order_valid ? c->p_order : nullptr; A.n_valid = c->p_order_valid ? c->p_hist + c->p_bins_cap + 1050 : nullptr;
    if (slabs) {
        k_particle_erode<true><<<(count + 127) / 128, 128, 0, c->stream>>>(d, c->sp, A, count, make_pslabs(c));
    } else {
        PSlabs none;
        memset(&none, 0, sizeof(none));
        k_particle_erode<false><<<(count + 127) / 128, 128, 0, c->stream>>>(d, c->sp, A, count, none);
    }
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}

// ownership at connect time: unspawned droplets are dealt out round robin (their first position is a hash of the id)
__global__ void k_own_init(unsigned char* own, uint32_t count, int n, int me) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id < count) own[id] = (int)(id % (uint32_t)n) == me ? 1 : 0;
}
// ownership after the droplet array was replaced from outside (hg_upload_particles on every slab, checkpoint load):
// a spawned droplet belongs to the slab that holds its row, an unspawned one is dealt out as at connect time
__global__ void k_own_from_positions(PSlabs S, int H, uint32_t count) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    const hg_particle p = S.parts[S.me][id];
    int o = (int)(id % (uint32_t)S.n);
    if (p.iters != 0) o = slab_of_row(S, min(max((int)p.position[1], 0), H - 1));
    S.own[S.me][id] = o == S.me ? 1 : 0;
}
int hg_particle_own_init(hg_ctx* c) {
    if (!c->p_own || !c->particle_count) return HG_OK;
    if (c->peers_connected && c->peer_parts[c->slabs.me]) k_own_from_positions<<<(c->particle_count + 255) / 256, 256, 0, c->stream>>>(make_pslabs(c), c->g.H, c->particle_count);
    else k_own_init<<<(c->particle_count + 255) / 256, 256, 0, c->stream>>>(c->p_own, c->particle_count, c->slabs.n > 0 ? c->slabs.n : 1, c->slabs.me);
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}

int hg_preload_particle_kernels(void) {
    cudaFuncAttributes a;
    HG_CUDA(cudaFuncGetAttributes(&a, k_particle_spawn));
    HG_CUDA(cudaFuncGetAttributes(&a, k_particle_move));
    HG_CUDA(cudaFuncGetAttributes(&a, k_particle_erode<true>));
    HG_CUDA(cudaFuncGetAttributes(&a, k_particle_erode<false>));
    HG_CUDA(cudaFuncGetAttributes(&a, k_own_init));
    HG_CUDA(cudaFuncGetAttributes(&a, k_own_from_positions));
    HG_CUDA(cudaFuncGetAttributes(&a, k_bin_keys));
    HG_CUDA(cudaFuncGetAttributes(&a, k_bin_scan1));
    HG_CUDA(cudaFuncGetAttributes(&a, k_bin_scan2));
    HG_CUDA(cudaFuncGetAttributes(&a, k_bin_scatter));
    return HG_OK;
}
