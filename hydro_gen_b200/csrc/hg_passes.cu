// hg_passes.cu — the PASSES schedule: the reference's grid dispatches, one kernel each
// (src/erosion.cpp:158-200), on SoA planes.  This is the validation path: it
// materialises H.a, V, TC and TD exactly as the reference's textures hold them so
// every intermediate can be downloaded and compared.  The arithmetic is the shared
// per-cell code of hg_cell.cuh; the product path is the fused kernel (hg_fused.cu).
#include "hg_internal.cuh"

namespace {

struct Dom { int W, H, pitch; };   // passes run on the full map: local row = global row + HG_HALO_ROWS

__device__ __forceinline__ size_t cidx(const Dom& d, int x, int y) { return (size_t)(y + HG_HALO_ROWS) * d.pitch + x; }
__device__ __forceinline__ bool oob(const Dom& d, int x, int y) { return x < 0 || x > d.W - 1 || y < 0 || y > d.H - 1; }
__device__ __forceinline__ float ldv(const float* __restrict__ p, const Dom& d, int x, int y, float oobv) {
    return oob(d, x, y) ? oobv : __ldg(p + cidx(d, x, y));
}

struct FluxArgs {
    const float *rock, *dirt, *water, *total, *fL, *fR, *fT, *fB, *vw;
    float *o_rock, *o_dirt, *o_water, *o_total, *o_fL, *o_fR, *o_fT, *o_fB, *o_u, *o_v, *o_vz, *o_vw;
};
// hydro_flux.glsl:77-166
__global__ void __launch_bounds__(256) k_flux(Dom d, HgStepParams P, FluxArgs A) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    size_t i = cidx(d, x, y);
    float rock = A.rock[i], dirt = A.dirt[i], water = A.water[i], a = A.total[i];
    HgFluxOut o = hg_flux_cell(P, x, y, d.W, d.H, a,
        ldv(A.total, d, x - 1, y, HG_OOB_HEIGHT), ldv(A.total, d, x + 1, y, HG_OOB_HEIGHT),
        ldv(A.total, d, x, y + 1, HG_OOB_HEIGHT), ldv(A.total, d, x, y - 1, HG_OOB_HEIGHT),
        A.fL[i], A.fR[i], A.fT[i], A.fB[i],
        ldv(A.fR, d, x - 1, y, 0.0f), ldv(A.fL, d, x + 1, y, 0.0f),
        ldv(A.fB, d, x, y + 1, 0.0f), ldv(A.fT, d, x, y - 1, 0.0f), water);
    A.o_fL[i] = o.fL; A.o_fR[i] = o.fR; A.o_fT[i] = o.fT; A.o_fB[i] = o.fB;
    A.o_rock[i] = rock; A.o_dirt[i] = dirt; A.o_water[i] = o.water;
    A.o_total[i] = rock + o.water + dirt;            // hydro_flux.glsl:137 (this order)
    A.o_u[i] = o.u; A.o_v[i] = o.v; A.o_vz[i] = o.vz; A.o_vw[i] = A.vw[i];
}

struct EroArgs {
    const float *rock, *dirt, *water, *sr, *sd, *u, *v, *vz;
    float *o_rock, *o_dirt, *o_water, *o_total, *o_sr, *o_sd;
};
// hydro_erosion.glsl:37-92
__global__ void __launch_bounds__(256) k_erosion(Dom d, HgStepParams P, EroArgs A) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    size_t i = cidx(d, x, y);
    float water = A.water[i];
    HgEroOut o = hg_erosion_cell(P, A.rock[i], A.dirt[i], A.sr[i], A.sd[i], A.u[i], A.v[i], A.vz[i],
        ldv(A.rock, d, x + 1, y, 0.0f), ldv(A.dirt, d, x + 1, y, 0.0f),
        ldv(A.rock, d, x - 1, y, 0.0f), ldv(A.dirt, d, x - 1, y, 0.0f),
        ldv(A.rock, d, x, y - 1, 0.0f), ldv(A.dirt, d, x, y - 1, 0.0f),
        ldv(A.rock, d, x, y + 1, 0.0f), ldv(A.dirt, d, x, y + 1, 0.0f));
    A.o_rock[i] = o.rock; A.o_dirt[i] = o.dirt; A.o_water[i] = water;
    A.o_total[i] = o.rock + o.dirt + water;
    A.o_sr[i] = o.sr; A.o_sd[i] = o.sd;
}

struct SedArgs {
    const float *rock, *dirt, *water, *sr, *sd, *u, *v;
    float *o_rock, *o_dirt, *o_water, *o_total, *o_sr, *o_sd;
};
// sediment_transport.glsl:66-93
__global__ void __launch_bounds__(256) k_sediment(Dom d, HgStepParams P, SedArgs A) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    size_t i = cidx(d, x, y);
    HgBack b = hg_backtrace(P, x, y, d.W, d.H, A.u[i], A.v[i]);
    float sr = hg_bilerp(ldv(A.sr, d, b.px, b.py, 0.0f), ldv(A.sr, d, b.px + 1, b.py, 0.0f),
                         ldv(A.sr, d, b.px, b.py + 1, 0.0f), ldv(A.sr, d, b.px + 1, b.py + 1, 0.0f), b.sx, b.sy);
    float sd = hg_bilerp(ldv(A.sd, d, b.px, b.py, 0.0f), ldv(A.sd, d, b.px + 1, b.py, 0.0f),
                         ldv(A.sd, d, b.px, b.py + 1, 0.0f), ldv(A.sd, d, b.px + 1, b.py + 1, 0.0f), b.sx, b.sy);
    float rock = A.rock[i], dirt = A.dirt[i], water = A.water[i];
    water *= P.evap;
    A.o_rock[i] = rock; A.o_dirt[i] = dirt; A.o_water[i] = water;
    A.o_total[i] = rock + dirt + water;
    A.o_sr[i] = sr; A.o_sd[i] = sd;
}

struct ThFluxArgs { const float *rock, *dirt; float* tc[4]; float* td[4]; };
// thermal_erosion.glsl:28-115
__global__ void __launch_bounds__(256) k_thermal_flux(Dom d, HgStepParams P, ThFluxArgs A, int layer) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    size_t i = cidx(d, x, y);
    const int ox[8] = {-1, 1, 0, 0, -1, 1, -1, 1};
    const int oy[8] = {0, 0, 1, -1, 1, 1, -1, -1};
    float rock = A.rock[i], dirt = A.dirt[i];
    float d_h[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float dh = 0.0f;
        dh += rock - ldv(A.rock, d, x + ox[k], y + oy[k], HG_OOB_HEIGHT);
        if (layer >= 1) dh += dirt - ldv(A.dirt, d, x + ox[k], y + oy[k], HG_OOB_HEIGHT);
        d_h[k] = dh;
    }
    float out[8];
    hg_thermal_outflow(P, layer, layer == 0 ? rock : dirt, d_h, out);
#pragma unroll
    for (int k = 0; k < 4; k++) { A.tc[k][i] = out[k]; A.td[k][i] = out[4 + k]; }
}

struct ThTransArgs {
    const float *rock, *dirt, *water; const float* tc[4]; const float* td[4];
    float *o_rock, *o_dirt, *o_water, *o_total;
};
// thermal_transport.glsl:31-65
__global__ void __launch_bounds__(256) k_thermal_transport(Dom d, ThTransArgs A, int layer) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    size_t i = cidx(d, x, y);
    float neg = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) neg -= A.tc[k][i];
#pragma unroll
    for (int k = 0; k < 4; k++) neg -= A.td[k][i];
    // c = (L,R,T,B), d = (LT,RT,LB,RB)
    float delta = hg_thermal_delta(neg,
        ldv(A.tc[1], d, x - 1, y, 0.0f), ldv(A.tc[0], d, x + 1, y, 0.0f),
        ldv(A.tc[3], d, x, y + 1, 0.0f), ldv(A.tc[2], d, x, y - 1, 0.0f),
        ldv(A.td[3], d, x - 1, y + 1, 0.0f), ldv(A.td[2], d, x + 1, y + 1, 0.0f),
        ldv(A.td[1], d, x - 1, y - 1, 0.0f), ldv(A.td[0], d, x + 1, y - 1, 0.0f));
    float rock = A.rock[i], dirt = A.dirt[i], water = A.water[i];
    if (layer == 0) rock += delta; else dirt += delta;
    A.o_rock[i] = rock; A.o_dirt[i] = dirt; A.o_water[i] = water;
    A.o_total[i] = rock + dirt + water;
}

struct SmoothArgs {
    const float *rock, *dirt, *water, *total; const float* m[4];
    float *o_rock, *o_dirt, *o_water, *o_total; float* om[4];
};
// smoothing.glsl:22-103
__global__ void __launch_bounds__(256) k_smooth(Dom d, HgStepParams P, SmoothArgs A, int momentum) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    size_t i = cidx(d, x, y);
    float rock = A.rock[i], dirt = A.dirt[i], water = A.water[i];
    if (x == 0 || y == 0 || x == d.W - 1 || y == d.H - 1) {
        A.o_rock[i] = rock; A.o_dirt[i] = dirt; A.o_water[i] = water; A.o_total[i] = A.total[i];
        if (momentum) { A.om[0][i] = 0.0f; A.om[1][i] = 0.0f; A.om[2][i] = 0.0f; A.om[3][i] = 0.0f; }
        return;
    }
    size_t l = i - 1, r = i + 1, t = i + d.pitch, b = i - d.pitch;
    hg_smooth_cell(P, rock, dirt, A.rock[l], A.dirt[l], A.rock[r], A.dirt[r], A.rock[t], A.dirt[t], A.rock[b], A.dirt[b]);
    if (P.particle_count != 0 && momentum) {
        float mx = A.m[0][i], my = A.m[1][i], mz = A.m[2][i], mw = A.m[3][i];
        hg_smooth_momentum(P, mx, my, mz, mw, water);
        A.om[0][i] = mx; A.om[1][i] = my; A.om[2][i] = mz; A.om[3][i] = mw;
    }
    A.o_rock[i] = rock; A.o_dirt[i] = dirt; A.o_water[i] = water;
    A.o_total[i] = rock + dirt + water;
}

// H.a = (rock + dirt) + water: what the last writer of a completed step (smooth, rain,
// init) leaves in the alpha channel; used when entering the PASSES schedule.
__global__ void __launch_bounds__(256) k_fill_total(Dom d, const float* rock, const float* dirt, const float* water, float* total) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    size_t i = cidx(d, x, y);
    total[i] = rock[i] + dirt[i] + water[i];
}

inline dim3 grid_for(const hg_ctx* c, dim3 b) { return dim3((c->g.W + b.x - 1) / b.x, (c->g.H + b.y - 1) / b.y); }

}  // namespace

int hg_fill_total(hg_ctx* c) {
    Dom d{c->g.W, c->g.H, c->g.pitch};
    dim3 b(32, 8);
    k_fill_total<<<grid_for(c, b), b, 0, c->stream>>>(d, hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_total(c, 1));
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}

static int require_full_map(hg_ctx* c) {
    if (c->g.row0 != 0 || c->g.rows != c->g.H) {
        hg_set_error("the PASSES schedule runs on a whole map only (this context is a slab)");
        return HG_ERR_STATE;
    }
    return hg_ensure_aux(c);
}

int hg_launch_pass(hg_ctx* c, int pass) {
    int rc = require_full_map(c);
    if (rc) return rc;
    rc = hg_particle_layout(c, false);      // the 1:1 pass kernels work on the planes
    if (rc) return rc;
    Dom d{c->g.W, c->g.H, c->g.pitch};
    dim3 b(32, 8), g = grid_for(c, b);
    switch (pass) {
    case HG_PASS_FLUX: {
        FluxArgs A{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_total(c, 1),
                   hg_cur(c, PL_FL, 1), hg_cur(c, PL_FR, 1), hg_cur(c, PL_FT, 1), hg_cur(c, PL_FB, 1), hg_vel(c, 3, 1),
                   hg_cur(c, PL_ROCK, 0), hg_cur(c, PL_DIRT, 0), hg_cur(c, PL_WATER, 0), hg_total(c, 0),
                   hg_cur(c, PL_FL, 0), hg_cur(c, PL_FR, 0), hg_cur(c, PL_FT, 0), hg_cur(c, PL_FB, 0),
                   hg_vel(c, 0, 0), hg_vel(c, 1, 0), hg_vel(c, 2, 0), hg_vel(c, 3, 0)};
        k_flux<<<g, b, 0, c->stream>>>(d, c->sp, A);
        HG_LAUNCH_CHECK(c);
        c->ri[0] ^= 1; c->ri[1] ^= 1; c->ri[2] ^= 1;
        break;
    }
    case HG_PASS_EROSION: {
        EroArgs A{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_cur(c, PL_SR, 1), hg_cur(c, PL_SD, 1),
                  hg_vel(c, 0, 1), hg_vel(c, 1, 1), hg_vel(c, 2, 1),
                  hg_cur(c, PL_ROCK, 0), hg_cur(c, PL_DIRT, 0), hg_cur(c, PL_WATER, 0), hg_total(c, 0),
                  hg_cur(c, PL_SR, 0), hg_cur(c, PL_SD, 0)};
        k_erosion<<<g, b, 0, c->stream>>>(d, c->sp, A);
        HG_LAUNCH_CHECK(c);
        c->ri[0] ^= 1; c->ri[3] ^= 1;
        break;
    }
    case HG_PASS_SEDIMENT: {
        SedArgs A{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_cur(c, PL_SR, 1), hg_cur(c, PL_SD, 1),
                  hg_vel(c, 0, 1), hg_vel(c, 1, 1),
                  hg_cur(c, PL_ROCK, 0), hg_cur(c, PL_DIRT, 0), hg_cur(c, PL_WATER, 0), hg_total(c, 0),
                  hg_cur(c, PL_SR, 0), hg_cur(c, PL_SD, 0)};
        k_sediment<<<g, b, 0, c->stream>>>(d, c->sp, A);
        HG_LAUNCH_CHECK(c);
        c->ri[0] ^= 1; c->ri[3] ^= 1;
        break;
    }
    case HG_PASS_THERMAL: {
        // run_thermal_erosion, src/erosion.cpp:103-121
        for (int layer = 0; layer < HG_SED_LAYERS; layer++) {
            ThFluxArgs F{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), {}, {}};
            for (int k = 0; k < 4; k++) { F.tc[k] = hg_aux_plane(c, AX_TC + k); F.td[k] = hg_aux_plane(c, AX_TD + k); }
            k_thermal_flux<<<g, b, 0, c->stream>>>(d, c->sp, F, layer);
            HG_LAUNCH_CHECK(c);
            ThTransArgs T{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), {}, {},
                          hg_cur(c, PL_ROCK, 0), hg_cur(c, PL_DIRT, 0), hg_cur(c, PL_WATER, 0), hg_total(c, 0)};
            for (int k = 0; k < 4; k++) { T.tc[k] = hg_aux_plane(c, AX_TC + k); T.td[k] = hg_aux_plane(c, AX_TD + k); }
            k_thermal_transport<<<g, b, 0, c->stream>>>(d, T, layer);
            HG_LAUNCH_CHECK(c);
            c->ri[0] ^= 1;
        }
        break;
    }
    case HG_PASS_SMOOTH: {
        int momentum = (c->erosion_type == HG_PARTICLES);
        SmoothArgs A{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_total(c, 1), {},
                     hg_cur(c, PL_ROCK, 0), hg_cur(c, PL_DIRT, 0), hg_cur(c, PL_WATER, 0), hg_total(c, 0), {}};
        for (int k = 0; k < 4; k++) { A.m[k] = hg_vel(c, k, 1); A.om[k] = hg_vel(c, k, 0); }
        k_smooth<<<g, b, 0, c->stream>>>(d, c->sp, A, momentum);
        HG_LAUNCH_CHECK(c);
        c->ri[0] ^= 1;
        if (momentum) c->ri[2] ^= 1;
        break;
    }
    default:
        hg_set_error("unknown pass %d", pass);
        return HG_ERR_INVALID;
    }
    return HG_OK;
}

// Erosion::dispatch_grid, src/erosion.cpp:158-200
int hg_launch_passes_step(hg_ctx* c) {
    for (int p = HG_PASS_FLUX; p <= HG_PASS_SMOOTH; p++) {
        int rc = hg_launch_pass(c, p);
        if (rc) return rc;
    }
    return HG_OK;
}

// thermal x2 + smooth with the momentum map bound: the tail of Erosion::dispatch_particle
// (src/erosion.cpp:146-155)
int hg_launch_thermal_smooth_particle(hg_ctx* c) {
    int rc = hg_launch_pass(c, HG_PASS_THERMAL);
    if (rc) return rc;
    return hg_launch_pass(c, HG_PASS_SMOOTH);
}
