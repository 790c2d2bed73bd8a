// hg_init_rain.cu — heightmap initialisation (State::World::gen_heightmap,
// src/state.cpp:116-147 -> glsl/heightmap.glsl) and rain (Erosion::dispatch_grid_rain,
// src/erosion.cpp:76-89 -> glsl/rain.glsl).  Both are pointwise in the GLOBAL cell
// coordinate and ALU-bound (32 gradient-noise / 8 simplex evaluations per cell), so a
// slab simply evaluates its own rows plus its ghost rows: no exchange is needed.
#include "hg_internal.cuh"
#include "hg_noise.cuh"

namespace {

struct SlabDom { int W, H, pitch, row0, rows; };

// local row index (ghost rows first) of global row gy
__device__ __forceinline__ size_t sidx(const SlabDom& d, int x, int gy) {
    return (size_t)(gy - d.row0 + HG_HALO_ROWS) * d.pitch + x;
}

struct InitArgs { float *rock, *dirt, *water, *total, *zero[10]; int nzero; };

__global__ void __launch_bounds__(256) k_heightmap(SlabDom d, hg_map_settings_data cfg, InitArgs A) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int gy = d.row0 - HG_HALO_ROWS + (int)(blockIdx.y * blockDim.y + threadIdx.y);
    if (x >= d.W || gy < 0 || gy >= d.H || gy >= d.row0 + d.rows + HG_HALO_ROWS) return;
    float rock, dirt;
    hg_heightmap_cell(cfg, x, gy, d.W, d.H, rock, dirt);
    size_t i = sidx(d, x, gy);
    A.rock[i] = rock;
    A.dirt[i] = dirt;
    A.water[i] = 0.0f;
    if (A.total) A.total[i] = rock + dirt + 0.0f;
    for (int k = 0; k < A.nzero; k++) A.zero[k][i] = 0.0f;
}

#ifndef HG_RAIN_ROWS_PER_THREAD
#define HG_RAIN_ROWS_PER_THREAD 4
#endif
struct RainArgs { const float *rock, *dirt, *water, *total; float *o_rock, *o_dirt, *o_water, *o_total; };

__global__ void __launch_bounds__(256) k_rain(SlabDom d, hg_rain_data set, hg_map_settings_data map_set, float time, RainArgs A) {
    // the permute tables of the table-form simplex noise (hg_noise.cuh): permute(k), k = 0..579, as int and as float
    __shared__ int perm_i[HG_PERM_N];
    __shared__ float perm_f[HG_PERM_N];
    __shared__ __align__(16) HgGrad perm_g[HG_PERM_N];      // gradient of permute(k): one 16-byte load per simplex corner
    for (int k = threadIdx.y * blockDim.x + threadIdx.x; k < HG_PERM_N; k += blockDim.x * blockDim.y) {
        const float p = hg_permute((float)k);
        perm_f[k] = p; perm_i[k] = (int)p;
        perm_g[k] = hg_simplex_grad(p);
    }
    __syncthreads();
    const HgPermTab T{perm_i, perm_f, perm_g};
    // HG_RAIN_ROWS_PER_THREAD rows per thread, blockDim.y apart: the table fill above is paid once for that many cells
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= d.W) return;
#pragma unroll 1
    for (int r = 0; r < HG_RAIN_ROWS_PER_THREAD; r++) {
        int gy = d.row0 - HG_HALO_ROWS + (int)((blockIdx.y * HG_RAIN_ROWS_PER_THREAD + r) * blockDim.y + threadIdx.y);
        if (gy < 0 || gy >= d.H || gy >= d.row0 + d.rows + HG_HALO_ROWS) continue;
        size_t i = sidx(d, x, gy);
        float rock = A.rock[i], dirt = A.dirt[i], water = A.water[i];
        // H.a as its last writer left it: (rock + dirt) + water (smoothing.glsl:101, rain.glsl:54,
        // heightmap.glsl:146); read back when the PASSES schedule materialises it
        float total = A.total ? A.total[i] : rock + dirt + water;
        water += hg_rain_cell(set, map_set, time, x, gy, total, &T);
        A.o_rock[i] = rock; A.o_dirt[i] = dirt; A.o_water[i] = water;
        if (A.o_total) A.o_total[i] = rock + dirt + water;
    }
}

}  // namespace

int hg_launch_heightmap(hg_ctx* c) {
    c->p_aos = false;      // droplet mode: the terrain is generated into the planes; everything is overwritten
    SlabDom d{c->g.W, c->g.H, c->g.pitch, c->g.row0, c->g.rows};
    dim3 b(32, 8), g((c->g.W + 31) / 32, (c->g.rows_alloc + 7) / 8);
    InitArgs A{};
    // gen_heightmap writes the WRITE textures of heightmap, velocity, flux, sediment, then swaps all four
    A.rock = hg_cur(c, PL_ROCK, 0); A.dirt = hg_cur(c, PL_DIRT, 0); A.water = hg_cur(c, PL_WATER, 0);
    A.total = c->aux ? hg_total(c, 0) : nullptr;   // kept current whenever the planes exist
    A.nzero = 0;
    for (int p = PL_FL; p <= PL_SD; p++) A.zero[A.nzero++] = hg_cur(c, p, 0);
    if (c->aux) for (int ch = 0; ch < 4; ch++) A.zero[A.nzero++] = hg_vel(c, ch, 0);
    k_heightmap<<<g, b, 0, c->stream>>>(d, c->map, A);
    HG_LAUNCH_CHECK(c);
    c->ri[0] ^= 1; c->ri[1] ^= 1; c->ri[2] ^= 1; c->ri[3] ^= 1;
    return HG_OK;
}

int hg_launch_rain(hg_ctx* c, float time) {
    {   // rain is added to the ghost rows too: the neighbours' edge rows of the last exchange must be there first
        int rcw = hg_slab_wait_pending(c);
        if (rcw) return rcw;
    }
    SlabDom d{c->g.W, c->g.H, c->g.pitch, c->g.row0, c->g.rows};
    dim3 b(32, 8), g((c->g.W + 31) / 32, (c->g.rows_alloc + 8 * HG_RAIN_ROWS_PER_THREAD - 1) / (8 * HG_RAIN_ROWS_PER_THREAD));
    // The reference writes the other heightmap texture and swaps (erosion.cpp:76-89).  Rain is
    // pointwise, so on the FUSED schedule it runs in place: H, F and S then stay in the same
    // ping-pong set, which the fused kernel's nine-plane TMA box needs.
    const bool in_place = c->schedule == HG_SCHEDULE_FUSED && c->erosion_type == HG_GRID;
    const int w = in_place ? 1 : 0;
    RainArgs A{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_total_live(c) ? hg_total(c, 1) : nullptr,
               hg_cur(c, PL_ROCK, w), hg_cur(c, PL_DIRT, w), hg_cur(c, PL_WATER, w), c->aux ? hg_total(c, w) : nullptr};
    k_rain<<<g, b, 0, c->stream>>>(d, c->rain, c->map, time, A);
    HG_LAUNCH_CHECK(c);
    if (!in_place) c->ri[0] ^= 1;
    return HG_OK;
}

int hg_preload_init_rain_kernels(void) {
    cudaFuncAttributes a;
    HG_CUDA(cudaFuncGetAttributes(&a, k_rain));
    HG_CUDA(cudaFuncGetAttributes(&a, k_heightmap));
    return HG_OK;
}
