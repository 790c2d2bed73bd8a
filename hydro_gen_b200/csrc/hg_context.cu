// hg_context.cu — the C ABI of include/hydrogen_b200.h: context lifetime, settings,
// field transfer in the reference's RGBA32F texture format, the per-step dispatch
// schedule of src/erosion.cpp:76-200 and the main-loop rule of src/main.cpp:310-324.
#include <stdarg.h>
#include <stdlib.h>
#include <new>
#include <vector>
#include "hg_internal.cuh"

static thread_local char g_err[512] = "";

void hg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* hg_last_error(void) { return g_err; }
#ifdef HG_CONTRACTED
extern "C" const char* hg_version(void) { return "hydrogen_b200 0.2 sm_100a contracted (-fmad=true: within tolerance of the reference, not bit-identical)"; }
#else
extern "C" const char* hg_version(void) { return "hydrogen_b200 0.2 sm_100a"; }
#endif

// ------------------------------------------------------------------ lifetime

static void free_ctx(hg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    hg_unregister_gl(c);
    hg_slab_disconnect(c);
    if (c->arena) cudaFree(c->arena);
    if (c->aux) cudaFree(c->aux);
    if (c->particles) cudaFree(c->particles);
    if (c->pa) cudaFree(c->pa);
    if (c->p_own) cudaFree(c->p_own);
    if (c->p_order) cudaFree(c->p_order);
    if (c->p_keys) cudaFree(c->p_keys);
    if (c->p_hist) cudaFree(c->p_hist);
    if (c->lockmap) cudaFree(c->lockmap);
    if (c->staging) cudaFree(c->staging);
    if (c->stage_up) cudaFree(c->stage_up);
    if (c->stage_down) cudaFree(c->stage_down);
    if (c->up_stream) cudaStreamDestroy(c->up_stream);
    if (c->down_stream) cudaStreamDestroy(c->down_stream);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (int b = 0; b < 2; b++) {
        if (c->ev_h2d[b]) cudaEventDestroy(c->ev_h2d[b]);
        if (c->ev_unpacked[b]) cudaEventDestroy(c->ev_unpacked[b]);
        if (c->ev_packed_b[b]) cudaEventDestroy(c->ev_packed_b[b]);
        if (c->ev_d2h[b]) cudaEventDestroy(c->ev_d2h[b]);
    }
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    if (c->ev_comp) cudaEventDestroy(c->ev_comp);
    if (c->ev_packed) cudaEventDestroy(c->ev_packed);
    if (c->ev_down) cudaEventDestroy(c->ev_down);
    if (c->h_sticky) cudaFreeHost(c->h_sticky);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->far_list) cudaFree(c->far_list);
    if (c->plan_stream) { cudaStreamSynchronize(c->plan_stream); cudaStreamDestroy(c->plan_stream); }
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_plan) cudaEventDestroy(c->ev_plan);
    if (c->plan[0]) cudaFree(c->plan[0]);
    if (c->cta_ns) cudaFree(c->cta_ns);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int refresh_params(hg_ctx* c) {
    c->sp = hg_make_step_params(c->erosion);
    return HG_OK;
}

int hg_ensure_aux(hg_ctx* c) {
    if (c->aux) return HG_OK;
    size_t bytes = (size_t)HG_NAUX * c->g.plane_elems * sizeof(float);
    HG_CUDA(cudaMalloc(&c->aux, bytes));
    HG_CUDA(cudaMemsetAsync(c->aux, 0, bytes, c->stream));
    return HG_OK;
}

static int create_impl(hg_ctx* c) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        hg_set_error("no CUDA device available (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return HG_ERR_NO_DEVICE;
    }
    if (c->device < 0 || c->device >= ndev) { hg_set_error("device %d out of range (%d devices)", c->device, ndev); return HG_ERR_INVALID; }
    HG_CUDA(cudaSetDevice(c->device));
    HG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    HG_CUDA(cudaEventCreate(&c->ev0));
    HG_CUDA(cudaEventCreate(&c->ev1));
    // arena = 2 sets x 9 planes, then one 4 KiB page of halo flags
    c->arena_bytes = (size_t)2 * HG_NPLANES * c->g.plane_elems * sizeof(float) + 4096;
    HG_CUDA(cudaMalloc(&c->arena, c->arena_bytes));
    HG_CUDA(cudaMemsetAsync(c->arena, 0, c->arena_bytes, c->stream));
    HG_CUDA(cudaMalloc(&c->d_counters, 16 * sizeof(unsigned long long)));
    HG_CUDA(cudaMemsetAsync(c->d_counters, 0, 16 * sizeof(unsigned long long), c->stream));
    if (c->particle_count) {
        // gl::gen_buffer(particle_buffer, particle_count * sizeof(Particle)), state.cpp:28-30; zeroed (hazard 6)
        HG_CUDA(cudaMalloc(&c->particles, (size_t)c->particle_count * sizeof(hg_particle)));
        HG_CUDA(cudaMemsetAsync(c->particles, 0, (size_t)c->particle_count * sizeof(hg_particle), c->stream));
    }
    if (c->erosion_type == HG_PARTICLES) {
        int rc = hg_ensure_aux(c);   // momentum map + thermal planes
        if (rc) return rc;
        if (c->g.row0 != 0 || c->g.rows != c->g.H) {
            // a droplet SLAB: peers take the addresses of the texel images and of the ownership bytes at connect time
            HG_CUDA(cudaMalloc(&c->pa, (size_t)4 * c->g.plane_elems * sizeof(float4)));
            HG_CUDA(cudaMemsetAsync(c->pa, 0, (size_t)4 * c->g.plane_elems * sizeof(float4), c->stream));
            HG_CUDA(cudaMalloc(&c->p_own, c->particle_count));
            HG_CUDA(cudaMemsetAsync(c->p_own, 0, c->particle_count, c->stream));
            int rco = hg_particle_order_alloc(c, (1 << 21) + 1);
            if (rco) return rco;
        }
    }
    HG_CUDA(cudaStreamSynchronize(c->stream));
    return HG_OK;
}

extern "C" hg_ctx* hg_create_slab(uint32_t map_w, uint32_t map_h, uint32_t row0, uint32_t rows,
                                  uint32_t particle_count, int erosion_type, int device) {
    if (map_w == 0 || map_h == 0 || map_w % HG_WRKGRP || map_h % HG_WRKGRP) {
        hg_set_error("map size %ux%u must be a positive multiple of %d (the reference dispatches map/8 groups)", map_w, map_h, HG_WRKGRP);
        return nullptr;
    }
    if (map_w > 1u << 20 || map_h > 1u << 20) { hg_set_error("map size %ux%u too large", map_w, map_h); return nullptr; }
    if (rows == 0 || (uint64_t)row0 + rows > map_h) { hg_set_error("slab rows [%u,%u) outside map height %u", row0, row0 + rows, map_h); return nullptr; }
    if (erosion_type != HG_GRID && erosion_type != HG_PARTICLES) { hg_set_error("unknown erosion type %d", erosion_type); return nullptr; }
    if (erosion_type == HG_PARTICLES && particle_count == 0) { hg_set_error("particle mode needs particle_count > 0"); return nullptr; }
    hg_ctx* c = new (std::nothrow) hg_ctx();
    if (!c) { hg_set_error("out of host memory"); return nullptr; }
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->erosion_type = erosion_type;
    c->schedule = HG_SCHEDULE_FUSED;
    c->particle_count = particle_count;
    c->g.W = (int)map_w; c->g.H = (int)map_h;
    c->g.row0 = (int)row0; c->g.rows = (int)rows;
    c->g.pitch = (int)map_w;
    c->g.rows_alloc = (int)rows + 2 * HG_HALO_ROWS;
    c->g.plane_elems = (size_t)c->g.rows_alloc * c->g.pitch;
    c->erosion = hg_default_erosion(erosion_type == HG_PARTICLES, particle_count);
    c->rain = hg_default_rain();
    c->map = hg_default_map(0.0f);
    c->tune_variant = -1;
    if (const char* e = getenv("HG_FUSED_SEG")) c->tune_seg = atoi(e);   // tuning aids
    if (const char* e = getenv("HG_FUSED_VARIANT")) c->tune_variant = atoi(e);
    if (const char* e = getenv("HG_FUSED_PUSH")) c->no_fused_push = atoi(e) == 0;
    if (const char* e = getenv("HG_FUSED_BALANCE")) c->no_balance = atoi(e) == 0;
    if (const char* e = getenv("HG_DROPS_VARIANT")) c->tune_drops_variant = atoi(e);
    c->p_rebin_period = 8;
    if (const char* e = getenv("HG_DROPS_REBIN")) c->p_rebin_period = atoi(e);
    refresh_params(c);
    if (create_impl(c) != HG_OK) { free_ctx(c); return nullptr; }
    return c;
}

extern "C" hg_ctx* hg_create(uint32_t map_w, uint32_t map_h, uint32_t particle_count, int erosion_type, int device) {
    return hg_create_slab(map_w, map_h, 0, map_h, particle_count, erosion_type, device);
}

extern "C" void hg_destroy(hg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_ctx(ctx);
}

// ------------------------------------------------------------------ settings

extern "C" int hg_set_erosion(hg_ctx* c, const hg_erosion_data* d) {
    HG_CHECK_CTX(c);
    if (!d) { hg_set_error("null settings"); return HG_ERR_INVALID; }
    c->erosion = *d;
    return refresh_params(c);
}
extern "C" int hg_set_rain(hg_ctx* c, const hg_rain_data* d) {
    HG_CHECK_CTX(c);
    if (!d) { hg_set_error("null settings"); return HG_ERR_INVALID; }
    c->rain = *d;
    return HG_OK;
}
extern "C" int hg_set_map(hg_ctx* c, const hg_map_settings_data* d) {
    HG_CHECK_CTX(c);
    if (!d) { hg_set_error("null settings"); return HG_ERR_INVALID; }
    c->map = *d;
    return HG_OK;
}
extern "C" int hg_get_erosion(hg_ctx* c, hg_erosion_data* o) { HG_CHECK_CTX(c); if (!o) return HG_ERR_INVALID; *o = c->erosion; return HG_OK; }
extern "C" int hg_get_rain(hg_ctx* c, hg_rain_data* o) { HG_CHECK_CTX(c); if (!o) return HG_ERR_INVALID; *o = c->rain; return HG_OK; }
extern "C" int hg_get_map(hg_ctx* c, hg_map_settings_data* o) { HG_CHECK_CTX(c); if (!o) return HG_ERR_INVALID; *o = c->map; return HG_OK; }

extern "C" int hg_set_schedule(hg_ctx* c, int schedule) {
    HG_CHECK_CTX(c);
    if (schedule != HG_SCHEDULE_FUSED && schedule != HG_SCHEDULE_PASSES) { hg_set_error("unknown schedule %d", schedule); return HG_ERR_INVALID; }
    if (schedule == HG_SCHEDULE_PASSES && c->schedule != HG_SCHEDULE_PASSES) {
        if (c->g.row0 != 0 || c->g.rows != c->g.H) { hg_set_error("the PASSES schedule runs on a whole map only"); return HG_ERR_STATE; }
        int rc = hg_ensure_aux(c);
        if (rc) return rc;
        if (c->erosion_type == HG_GRID) {   // H.a was not maintained by the fused schedule
            c->schedule = schedule;
            return hg_fill_total(c);
        }
    }
    c->schedule = schedule;
    return HG_OK;
}

// ------------------------------------------------------------- droplet-mode layouts
namespace {
struct LayoutPlanes { float *rock, *dirt, *water, *total, *m[4]; };
__global__ void __launch_bounds__(256) k_planes_to_aos(LayoutPlanes P, float4* ha, float4* ma, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ha[i] = make_float4(P.rock[i], P.dirt[i], P.water[i], P.total[i]);
        ma[i] = make_float4(P.m[0][i], P.m[1][i], P.m[2][i], P.m[3][i]);
    }
}
__global__ void __launch_bounds__(256) k_aos_to_planes(LayoutPlanes P, const float4* ha, const float4* ma, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 h = ha[i], m = ma[i];
        P.rock[i] = h.x; P.dirt[i] = h.y; P.water[i] = h.z; P.total[i] = h.w;
        P.m[0][i] = m.x; P.m[1][i] = m.y; P.m[2][i] = m.z; P.m[3][i] = m.w;
    }
}
}  // namespace

// The droplet kernels gather and scatter single texels at random positions: with one 16-byte texel per cell a droplet
// touches ~11 DRAM sectors per step instead of 81 with five separate planes, and its deposits are two vector
// reductions per corner instead of five scalar ones.  So in droplet mode the product path (move, erode, fused
// thermal/smoothing tail) works on H and M in the reference's own texture layout; the 1:1 pass kernels, heightmap
// init and the mass diagnostic keep the planes.  Only the read images / planes are converted (a step writes the
// other set completely).
int hg_particle_layout(hg_ctx* c, bool want_aos) {
    if (c->erosion_type != HG_PARTICLES || c->p_aos == want_aos) return HG_OK;
    int rc = hg_ensure_aux(c);
    if (rc) return rc;
    if (!c->pa) {
        HG_CUDA(cudaMalloc(&c->pa, (size_t)4 * c->g.plane_elems * sizeof(float4)));
        HG_CUDA(cudaMemsetAsync(c->pa, 0, (size_t)4 * c->g.plane_elems * sizeof(float4), c->stream));
    }
    LayoutPlanes P{hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1), hg_total(c, 1),
                   {hg_vel(c, 0, 1), hg_vel(c, 1, 1), hg_vel(c, 2, 1), hg_vel(c, 3, 1)}};
    const size_t n = c->g.plane_elems;
    if (want_aos) k_planes_to_aos<<<148 * 8, 256, 0, c->stream>>>(P, hg_pa_h(c, 1), hg_pa_m(c, 1), n);
    else k_aos_to_planes<<<148 * 8, 256, 0, c->stream>>>(P, hg_pa_h(c, 1), hg_pa_m(c, 1), n);
    HG_LAUNCH_CHECK(c);
    c->p_aos = want_aos;
    return HG_OK;
}

// ------------------------------------------------------------- field transfer

namespace {

struct Chan4 { float* p[4]; };   // nullptr = channel not stored (reads as 0 / write ignored)

// dst/src RGBA32F rows [r0, r0+nr) of the slab <-> planes (local row = slab row + HG_HALO_ROWS)
__global__ void __launch_bounds__(256) k_pack(Chan4 ch, int W, int pitch, int r0, int nr, float4* out, int synth_total) {
    size_t n = (size_t)nr * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / W), x = (int)(i - (size_t)r * W);
        size_t s = (size_t)(r0 + r + HG_HALO_ROWS) * pitch + x;
        float4 v;
        v.x = ch.p[0] ? ch.p[0][s] : 0.0f;
        v.y = ch.p[1] ? ch.p[1][s] : 0.0f;
        v.z = ch.p[2] ? ch.p[2][s] : 0.0f;
        v.w = ch.p[3] ? ch.p[3][s] : 0.0f;
        if (synth_total) v.w = v.x + v.y + v.z;   // H.a as the last writer of a step leaves it
        out[i] = v;
    }
}
__global__ void __launch_bounds__(256) k_unpack(Chan4 ch, int W, int pitch, int r0, int nr, const float4* in) {
    size_t n = (size_t)nr * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / W), x = (int)(i - (size_t)r * W);
        size_t s = (size_t)(r0 + r + HG_HALO_ROWS) * pitch + x;
        float4 v = in[i];
        if (ch.p[0]) ch.p[0][s] = v.x;
        if (ch.p[1]) ch.p[1][s] = v.y;
        if (ch.p[2]) ch.p[2][s] = v.z;
        if (ch.p[3]) ch.p[3][s] = v.w;
    }
}

__global__ void __launch_bounds__(256) k_mass(const float* rock, const float* dirt, const float* water, const float* sr, const float* sd,
                                              int W, int pitch, int rows, double* out) {
    double acc[5] = {0, 0, 0, 0, 0};
    size_t n = (size_t)rows * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / W), x = (int)(i - (size_t)r * W);
        size_t s = (size_t)(r + HG_HALO_ROWS) * pitch + x;
        acc[0] += rock[s]; acc[1] += dirt[s]; acc[2] += water[s]; acc[3] += sr[s]; acc[4] += sd[s];
    }
    __shared__ double sh[5][8];
    for (int k = 0; k < 5; k++) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0;
        for (int w = 0; w < 8; w++) v += sh[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, v);
    }
}

int field_channels(hg_ctx* c, int field, Chan4* ch, int* synth_total, bool for_write) {
    *synth_total = 0;
    for (int k = 0; k < 4; k++) ch->p[k] = nullptr;
    switch (field) {
    case HG_FIELD_HEIGHTMAP:
        ch->p[0] = hg_cur(c, PL_ROCK, 1); ch->p[1] = hg_cur(c, PL_DIRT, 1); ch->p[2] = hg_cur(c, PL_WATER, 1);
        if (hg_total_live(c)) ch->p[3] = hg_total(c, 1);
        else if (!for_write) *synth_total = 1;
        return HG_OK;
    case HG_FIELD_FLUX:
        for (int k = 0; k < 4; k++) ch->p[k] = hg_cur(c, PL_FL + k, 1);
        return HG_OK;
    case HG_FIELD_SEDIMENT:
        ch->p[0] = hg_cur(c, PL_SR, 1); ch->p[1] = hg_cur(c, PL_SD, 1);
        return HG_OK;
    case HG_FIELD_VELOCITY:
        if (!c->aux) { hg_set_error("velocity is not materialised by the FUSED schedule (it is dead between steps); use HG_SCHEDULE_PASSES"); return HG_ERR_STATE; }
        for (int k = 0; k < 4; k++) ch->p[k] = hg_vel(c, k, 1);
        return HG_OK;
    case HG_FIELD_THERMAL_C:
    case HG_FIELD_THERMAL_D:
        if (!c->aux) { hg_set_error("thermal outflow is not materialised by the FUSED schedule; use HG_SCHEDULE_PASSES"); return HG_ERR_STATE; }
        for (int k = 0; k < 4; k++) ch->p[k] = hg_aux_plane(c, (field == HG_FIELD_THERMAL_C ? AX_TC : AX_TD) + k);
        return HG_OK;
    default:
        hg_set_error("unknown field %d", field);
        return HG_ERR_INVALID;
    }
}

int ensure_staging(hg_ctx* c) {
    if (c->staging) return HG_OK;
    size_t want = (size_t)32 << 20;                       // 32 Mi floats = 128 MiB
    size_t need = (size_t)c->g.rows * c->g.W * 4;
    c->staging_elems = need < want ? need : want;
    size_t row = (size_t)c->g.W * 4;
    if (c->staging_elems < row) c->staging_elems = row;
    HG_CUDA(cudaMalloc(&c->staging, c->staging_elems * sizeof(float)));
    return HG_OK;
}

int transfer(hg_ctx* c, int field, float* host, bool upload) {
    if (!host) { hg_set_error("null host buffer"); return HG_ERR_INVALID; }
    if (c->p_aos && (field == HG_FIELD_HEIGHTMAP || field == HG_FIELD_VELOCITY)) {
        // droplet mode, texture layout: the image IS the reference's RGBA32F texture; the owned rows are contiguous
        float4* img = (field == HG_FIELD_HEIGHTMAP ? hg_pa_h(c, 1) : hg_pa_m(c, 1)) + (size_t)HG_HALO_ROWS * c->g.pitch;
        const size_t bytes = (size_t)c->g.rows * c->g.W * 4 * sizeof(float);
        if (upload) HG_CUDA(cudaMemcpyAsync(img, host, bytes, cudaMemcpyHostToDevice, c->stream));
        else HG_CUDA(cudaMemcpyAsync(host, img, bytes, cudaMemcpyDeviceToHost, c->stream));
        return HG_OK;
    }
    Chan4 ch; int synth;
    int rc = field_channels(c, field, &ch, &synth, upload);
    if (rc) return rc;
    rc = ensure_staging(c);
    if (rc) return rc;
    int W = c->g.W;
    int rows_per = (int)(c->staging_elems / ((size_t)W * 4));
    for (int r0 = 0; r0 < c->g.rows; r0 += rows_per) {
        int nr = c->g.rows - r0 < rows_per ? c->g.rows - r0 : rows_per;
        size_t n = (size_t)nr * W;
        int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
        float* hp = host + (size_t)r0 * W * 4;
        if (upload) {
            HG_CUDA(cudaMemcpyAsync(c->staging, hp, n * 4 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            k_unpack<<<blocks, 256, 0, c->stream>>>(ch, W, c->g.pitch, r0, nr, reinterpret_cast<const float4*>(c->staging));
            HG_LAUNCH_CHECK(c);
        } else {
            k_pack<<<blocks, 256, 0, c->stream>>>(ch, W, c->g.pitch, r0, nr, reinterpret_cast<float4*>(c->staging), synth);
            HG_LAUNCH_CHECK(c);
            HG_CUDA(cudaMemcpyAsync(hp, c->staging, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    return HG_OK;
}

}  // namespace

int hg_preload_context_kernels(void) {
    cudaFuncAttributes a;
    HG_CUDA(cudaFuncGetAttributes(&a, k_pack));
    HG_CUDA(cudaFuncGetAttributes(&a, k_unpack));
    HG_CUDA(cudaFuncGetAttributes(&a, k_mass));
    HG_CUDA(cudaFuncGetAttributes(&a, k_planes_to_aos));
    HG_CUDA(cudaFuncGetAttributes(&a, k_aos_to_planes));
    return HG_OK;
}

extern "C" int hg_upload_async(hg_ctx* c, int field, const float* src) { HG_CHECK_CTX(c); return transfer(c, field, const_cast<float*>(src), true); }
extern "C" int hg_download_async(hg_ctx* c, int field, float* dst) { HG_CHECK_CTX(c); return transfer(c, field, dst, false); }
extern "C" int hg_upload(hg_ctx* c, int field, const float* src) {
    int rc = hg_upload_async(c, field, src);
    if (rc) return rc;
    HG_CUDA(cudaStreamSynchronize(c->stream));
    return HG_OK;
}
extern "C" int hg_download(hg_ctx* c, int field, float* dst) {
    int rc = hg_download_async(c, field, dst);
    if (rc) return rc;
    HG_CUDA(cudaStreamSynchronize(c->stream));
    return HG_OK;
}

extern "C" int hg_slab_set_ghost(hg_ctx* c, int field, int side, const float* rows) {
    HG_CHECK_CTX(c);
    if (!rows || (side != 0 && side != 1)) { hg_set_error("bad ghost upload"); return HG_ERR_INVALID; }
    Chan4 ch; int synth;
    int rc = field_channels(c, field, &ch, &synth, true);
    if (rc) return rc;
    rc = ensure_staging(c);
    if (rc) return rc;
    size_t n = (size_t)HG_HALO_ROWS * c->g.W;
    if (n * 4 > c->staging_elems) { hg_set_error("staging too small for ghost rows"); return HG_ERR_STATE; }
    HG_CUDA(cudaMemcpyAsync(c->staging, rows, n * 4 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    int r0 = side == 0 ? -HG_HALO_ROWS : c->g.rows;
    k_unpack<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(ch, c->g.W, c->g.pitch, r0, HG_HALO_ROWS, reinterpret_cast<const float4*>(c->staging));
    HG_LAUNCH_CHECK(c);
    HG_CUDA(cudaStreamSynchronize(c->stream));
    return HG_OK;
}

extern "C" int hg_upload_particles(hg_ctx* c, const hg_particle* src, uint32_t count) {
    HG_CHECK_CTX(c);
    if (!src || count > c->particle_count) { hg_set_error("bad particle upload (count %u of %u)", count, c->particle_count); return HG_ERR_INVALID; }
    HG_CUDA(cudaMemcpyAsync(c->particles, src, (size_t)count * sizeof(hg_particle), cudaMemcpyHostToDevice, c->stream));
    c->p_order_valid = false;      // positions changed under the processing order
    if (c->p_own) {                // droplet slab: every slab receives the same array and takes the droplets in its rows
        int rco = hg_particle_own_init(c);
        if (rco) return rco;
    }
    HG_CUDA(cudaStreamSynchronize(c->stream));
    return HG_OK;
}
extern "C" int hg_download_particles(hg_ctx* c, hg_particle* dst, uint32_t count) {
    HG_CHECK_CTX(c);
    if (!dst || count > c->particle_count) { hg_set_error("bad particle download (count %u of %u)", count, c->particle_count); return HG_ERR_INVALID; }
    HG_CUDA(cudaMemcpyAsync(dst, c->particles, (size_t)count * sizeof(hg_particle), cudaMemcpyDeviceToHost, c->stream));
    HG_CUDA(cudaStreamSynchronize(c->stream));
    return HG_OK;
}

extern "C" int hg_slab_particle_owners(hg_ctx* c, unsigned char* dst, uint32_t count) {
    HG_CHECK_CTX(c);
    if (!dst || count > c->particle_count) { hg_set_error("bad owner download (count %u of %u)", count, c->particle_count); return HG_ERR_INVALID; }
    if (!c->p_own) { memset(dst, 1, count); return HG_OK; }      // a whole map owns all its droplets
    HG_CUDA(cudaMemcpyAsync(dst, c->p_own, count, cudaMemcpyDeviceToHost, c->stream));
    HG_CUDA(cudaStreamSynchronize(c->stream));
    for (uint32_t k = 0; k < count; k++) dst[k] = dst[k] ? 1 : 0;      // 2 = handed over during the last erode pass: this slab holds its state
    return HG_OK;
}

extern "C" void* hg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { hg_set_error("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
    return p;
}
extern "C" void hg_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int hg_mass(hg_ctx* c, double out5[5]) {
    HG_CHECK_CTX(c);
    if (!out5) return HG_ERR_INVALID;
    {
        int rcl = hg_particle_layout(c, false);
        if (rcl) return rcl;
    }
    double* d = reinterpret_cast<double*>(c->d_counters + 2);
    HG_CUDA(cudaMemsetAsync(d, 0, 5 * sizeof(double), c->stream));
    k_mass<<<148 * 8, 256, 0, c->stream>>>(hg_cur(c, PL_ROCK, 1), hg_cur(c, PL_DIRT, 1), hg_cur(c, PL_WATER, 1),
                                           hg_cur(c, PL_SR, 1), hg_cur(c, PL_SD, 1), c->g.W, c->g.pitch, c->g.rows, d);
    HG_LAUNCH_CHECK(c);
    HG_CUDA(cudaMemcpyAsync(out5, d, 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HG_CUDA(cudaStreamSynchronize(c->stream));
    return HG_OK;
}

// ---- pipelined host step --------------------------------------------------------------
// One Erosion::dispatch_grid whose inputs come from, and whose results go to, HOST images in the
// reference's texture format (RGBA32F H, F, S).  Three streams: uploads + unpack, the step, pack +
// downloads; consecutive calls overlap (PCIe is full duplex): the upload of call k+1 runs while
// call k computes and downloads.  Host buffers should be pinned (hg_host_alloc); outputs are
// valid after hg_sync.  Ordering: unpack(k+1) overwrites the planes pack(k) reads, so it waits
// for ev_packed; staging buffers are reused in stream order.
static int ensure_host_pipe(hg_ctx* c) {
    if (c->stage_up) return HG_OK;
    const size_t bytes = (size_t)2 * 3 * c->g.rows * c->g.W * 4 * sizeof(float);
    HG_CUDA(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    HG_CUDA(cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
    HG_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    HG_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    HG_CUDA(cudaMalloc(&c->stage_up, bytes));
    HG_CUDA(cudaMalloc(&c->stage_down, bytes));
    HG_CUDA(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
    HG_CUDA(cudaEventCreateWithFlags(&c->ev_comp, cudaEventDisableTiming));
    HG_CUDA(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    HG_CUDA(cudaEventCreate(&c->ev_down));
    for (int b = 0; b < 2; b++) {
        HG_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[b], cudaEventDisableTiming));
        HG_CUDA(cudaEventCreateWithFlags(&c->ev_unpacked[b], cudaEventDisableTiming));
        HG_CUDA(cudaEventCreateWithFlags(&c->ev_packed_b[b], cudaEventDisableTiming));
        HG_CUDA(cudaEventCreateWithFlags(&c->ev_d2h[b], cudaEventDisableTiming));
    }
    return HG_OK;
}

// One Erosion::dispatch_grid from HOST images to HOST images.  Call k (staging buffers k & 1):
//   h2d_stream : [unpack(k-2) done with the buffer]      copy H, F, S into stage_up
//   up_stream  : [copies(k) done] [everything earlier on the main stream done] [pack(k-1) done with the planes]   unpack
//   stream     : [unpack(k) done]                         fused step + halo exchange
//   down_stream: [step(k) done] [copies-out(k-2) done with the buffer]   pack into stage_down
//   d2h_stream : [pack(k) done]                           copy H, F, S out
// so the two PCIe directions run back to back without waiting for a kernel of the neighbouring call: the period of
// the pipeline is the longer of the two copies (bench.py: e2e against e2e.copy_ceiling).
extern "C" int hg_step_host_async(hg_ctx* c, const float* in_h, const float* in_f, const float* in_s,
                                  float* out_h, float* out_f, float* out_s) {
    HG_CHECK_CTX(c);
    if (c->erosion_type != HG_GRID || c->schedule != HG_SCHEDULE_FUSED) { hg_set_error("hg_step_host_async needs a grid context on the FUSED schedule"); return HG_ERR_STATE; }
    if (!in_h || !in_f || !in_s || !out_h || !out_f || !out_s) { hg_set_error("null host image"); return HG_ERR_INVALID; }
    int rc = ensure_host_pipe(c);
    if (rc) return rc;
    const int fields[3] = {HG_FIELD_HEIGHTMAP, HG_FIELD_FLUX, HG_FIELD_SEDIMENT};
    const float* in[3] = {in_h, in_f, in_s};
    float* out[3] = {out_h, out_f, out_s};
    const size_t n = (size_t)c->g.rows * c->g.W;            // texels per field
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    const int b = (int)(c->host_pipe_calls & 1u);
    const bool reuse = c->host_pipe_calls >= 2;              // the buffers of call k-2 are in flight
    float* const sup = c->stage_up + (size_t)b * 3 * n * 4;
    float* const sdn = c->stage_down + (size_t)b * 3 * n * 4;
    // copies in
    if (reuse) HG_CUDA(cudaStreamWaitEvent(c->h2d_stream, c->ev_unpacked[b], 0));
    for (int k = 0; k < 3; k++)
        HG_CUDA(cudaMemcpyAsync(sup + (size_t)k * n * 4, in[k], n * 4 * sizeof(float), cudaMemcpyHostToDevice, c->h2d_stream));
    HG_CUDA(cudaEventRecord(c->ev_h2d[b], c->h2d_stream));
    // unpack: the planes must be free -- everything already enqueued on the main stream (a previous plain dispatch, an
    // upload ...) and the pack of the previous call
    HG_CUDA(cudaEventRecord(c->ev_comp, c->stream));
    HG_CUDA(cudaStreamWaitEvent(c->up_stream, c->ev_comp, 0));
    HG_CUDA(cudaStreamWaitEvent(c->up_stream, c->ev_h2d[b], 0));
    if (c->host_pipe_busy) HG_CUDA(cudaStreamWaitEvent(c->up_stream, c->ev_packed, 0));
    for (int k = 0; k < 3; k++) {
        Chan4 ch; int synth;
        rc = field_channels(c, fields[k], &ch, &synth, true);
        if (rc) return rc;
        k_unpack<<<blocks, 256, 0, c->up_stream>>>(ch, c->g.W, c->g.pitch, 0, c->g.rows, reinterpret_cast<const float4*>(sup + (size_t)k * n * 4));
        HG_LAUNCH_CHECK(c);
    }
    HG_CUDA(cudaEventRecord(c->ev_up, c->up_stream));
    HG_CUDA(cudaEventRecord(c->ev_unpacked[b], c->up_stream));
    // the step
    HG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_up, 0));
    rc = hg_launch_fused_step(c);
    if (rc) return rc;
    rc = hg_slab_exchange(c);
    if (rc) return rc;
    HG_CUDA(cudaEventRecord(c->ev_comp, c->stream));
    // pack
    HG_CUDA(cudaStreamWaitEvent(c->down_stream, c->ev_comp, 0));
    if (reuse) HG_CUDA(cudaStreamWaitEvent(c->down_stream, c->ev_d2h[b], 0));
    for (int k = 0; k < 3; k++) {
        Chan4 ch; int synth;
        rc = field_channels(c, fields[k], &ch, &synth, false);
        if (rc) return rc;
        k_pack<<<blocks, 256, 0, c->down_stream>>>(ch, c->g.W, c->g.pitch, 0, c->g.rows, reinterpret_cast<float4*>(sdn + (size_t)k * n * 4), synth);
        HG_LAUNCH_CHECK(c);
    }
    HG_CUDA(cudaEventRecord(c->ev_packed, c->down_stream));
    HG_CUDA(cudaEventRecord(c->ev_packed_b[b], c->down_stream));
    // copies out
    HG_CUDA(cudaStreamWaitEvent(c->d2h_stream, c->ev_packed_b[b], 0));
    for (int k = 0; k < 3; k++)
        HG_CUDA(cudaMemcpyAsync(out[k], sdn + (size_t)k * n * 4, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->d2h_stream));
    HG_CUDA(cudaEventRecord(c->ev_d2h[b], c->d2h_stream));
    HG_CUDA(cudaEventRecord(c->ev_down, c->d2h_stream));
    c->host_pipe_busy = true;
    c->host_pipe_calls++;
    return HG_OK;
}

// ------------------------------------------------------- streams, sync, timing

extern "C" int hg_sync(hg_ctx* c) {
    HG_CHECK_CTX(c);
    HG_CUDA(cudaStreamSynchronize(c->stream));
    if (c->up_stream) HG_CUDA(cudaStreamSynchronize(c->up_stream));
    if (c->down_stream) HG_CUDA(cudaStreamSynchronize(c->down_stream));
    if (c->h2d_stream) HG_CUDA(cudaStreamSynchronize(c->h2d_stream));
    if (c->d2h_stream) HG_CUDA(cudaStreamSynchronize(c->d2h_stream));
    return hg_slab_check_sticky(c);
}
extern "C" int hg_set_stream(hg_ctx* c, void* s) {
    HG_CHECK_CTX(c);
    HG_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = static_cast<cudaStream_t>(s);
    return HG_OK;
}
extern "C" void* hg_get_stream(hg_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }
extern "C" int hg_timer_start(hg_ctx* c) { HG_CHECK_CTX(c); HG_CUDA(cudaEventRecord(c->ev0, c->stream)); return HG_OK; }
extern "C" int hg_timer_stop(hg_ctx* c, float* ms) {
    HG_CHECK_CTX(c);
    if (c->host_pipe_busy) HG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_down, 0));   // include the last download
    HG_CUDA(cudaEventRecord(c->ev1, c->stream));
    HG_CUDA(cudaEventSynchronize(c->ev1));
    if (ms) HG_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return HG_OK;
}
extern "C" uint64_t hg_launch_count(hg_ctx* c) { return c ? c->launches : 0; }
extern "C" int hg_far_fetch_count(hg_ctx* c, uint64_t* cells) {
    HG_CHECK_CTX(c);
    unsigned long long v = 0;
    HG_CUDA(cudaMemcpyAsync(&v, c->d_counters, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    HG_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(v), c->stream));
    HG_CUDA(cudaStreamSynchronize(c->stream));
    if (cells) *cells = v;
    return HG_OK;
}

// -------------------------------------------------------------------- dispatch

extern "C" int hg_gen_heightmap(hg_ctx* c) { HG_CHECK_CTX(c); return hg_launch_heightmap(c); }

extern "C" int hg_dispatch_grid_rain(hg_ctx* c, float time) {
    HG_CHECK_CTX(c);
    if (c->erosion_type != HG_GRID) { hg_set_error("dispatch_grid_rain on a particle context (prog.grid is null in the reference)"); return HG_ERR_STATE; }
    int rc = hg_slab_check_sticky(c);
    if (rc) return rc;
    rc = hg_launch_rain(c, time);
    if (rc) return rc;
    // On the FUSED schedule rain is added in place to the planes the next step reads, ghost rows included (pointwise
    // in the global coordinate, so no push is needed).  A peer's far fetch of that step reads this rank's water
    // through its peer pointer: an all-rank generation makes sure every rank's rain has landed first.
    return hg_slab_barrier(c, false);
}

extern "C" int hg_dispatch_grid(hg_ctx* c) {
    HG_CHECK_CTX(c);
    if (c->erosion_type != HG_GRID) { hg_set_error("dispatch_grid on a particle context"); return HG_ERR_STATE; }
    if (c->schedule == HG_SCHEDULE_PASSES) return hg_launch_passes_step(c);
    int rc = hg_slab_check_sticky(c);
    if (rc) return rc;
    rc = hg_launch_fused_step(c);
    if (rc) return rc;
    return hg_slab_exchange(c);
}

// Average duration of the fused step kernel alone (CUDA events around that one launch), over
// n_steps real steps: advances the simulation like hg_dispatch_grid.  Blocking.
extern "C" int hg_profile_fused(hg_ctx* c, uint32_t n_steps, float* avg_kernel_ms) {
    HG_CHECK_CTX(c);
    if (c->erosion_type != HG_GRID || c->schedule != HG_SCHEDULE_FUSED || !n_steps) { hg_set_error("hg_profile_fused needs a grid context on the FUSED schedule"); return HG_ERR_STATE; }
    cudaEvent_t e0, e1;
    HG_CUDA(cudaEventCreate(&e0));
    HG_CUDA(cudaEventCreate(&e1));
    double total = 0.0;
    int rc = HG_OK;
    for (uint32_t k = 0; k < n_steps && rc == HG_OK; k++) {
        c->prof_ev0 = e0; c->prof_ev1 = e1;
        rc = hg_launch_fused_step(c);
        c->prof_ev0 = c->prof_ev1 = nullptr;
        if (rc == HG_OK) rc = hg_slab_exchange(c);
        if (rc == HG_OK && cudaEventSynchronize(e1) == cudaSuccess) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (avg_kernel_ms) *avg_kernel_ms = (float)(total / n_steps);
    return rc;
}

extern "C" int hg_dispatch_pass(hg_ctx* c, int pass) {
    HG_CHECK_CTX(c);
    if (c->schedule != HG_SCHEDULE_PASSES) { hg_set_error("hg_dispatch_pass needs HG_SCHEDULE_PASSES"); return HG_ERR_STATE; }
    return hg_launch_pass(c, pass);
}

extern "C" int hg_dispatch_particle_pass(hg_ctx* c, int which, float time, int should_rain) {
    HG_CHECK_CTX(c);
    if (c->erosion_type != HG_PARTICLES) { hg_set_error("particle pass on a grid context"); return HG_ERR_STATE; }
    return which == 0 ? hg_launch_particle_move(c, time, should_rain) : hg_launch_particle_erode(c);
}

// Erosion::dispatch_particle, src/erosion.cpp:132-156
extern "C" int hg_dispatch_particle(hg_ctx* c, float time, int should_rain) {
    HG_CHECK_CTX(c);
    if (c->erosion_type != HG_PARTICLES) { hg_set_error("dispatch_particle on a grid context"); return HG_ERR_STATE; }
    int rc;
    if (c->g.row0 != 0 || c->g.rows != c->g.H) {
        // droplets on row slabs: spawn + hand-over | all ranks | move, erode (+ hand-over of drifted droplets, NVLink
        // atomics for corner texels in a neighbour's rows) | all ranks | edge rows of the eroded images | all ranks |
        // thermal/smoothing tail | edge rows of the new images (the wait sits in front of the next dispatch)
        if (!c->peers_connected) { hg_set_error("a droplet slab must be connected to its peers (hg_slab_connect*) before it steps"); return HG_ERR_STATE; }
        if (c->schedule != HG_SCHEDULE_FUSED) { hg_set_error("droplet slabs run on the FUSED schedule"); return HG_ERR_STATE; }
        rc = hg_slab_check_sticky(c);
        if (rc) return rc;
        rc = hg_particle_layout(c, true);
        if (rc) return rc;
        rc = hg_slab_wait_pending(c);
        if (rc) return rc;
        rc = hg_launch_particle_spawn(c, time, should_rain);
        if (rc) return rc;
        rc = hg_slab_barrier(c, false);
        if (rc) return rc;
        rc = hg_slab_wait_pending(c);
        if (rc) return rc;
        rc = hg_launch_particle_move(c, time, should_rain);
        if (rc) return rc;
        rc = hg_launch_particle_erode(c);
        if (rc) return rc;
        rc = hg_slab_barrier(c, false);
        if (rc) return rc;
        rc = hg_slab_push_images(c);      // waits for the generation above first
        if (rc) return rc;
        rc = hg_launch_fused_thermal_smooth_particle(c);      // waits for the pushed rows
        if (rc) return rc;
        return hg_slab_push_images(c);
    }
    rc = hg_launch_particle_move(c, time, should_rain);
    if (rc) return rc;
    rc = hg_launch_particle_erode(c);
    if (rc) return rc;
    // thermal x2 + smoothing: one fused kernel (TC/TD stay on chip), or the five 1:1 pass kernels on the PASSES schedule
    if (c->schedule == HG_SCHEDULE_FUSED && !getenv("HG_DROPS_PASSES")) return hg_launch_fused_thermal_smooth_particle(c);
    return hg_launch_thermal_smooth_particle(c);
}

// src/main.cpp:310-324
// ev: null, or 2 * n_steps events: the pair (2k, 2k+1) is recorded around the fused step kernel of iteration k
static int run_steps(hg_ctx* c, uint32_t n_steps, float time0, float dtime, int should_rain, cudaEvent_t* ev) {
    for (uint32_t k = 0; k < n_steps; k++) {
        float time = time0 + (float)k * dtime;
        if (ev) { c->prof_ev0 = ev[2 * k]; c->prof_ev1 = ev[2 * k + 1]; }
        c->erosion_steps++;
        int rc;
        if (c->erosion_type == HG_GRID) {
            if (should_rain && c->rain.period != 0 && !(c->erosion_steps % (uint32_t)c->rain.period)) {
                rc = hg_dispatch_grid_rain(c, time);
                if (rc) return rc;
            }
            rc = hg_dispatch_grid(c);
        } else {
            rc = hg_dispatch_particle(c, time, should_rain);
        }
        c->prof_ev0 = c->prof_ev1 = nullptr;
        if (rc) return rc;
    }
    return HG_OK;
}

extern "C" int hg_run(hg_ctx* c, uint32_t n_steps, float time0, float dtime, int should_rain) {
    HG_CHECK_CTX(c);
    return run_steps(c, n_steps, time0, dtime, should_rain, nullptr);
}

// hg_run that also times the fused step kernel of every iteration (one CUDA event pair per iteration on the
// handle's stream, no host synchronisation inside the run) and returns the average: the kernel's duration
// INSIDE a real run, rain, fix-up and planner kernels around it as usual.  Blocking.
extern "C" int hg_run_profiled(hg_ctx* c, uint32_t n_steps, float time0, float dtime, int should_rain, float* avg_kernel_ms, float* total_ms) {
    HG_CHECK_CTX(c);
    if (c->erosion_type != HG_GRID || c->schedule != HG_SCHEDULE_FUSED || !n_steps) { hg_set_error("hg_run_profiled needs a grid context on the FUSED schedule"); return HG_ERR_STATE; }
    std::vector<cudaEvent_t> ev(2 * (size_t)n_steps, nullptr);
    int rc = HG_OK;
    for (auto& e : ev) if (cudaEventCreate(&e) != cudaSuccess) { hg_set_error("cudaEventCreate failed"); rc = HG_ERR_CUDA; break; }
    if (rc == HG_OK) {
        cudaEventRecord(c->ev0, c->stream);         // the whole run, on the device: from before the first launch ...
        rc = run_steps(c, n_steps, time0, dtime, should_rain, ev.data());
        cudaEventRecord(c->ev1, c->stream);         // ... to after the last one
    }
    if (rc == HG_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) { hg_set_error("cudaStreamSynchronize failed"); rc = HG_ERR_CUDA; }
    double total = 0.0;
    if (rc == HG_OK) {
        for (uint32_t k = 0; k < n_steps; k++) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ev[2 * k], ev[2 * k + 1]) == cudaSuccess) total += ms;
        }
    }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (avg_kernel_ms) *avg_kernel_ms = (float)(total / n_steps);
    if (total_ms) { float ms = 0.0f; if (rc == HG_OK) cudaEventElapsedTime(&ms, c->ev0, c->ev1); *total_ms = ms; }
    return rc;
}
extern "C" int hg_get_steps(hg_ctx* c, uint32_t* s) { HG_CHECK_CTX(c); if (!s) return HG_ERR_INVALID; *s = c->erosion_steps; return HG_OK; }
extern "C" int hg_set_steps(hg_ctx* c, uint32_t s) { HG_CHECK_CTX(c); c->erosion_steps = s; return HG_OK; }

// ------------------------------------------------------------- publishing to a renderer
// The only consumer of the fields in the reference is its renderer, which samples the READ textures of the heightmap
// and sediment pairs (src/rendering.cpp:103-104; gl::Tex_pair read index, src/shaderprogram.cpp:51-82).  Two steps:
// (1) pack a field into a linear DEVICE image in the reference's texture format -- RGBA32F, [row][x][4], H.a
//     included -- which is all that any interop (GL, Vulkan/EGL external memory, a CUDA renderer) needs and is what
//     the tests check against hg_download; (2) with -DHG_WITH_GL, copy that image into the mapped GL texture.
extern "C" int hg_pack_device(hg_ctx* c, int field, float* dst_rgba32f_device) {
    HG_CHECK_CTX(c);
    if (!dst_rgba32f_device) { hg_set_error("null device image"); return HG_ERR_INVALID; }
    const size_t n = (size_t)c->g.rows * c->g.W;
    if (c->p_aos && (field == HG_FIELD_HEIGHTMAP || field == HG_FIELD_VELOCITY)) {      // droplet mode: already in texture layout
        const float4* img = (field == HG_FIELD_HEIGHTMAP ? hg_pa_h(c, 1) : hg_pa_m(c, 1)) + (size_t)HG_HALO_ROWS * c->g.pitch;
        HG_CUDA(cudaMemcpyAsync(dst_rgba32f_device, img, n * 4 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return HG_OK;
    }
    Chan4 ch; int synth;
    int rc = field_channels(c, field, &ch, &synth, false);
    if (rc) return rc;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_pack<<<blocks, 256, 0, c->stream>>>(ch, c->g.W, c->g.pitch, 0, c->g.rows, reinterpret_cast<float4*>(dst_rgba32f_device), synth);
    HG_LAUNCH_CHECK(c);
    return HG_OK;
}

#ifdef HG_WITH_GL
#include <cuda_gl_interop.h>
// gl_res[pair][index]: pair 0 = heightmap, 1 = sediment; index = the texture's place in its gl::Tex_pair
static int gl_publish_one(hg_ctx* c, int field, cudaGraphicsResource_t res) {
    int rc = hg_pack_device(c, field, c->gl_stage);
    if (rc) return rc;
    HG_CUDA(cudaGraphicsMapResources(1, &res, c->stream));
    cudaArray_t arr = nullptr;
    cudaError_t e = cudaGraphicsSubResourceGetMappedArray(&arr, res, 0, 0);
    if (e == cudaSuccess)      // a slab publishes its rows at its place in the map-sized texture
        e = cudaMemcpy2DToArrayAsync(arr, 0, (size_t)c->g.row0, c->gl_stage, (size_t)c->g.W * 4 * sizeof(float),
                                     (size_t)c->g.W * 4 * sizeof(float), (size_t)c->g.rows, cudaMemcpyDeviceToDevice, c->stream);
    cudaError_t e2 = cudaGraphicsUnmapResources(1, &res, c->stream);
    if (e != cudaSuccess || e2 != cudaSuccess) { hg_set_error("publishing to the GL texture failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2)); return HG_ERR_CUDA; }
    return HG_OK;
}

extern "C" int hg_register_gl(hg_ctx* c, const unsigned heightmap_tex[2], const unsigned sediment_tex[2]) {
    HG_CHECK_CTX(c);
    if (!heightmap_tex || !sediment_tex) { hg_set_error("null texture names"); return HG_ERR_INVALID; }
    hg_unregister_gl(c);
    const unsigned* names[2] = {heightmap_tex, sediment_tex};
    for (int p = 0; p < 2; p++)
        for (int k = 0; k < 2; k++) {
            cudaGraphicsResource_t r = nullptr;
            // RGBA32F GL_TEXTURE_2D of the map's size (gl::gen_texture, src/shaderprogram.cpp:14-33); CUDA only writes it
            cudaError_t e = cudaGraphicsGLRegisterImage(&r, (GLuint)names[p][k], GL_TEXTURE_2D, cudaGraphicsRegisterFlagsWriteDiscard);
            if (e != cudaSuccess) {
                hg_set_error("cudaGraphicsGLRegisterImage(texture %u) failed: %s (is the GL context current on this thread, on device %d?)", names[p][k], cudaGetErrorString(e), c->device);
                hg_unregister_gl(c);
                return HG_ERR_CUDA;
            }
            c->gl_res[p][k] = r;
        }
    HG_CUDA(cudaMalloc(&c->gl_stage, (size_t)c->g.rows * c->g.W * 4 * sizeof(float)));
    return HG_OK;
}

extern "C" int hg_unregister_gl(hg_ctx* c) {
    if (!c) return HG_ERR_INVALID;
    for (int p = 0; p < 2; p++)
        for (int k = 0; k < 2; k++)
            if (c->gl_res[p][k]) { cudaGraphicsUnregisterResource(static_cast<cudaGraphicsResource_t>(c->gl_res[p][k])); c->gl_res[p][k] = nullptr; }
    if (c->gl_stage) { cudaFree(c->gl_stage); c->gl_stage = nullptr; }
    return HG_OK;
}

// read_index: bit 0 = index of the heightmap texture the renderer will sample (Tex_pair::get_read_tex of
// State::World::Textures::heightmap), bit 1 = the same for the sediment pair.  Asynchronous on the handle's stream;
// GL may sample the textures once the stream has been synchronised (hg_sync) -- the reference's own frame does the
// equivalent with glMemoryBarrier between its dispatches and its draw (src/erosion.cpp:99).
extern "C" int hg_publish_gl(hg_ctx* c, int read_index) {
    HG_CHECK_CTX(c);
    if (!c->gl_stage) { hg_set_error("hg_publish_gl before hg_register_gl"); return HG_ERR_STATE; }
    int rc = gl_publish_one(c, HG_FIELD_HEIGHTMAP, static_cast<cudaGraphicsResource_t>(c->gl_res[0][read_index & 1]));
    if (rc) return rc;
    if (c->erosion_type == HG_PARTICLES) return HG_OK;      // the droplet mode keeps no suspended sediment field
    return gl_publish_one(c, HG_FIELD_SEDIMENT, static_cast<cudaGraphicsResource_t>(c->gl_res[1][(read_index >> 1) & 1]));
}
#else
extern "C" int hg_register_gl(hg_ctx*, const unsigned[2], const unsigned[2]) {
    hg_set_error("built without HG_WITH_GL (no OpenGL in the build image); hg_pack_device gives the same RGBA32F image in device memory");
    return HG_ERR_STATE;
}
extern "C" int hg_unregister_gl(hg_ctx*) { return HG_OK; }
extern "C" int hg_publish_gl(hg_ctx*, int) {
    hg_set_error("built without HG_WITH_GL (no OpenGL in the build image); hg_pack_device gives the same RGBA32F image in device memory");
    return HG_ERR_STATE;
}
#endif
