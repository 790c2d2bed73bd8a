// hg_internal.cuh — context object and device-side layout shared by the .cu files.
//
// Device layout: every field channel is one fp32 plane of
// (rows + 2*HG_HALO_ROWS) x W elements, row-major, ghost rows first.  The nine
// persistent planes (rock, dirt, water, fL, fR, fT, fB, sed_rock, sed_dirt) exist
// twice (ping-pong sets 0/1) inside ONE arena allocation so a single CUDA IPC
// handle exports a slab to its neighbours.  Each field keeps its own read index,
// mirroring gl::Tex_pair (src/shaderprogram.cpp:51-82).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/hydrogen_b200.h"
#include "hg_cell.cuh"

enum HgPlane {
    PL_ROCK = 0, PL_DIRT = 1, PL_WATER = 2,
    PL_FL = 3, PL_FR = 4, PL_FT = 5, PL_FB = 6,
    PL_SR = 7, PL_SD = 8,
    HG_NPLANES = 9
};
// aux planes, allocated on first use of the PASSES schedule or particle mode
enum HgAuxPlane {
    AX_TOTAL0 = 0, AX_TOTAL1 = 1,                 // H.a, ping-pong with the heightmap
    AX_V0 = 2,                                    // velocity / momentum: 4 channels x 2 sets
    AX_TC = 10, AX_TD = 14,                       // thermal outflow, 4 channels each
    HG_NAUX = 18
};

struct HgGeom {
    int W, H;          // global map size
    int row0, rows;    // slab: global rows [row0, row0+rows)
    int pitch;         // floats per row (= W)
    int rows_alloc;    // rows + 2*HG_HALO_ROWS
    size_t plane_elems;
};

// Every rank's slab as seen from this process: arena base (both plane sets + flag page),
// first owned row, owned rows.  n == 0 until hg_slab_connect*; entry `me` is this context.
struct HgSlabTable {
    float* arena[HG_MAX_SLABS];
    int row0[HG_MAX_SLABS];
    int rows[HG_MAX_SLABS];
    int n, me;
};

// pointers to local row 0 of the ghost region (i.e. global row row0 - HG_HALO_ROWS)
struct HgPlaneSet { float* p[HG_NPLANES]; };
struct HgConstPlaneSet { const float* p[HG_NPLANES]; };

struct hg_ctx {
    int device;
    int erosion_type;
    int schedule;
    HgGeom g;
    uint32_t particle_count;
    uint32_t erosion_steps;
    hg_erosion_data erosion;
    hg_rain_data rain;
    hg_map_settings_data map;
    HgStepParams sp;

    float* arena;              // 2 * HG_NPLANES planes + flag page
    size_t arena_bytes;
    float* aux;                // HG_NAUX planes (lazy)
    // read index per field group: 0 H, 1 F, 2 V, 3 S (Tex_pair::idx_read)
    int ri[4];
    // droplet mode: heightmap and momentum map in the reference's TEXTURE layout, one 16-byte texel per cell
    // ((rock, dirt, water, total), (mx, my, acc_x, acc_y)): four images of plane_elems texels [H set 0, H set 1,
    // M set 0, M set 1] (lazy).  p_aos: these images, not the planes, hold the current H and M (hg_particle_layout).
    float4* pa;
    bool p_aos;
    // droplet processing order (hg_particles.cu): droplet ids binned by map tile, so the threads of a warp gather from
    // and scatter to neighbouring texels; rebuilt every p_rebin_period dispatches (HG_DROPS_REBIN, 0 = id order)
    // droplets on row slabs: ownership bytes (1 = this rank moves and erodes the droplet) and every rank's images,
    // droplet array and ownership bytes (hg_slab_connect*)
    unsigned char* p_own;
    float4* peer_pa[HG_MAX_SLABS];
    hg_particle* peer_parts[HG_MAX_SLABS];
    unsigned char* peer_own[HG_MAX_SLABS];
    bool peer_p_ipc[HG_MAX_SLABS];
    uint32_t* p_order;         // particle_count ids, or null
    uint32_t* p_keys;          // bin of every droplet (scratch)
    uint32_t* p_hist;          // bins + block sums (scratch)
    int p_bins_cap;
    int p_rebin_period, p_rebin_age;
    bool p_order_valid;
    hg_particle* particles;
    uint32_t* lockmap;         // unused by the CUDA path (atomics replace the spin lock); kept for layout parity

    cudaStream_t stream;
    bool own_stream;
    cudaEvent_t ev0, ev1;
    cudaEvent_t prof_ev0, prof_ev1;   // set only inside hg_profile_fused
    uint64_t launches;
    // 16 words: [0] far-fetch cells (total), [1] halo/far errors, [2..6] mass (fp64), [8],[9] per-step far counters
    unsigned long long* d_counters;
    unsigned* far_list;        // cells whose back-trace left the on-chip window this step (lazy)
    int far_parity;
    int tune_variant;          // CTA shape of the fused kernel; -1 = default (HG_FUSED_VARIANT env at create)
    int tune_drops_variant;    // droplet-mode tail: 0 = warp-specialised (default), 1 = one warp group of 224 threads (HG_DROPS_VARIANT)
    int tune_seg;              // rows per CTA of the fused kernel; 0 = automatic (HG_FUSED_SEG env at create)
    // balanced partition of the fused step (hg_fused.cu, k_plan_segments): two plans (this step's / next step's), the
    // durations the CTAs reported, which plan is current; no_balance = HG_FUSED_BALANCE=0
    struct HgPlanItem* plan[2];
    unsigned* cta_ns;
    int plan_cur, plan_n;
    bool plan_valid, no_balance;      // plan_valid: plan[plan_cur] is being written on plan_stream (wait for ev_plan)
    cudaStream_t plan_stream;
    cudaEvent_t ev_main, ev_plan;
    float* staging;            // device staging for RGBA pack/unpack
    size_t staging_elems;
    // pipelined host step (hg_step_host_async): copy streams, full-size staging, ordering events
    // five streams: host->device copies | unpack kernels | (the step, on `stream`) | pack kernels | device->host copies;
    // the staging images are double-buffered (call k uses buffer k & 1), so a copy of call k+1 never waits for a kernel of call k
    cudaStream_t up_stream, down_stream, h2d_stream, d2h_stream;
    float* stage_up;           // 2 buffers x 3 fields x rows x W x 4 floats (H, F, S as RGBA32F)
    float* stage_down;
    cudaEvent_t ev_up, ev_comp, ev_packed, ev_down;
    cudaEvent_t ev_h2d[2], ev_unpacked[2], ev_packed_b[2], ev_d2h[2];
    unsigned host_pipe_calls;  // hg_step_host_async calls so far
    bool host_pipe_busy;       // ev_packed / ev_down have been recorded at least once

    // CUDA-GL interop (HG_WITH_GL): registered textures [heightmap | sediment][index in the Tex_pair], staging image
    void* gl_res[2][2];
    float* gl_stage;

    // multi-GPU slabs (hg_slab.cu): every rank's arena, ordered by row0
    HgSlabTable slabs;
    bool slab_ipc[HG_MAX_SLABS];
    uint32_t step_flag;        // halo generation counter
    bool peers_connected;
    unsigned* h_sticky;        // mapped pinned word raised by a timed-out halo wait (hg_slab.cu); d_sticky = its device alias
    unsigned* d_sticky;
    unsigned long long halo_timeout_ns;
    uint32_t pending_gen;      // generation signalled but not yet waited for (hg_slab_wait_pending), 0 = none
    uint32_t fused_push_gen;   // generation the last fused step pushed and signalled from inside its own kernels (hg_fused.cu), 0 = none
    bool no_fused_push;        // HG_FUSED_PUSH=0: keep the separate push kernel (A/B measurements)
};

// my flag word on every other rank (null for myself)
struct HgFlagArgs { unsigned* flag[HG_MAX_SLABS]; int n; };

void hg_set_error(const char* fmt, ...);

#define HG_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) {                                                        \
            hg_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return HG_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define HG_CHECK_CTX(ctx)                                   \
    do {                                                    \
        if (!(ctx)) { hg_set_error("null context"); return HG_ERR_INVALID; } \
        cudaError_t _e = cudaSetDevice((ctx)->device);      \
        if (_e != cudaSuccess) { hg_set_error("cudaSetDevice: %s", cudaGetErrorString(_e)); return HG_ERR_CUDA; } \
    } while (0)

#define HG_LAUNCH_CHECK(ctx)                                \
    do {                                                    \
        (ctx)->launches++;                                  \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) { hg_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); return HG_ERR_CUDA; } \
    } while (0)

static inline float* hg_plane(hg_ctx* c, int set, int plane) {
    return c->arena + ((size_t)set * HG_NPLANES + plane) * c->g.plane_elems;
}
static inline float* hg_aux_plane(hg_ctx* c, int idx) { return c->aux + (size_t)idx * c->g.plane_elems; }
static inline int hg_field_of_plane(int plane) { return plane <= PL_WATER ? 0 : plane <= PL_FB ? 1 : 3; }
// current read (rd=1) or write (rd=0) plane of the persistent state
static inline float* hg_cur(hg_ctx* c, int plane, int rd) {
    int f = hg_field_of_plane(plane);
    int set = rd ? c->ri[f] : 1 - c->ri[f];
    return hg_plane(c, set, plane);
}
static inline float* hg_total(hg_ctx* c, int rd) { return hg_aux_plane(c, rd ? AX_TOTAL0 + c->ri[0] : AX_TOTAL0 + 1 - c->ri[0]); }
static inline float* hg_vel(hg_ctx* c, int ch, int rd) { return hg_aux_plane(c, AX_V0 + 4 * (rd ? c->ri[2] : 1 - c->ri[2]) + ch); }

static inline float4* hg_pa_h(hg_ctx* c, int rd) { return c->pa + (size_t)(rd ? c->ri[0] : 1 - c->ri[0]) * c->g.plane_elems; }
static inline float4* hg_pa_m(hg_ctx* c, int rd) { return c->pa + (size_t)(2 + (rd ? c->ri[2] : 1 - c->ri[2])) * c->g.plane_elems; }
// Droplet mode keeps H and M either as SoA planes (the PASSES kernels, rain, mass, heightmap init) or as texture-layout
// images (the droplet kernels and the fused thermal/smoothing tail); converts when the other form is asked for.
int hg_particle_layout(hg_ctx* c, bool want_aos);

int hg_ensure_aux(hg_ctx* c);
int hg_fill_total(hg_ctx* c);
// the H.a planes are maintained only by the PASSES schedule and by particle mode
static inline bool hg_total_live(const hg_ctx* c) { return c->aux && (c->schedule == HG_SCHEDULE_PASSES || c->erosion_type == HG_PARTICLES); }

// implemented per file
int hg_launch_passes_step(hg_ctx* c);
int hg_launch_fused_thermal_smooth_particle(hg_ctx* c);
int hg_launch_pass(hg_ctx* c, int pass);
int hg_launch_fused_step(hg_ctx* c);
int hg_launch_rain(hg_ctx* c, float time);
int hg_launch_heightmap(hg_ctx* c);
int hg_launch_particle_spawn(hg_ctx* c, float time, int should_rain);   // slabs: respawn + hand-over before the move
int hg_particle_own_init(hg_ctx* c);
int hg_particle_order_alloc(hg_ctx* c, int nbins);
int hg_slab_push_images(hg_ctx* c);   // droplet slabs: edge rows of the H and M images to the neighbours' ghost rows + signal
int hg_launch_particle_move(hg_ctx* c, float time, int should_rain);
int hg_launch_particle_erode(hg_ctx* c);
int hg_launch_thermal_smooth_particle(hg_ctx* c);
int hg_preload_fused_kernels(void);      // force-load kernels (lazy module loading must not happen under a spinning halo wait)
int hg_preload_particle_kernels(void);
int hg_preload_init_rain_kernels(void);
int hg_preload_context_kernels(void);
int hg_slab_exchange(hg_ctx* c);     // push edge rows to neighbours + wait (no-op without peers)
// fused push (hg_fused.cu): the neighbours' planes of the set this step WRITES, pre-offset so that this slab's element
// index (ghost rows included) of an edge row addresses the matching ghost row; mask bit 0 / 1: a slab below / above exists
int hg_slab_peer_planes(const hg_ctx* c, float* out[2][HG_NPLANES], int* mask);
void hg_slab_signal_args(const hg_ctx* c, HgFlagArgs* out);
int hg_slab_barrier(hg_ctx* c, bool push);   // generation signal + all-rank wait, with or without the edge-row push
int hg_slab_wait_pending(hg_ctx* c); // enqueue the wait for the last signalled generation (before rain / a step touches the planes)
int hg_slab_check_sticky(hg_ctx* c); // HG_ERR_STATE once a halo wait has timed out on this context
void hg_slab_disconnect(hg_ctx* c);
