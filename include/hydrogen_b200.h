/* hydrogen_b200.h — C ABI of the B200 erosion path (libhydrogen_b200.so).
 *
 * The reference has no FFI: its erosion path is five C++ free functions and two
 * structs behind an OpenGL context.  Each entry point below
 * names the reference symbol it replaces.  A header-only C++ shim
 * (host/hydrogen_erosion.hpp) re-creates `namespace Erosion` / `State::World` with
 * the reference signatures on top of this ABI; INTEGRATION.md shows the patch
 * to src/main.cpp.
 *
 * Conventions: opaque handle; every call returns HG_OK (0) or an error code
 * and records a message readable with hg_last_error(); no torch / CUDA types
 * in the signatures (a stream is passed as void*).  A handle is not
 * thread-safe; all device work is enqueued on the handle's stream and is
 * asynchronous unless stated ("blocking").  There is NO CPU fallback: without a
 * CUDA device hg_create fails.
 *
 * Field images cross the boundary in the reference's texture format:
 * RGBA32F, row-major [y][x][4] (State::World::Textures, src/state.hpp:59-76):
 *   HG_FIELD_HEIGHTMAP (rock, dirt, water, total)   heightmap.glsl:138-143
 *   HG_FIELD_FLUX      (fL, fR, fT, fB)             hydro_flux.glsl:14
 *   HG_FIELD_VELOCITY  (u, v, d1+d2, 0) / momentum (mx, my, acc_x, acc_y) in particle mode
 *   HG_FIELD_SEDIMENT  (rock-sed, dirt-sed, 0, 0)
 *   HG_FIELD_THERMAL_C (L, R, T, B) / HG_FIELD_THERMAL_D (LT, RT, LB, RB)
 * Device-side they are fp32 SoA planes.
 */
#ifndef HYDROGEN_B200_H
#define HYDROGEN_B200_H

#include <stddef.h>
#include <stdint.h>
#include "hg_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hg_ctx hg_ctx;

enum hg_status {
    HG_OK = 0,
    HG_ERR_INVALID = 1,      /* bad argument (NULL, size not a multiple of 8, unknown field ...) */
    HG_ERR_CUDA = 2,         /* a CUDA runtime call failed; message in hg_last_error() */
    HG_ERR_STATE = 3,        /* call not valid in this mode / field not materialised */
    HG_ERR_NO_DEVICE = 4     /* no usable CUDA device: there is no CPU fallback */
};

/* Erosion::Programs::Erosion_type (src/erosion.hpp:28-31) */
enum hg_erosion_type { HG_GRID = 0, HG_PARTICLES = 1 };

enum hg_field {
    HG_FIELD_HEIGHTMAP = 0, HG_FIELD_FLUX = 1, HG_FIELD_VELOCITY = 2, HG_FIELD_SEDIMENT = 3,
    HG_FIELD_THERMAL_C = 4, HG_FIELD_THERMAL_D = 5
};

/* How hg_dispatch_grid executes a step.  Both give identical bits.
 *  FUSED : one row-marching kernel per step on the 9 persistent planes (the product path).
 *  PASSES: the reference's 8 dispatches one kernel each (src/erosion.cpp:158-200),
 *          materialising V, TC, TD so they can be downloaded; the validation path. */
enum hg_schedule { HG_SCHEDULE_FUSED = 0, HG_SCHEDULE_PASSES = 1 };

/* The reference's individual dispatches, for hg_dispatch_pass (PASSES schedule only). */
enum hg_pass { HG_PASS_FLUX = 0, HG_PASS_EROSION = 1, HG_PASS_SEDIMENT = 2, HG_PASS_THERMAL = 3, HG_PASS_SMOOTH = 4 };

/* ---- lifetime: State::World::gen_textures + Erosion::setup_shaders + State::setup_settings
 *      (src/state.cpp:3-44, src/erosion.cpp:21-74, src/state.cpp:57-106) ----
 * map_w, map_h: multiples of 8 (the reference dispatches map/8 groups, erosion.cpp:96-97).
 * Settings start at the reference defaults for the mode; fields start zeroed.
 * device: CUDA ordinal.  Returns NULL on failure (see hg_last_error). */
hg_ctx* hg_create(uint32_t map_w, uint32_t map_h, uint32_t particle_count, int erosion_type, int device);
/* Row slab [row0, row0+rows) of a map_w x map_h map for one rank of a multi-GPU run
 * (DESIGN.md §Multi-GPU); rank/world only label the slab. */
hg_ctx* hg_create_slab(uint32_t map_w, uint32_t map_h, uint32_t row0, uint32_t rows,
                       uint32_t particle_count, int erosion_type, int device);
/* State::World::delete_textures + State::delete_settings (src/state.cpp:46-55, 108-113) */
void hg_destroy(hg_ctx* ctx);
const char* hg_last_error(void);
/* library / build identification, e.g. "hydrogen_b200 0.1 sm_100a" */
const char* hg_version(void);

/* ---- settings: Erosion_settings/Rain_settings/Map_settings::push_data()
 *      (src/settings.hpp:22,56,64): the whole struct is replaced between steps ---- */
int hg_set_erosion(hg_ctx* ctx, const hg_erosion_data* data);
int hg_set_rain(hg_ctx* ctx, const hg_rain_data* data);
int hg_set_map(hg_ctx* ctx, const hg_map_settings_data* data);
int hg_get_erosion(hg_ctx* ctx, hg_erosion_data* out);
int hg_get_rain(hg_ctx* ctx, hg_rain_data* out);
int hg_get_map(hg_ctx* ctx, hg_map_settings_data* out);
int hg_set_schedule(hg_ctx* ctx, int schedule);

/* ---- State::World::gen_heightmap (src/state.cpp:116-147 -> glsl/heightmap.glsl) ---- */
int hg_gen_heightmap(hg_ctx* ctx);

/* ---- per-step dispatch (src/erosion.hpp:45-47) ----
 * `time` replaces State::World::Textures::time, which the reference fills from the
 * wall clock (main.cpp:290). */
int hg_dispatch_grid_rain(hg_ctx* ctx, float time);                 /* Erosion::dispatch_grid_rain */
int hg_dispatch_grid(hg_ctx* ctx);                                  /* Erosion::dispatch_grid */
int hg_dispatch_particle(hg_ctx* ctx, float time, int should_rain); /* Erosion::dispatch_particle */
int hg_dispatch_pass(hg_ctx* ctx, int pass);                        /* one dispatch of erosion.cpp:158-200 */
int hg_dispatch_particle_pass(hg_ctx* ctx, int which, float time, int should_rain); /* 0 move, 1 erode */

/* The erosion part of the main loop (src/main.cpp:310-324) for n_steps iterations:
 * erosion_steps++, rain when should_rain and erosion_steps % period == 0, then the
 * step.  time of iteration k (0-based) = time0 + k*dtime. */
int hg_run(hg_ctx* ctx, uint32_t n_steps, float time0, float dtime, int should_rain);
int hg_get_steps(hg_ctx* ctx, uint32_t* erosion_steps);
int hg_set_steps(hg_ctx* ctx, uint32_t erosion_steps);

/* ---- field transfer (the reference never reads fields back; the renderer samples the
 *      read textures, src/rendering.cpp:104-105).  Host buffers hold the slab's rows:
 *      rows*map_w*4 floats.  Blocking. ---- */
int hg_upload(hg_ctx* ctx, int field, const float* src_rgba32f);
int hg_download(hg_ctx* ctx, int field, float* dst_rgba32f);
int hg_upload_particles(hg_ctx* ctx, const hg_particle* src, uint32_t count);
int hg_download_particles(hg_ctx* ctx, hg_particle* dst, uint32_t count);
/* Asynchronous variants on the handle's stream; host memory should be pinned
 * (hg_host_alloc) for the copy to overlap. */
int hg_upload_async(hg_ctx* ctx, int field, const float* src_rgba32f);
int hg_download_async(hg_ctx* ctx, int field, float* dst_rgba32f);
/* One Erosion::dispatch_grid (src/erosion.cpp:158-200) from HOST images to HOST images in the
 * reference's texture format: uploads H, F, S (rows*map_w*4 floats each), steps, downloads
 * H, F, S.  Asynchronous and pipelined over three streams: consecutive calls overlap their
 * PCIe transfers in both directions with the step.  Neither inputs nor outputs may be
 * touched before hg_sync; outputs of one call must not be inputs of the next (use two sets).
 * Grid contexts on the FUSED schedule. */
int hg_step_host_async(hg_ctx* ctx, const float* in_h, const float* in_f, const float* in_s,
                       float* out_h, float* out_f, float* out_s);
void* hg_host_alloc(size_t bytes);   /* pinned host memory */
void hg_host_free(void* p);

/* Sum of rock, dirt, water, rock-sediment, dirt-sediment over the slab (fp64): the
 * mass diagnostic of the 1000-step drift test.  Blocking. */
int hg_mass(hg_ctx* ctx, double out5[5]);

/* ---- checkpoint (no counterpart in the reference: its fields live in GL textures and are lost
 *      at exit, src/main.cpp:334-345).  One file holds the settings structs (byte images of
 *      bindings.glsl:39-99), the step counter that drives the rain schedule (src/main.cpp:315-319),
 *      the slab's fields as RGBA32F little-endian images (grid: heightmap, flux, sediment;
 *      particles: heightmap, momentum map) and the droplet SSBO; layout in
 *      hydro_gen_b200/csrc/hg_checkpoint.cu, reader/writer in hydro_gen_b200/checkpoint.py.
 *      A run resumed from it continues bit for bit.  Load needs a context of the same geometry
 *      and mode.  Blocking. ---- */
int hg_checkpoint_save(hg_ctx* ctx, const char* path);
int hg_checkpoint_load(hg_ctx* ctx, const char* path);

/* ---- streams, sync, timing ---- */
int hg_sync(hg_ctx* ctx);
int hg_set_stream(hg_ctx* ctx, void* cuda_stream);   /* run on the caller's cudaStream_t */
void* hg_get_stream(hg_ctx* ctx);
/* CUDA-event timing on the handle's stream: start, enqueue work, stop (blocking), ms. */
int hg_timer_start(hg_ctx* ctx);
int hg_timer_stop(hg_ctx* ctx, float* elapsed_ms);
/* Average duration of the fused step kernel alone over n_steps real steps (CUDA events around
 * that one launch on the handle's stream); the roofline figure of bench.py.  Blocking. */
int hg_profile_fused(hg_ctx* ctx, uint32_t n_steps, float* avg_kernel_ms);
/* hg_run (src/main.cpp:310-324) that also times the fused step kernel of every iteration with one CUDA event pair per
 * iteration on the handle's stream, without synchronising inside the run: the kernel's average duration inside a real
 * run (bench.py's roofline figure, taken over its timed region), and the duration of the whole run between two events
 * on the stream (before the first launch, after the last).  Blocking. */
int hg_run_profiled(hg_ctx* ctx, uint32_t n_steps, float time0, float dtime, int should_rain, float* avg_kernel_ms, float* total_ms);
/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
uint64_t hg_launch_count(hg_ctx* ctx);
/* Cells whose sediment back-trace left the on-chip window and took the far-fetch path,
 * summed since the last call (fused schedule). Blocking. */
int hg_far_fetch_count(hg_ctx* ctx, uint64_t* cells);

/* ---- multi-GPU slabs (one process per GPU) ----
 * A slab keeps HG_HALO_ROWS ghost rows above and below.  The neighbours' ghost rows
 * are written directly over NVLink: each rank exports a CUDA IPC handle of its plane
 * storage, the host layer all-gathers them, and hg_slab_connect opens the two
 * neighbours' allocations.  After that every step pushes its edge rows into the
 * neighbours' ghost rows and waits on a device-side step flag; no host round trip. */
#define HG_HALO_ROWS 8
#define HG_IPC_HANDLE_BYTES 64
typedef struct hg_slab_export {
    unsigned char mem_handle[HG_IPC_HANDLE_BYTES];   /* cudaIpcMemHandle_t of the slab arena */
    uint64_t arena_bytes;
    uint32_t row0, rows, map_w, map_h;
    int32_t device;
    uint32_t _pad;
} hg_slab_export;
#define HG_MAX_SLABS 16
int hg_slab_export_handle(hg_ctx* ctx, hg_slab_export* out);
/* all[0..n): the exports of every rank, ordered by row0; my_index = this context's entry.
 * Neighbours (my_index +- 1) receive this slab's edge rows every step; every rank's
 * arena is mapped so a sediment back-trace that leaves the slab can be resolved by a
 * direct peer load (far fetch). */
int hg_slab_connect(hg_ctx* ctx, const hg_slab_export* all, int n, int my_index);
/* Droplet mode on slabs: every rank keeps the whole droplet array and owns the droplets whose position
 * lies in its rows; droplets that respawn or drift across a slab edge are handed over through peer pointers, corner
 * texels in a neighbour's rows are eroded in the neighbour's image (NVLink atomics), and the heightmap / momentum
 * images exchange their edge rows after the erode pass and after the thermal/smoothing tail.  Between processes the
 * three extra allocations travel like the arena: export, all-gather, connect (after hg_slab_connect). */
typedef struct hg_slab_export_particles_t {
    unsigned char images_handle[HG_IPC_HANDLE_BYTES];
    unsigned char droplets_handle[HG_IPC_HANDLE_BYTES];
    unsigned char owners_handle[HG_IPC_HANDLE_BYTES];
    uint32_t particle_count;
    uint32_t _pad;
} hg_slab_export_particles_t;
int hg_slab_export_particles(hg_ctx* ctx, hg_slab_export_particles_t* out);
int hg_slab_connect_particles(hg_ctx* ctx, const hg_slab_export_particles_t* all, int n, int my_index);
/* 1 byte per droplet: 1 = this slab owns (moves and erodes) the droplet and holds its current state.  Blocking. */
int hg_slab_particle_owners(hg_ctx* ctx, unsigned char* dst, uint32_t count);
/* Same, for slabs that live in THIS process (one host thread driving several GPUs, or
 * several slabs on one GPU in the tests): CUDA IPC handles cannot be opened by the
 * process that created them. */
int hg_slab_connect_local(hg_ctx* ctx, hg_ctx* const* all, int n, int my_index);
/* Fill the ghost rows from host arrays instead (testing): rows_rgba32f holds
 * HG_HALO_ROWS rows of the field below (side 0) or above (side 1) the slab. */
int hg_slab_set_ghost(hg_ctx* ctx, int field, int side, const float* rows_rgba32f);
/* Halo waits that timed out + far fetches that found no owner, since creation. Blocking.
 * A timed-out wait (a rank more than HG_HALO_TIMEOUT_S seconds behind, default 60) is also STICKY:
 * from then on hg_run, hg_dispatch_grid, hg_dispatch_grid_rain and hg_sync on this context return
 * HG_ERR_STATE, because its ghost rows are stale. */
int hg_slab_errors(hg_ctx* ctx, uint64_t* count);
/* Re-fill the ghost rows after the owned rows were replaced from outside (hg_upload on a connected
 * slab): every rank pushes its edge rows and waits for its neighbours'.  Collective over the slab
 * table: every rank calls it once, at the same point of its schedule.  hg_checkpoint_load calls it
 * itself on a connected slab (so loading is collective too).  No-op on an unconnected context. */
int hg_slab_refresh_halo(hg_ctx* ctx);

/* ---- publishing the fields to a renderer.  The reference's only consumer of the fields is its renderer,
 * which samples the READ textures of the heightmap and sediment pairs (src/rendering.cpp:103-104,
 * gl::Tex_pair::get_read_tex, src/shaderprogram.cpp:51-82).
 *
 * hg_pack_device: the field as a linear image in DEVICE memory in the reference's texture format (RGBA32F,
 * [row][x][4], rows*map_w*4 floats, H.a included): what hg_publish_gl copies into the GL texture, and the
 * entry for any other interop (external memory, a CUDA renderer).  Asynchronous on the handle's stream.
 *
 * hg_register_gl / hg_publish_gl / hg_unregister_gl: CUDA-GL interop with the reference's own textures, compiled
 * with -DHG_WITH_GL (cudaGraphicsGLRegisterImage on the four RGBA32F GL_TEXTURE_2D names of the two pairs; the GL
 * context must be current on the calling thread and live on the handle's device).  read_index: bit 0 = index
 * inside the heightmap pair of the texture the renderer samples next, bit 1 = the same for the sediment pair.
 * A slab publishes its rows at their place.  Without HG_WITH_GL (this image has no OpenGL) they return
 * HG_ERR_STATE. ---- */
int hg_pack_device(hg_ctx* ctx, int field, float* dst_rgba32f_device);
int hg_register_gl(hg_ctx* ctx, const unsigned heightmap_tex[2], const unsigned sediment_tex[2]);
int hg_publish_gl(hg_ctx* ctx, int read_index);
int hg_unregister_gl(hg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HYDROGEN_B200_H */
