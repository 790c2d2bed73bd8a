// hg_slab.cu — row-slab halo exchange between GPUs over NVLink peer memory.
//
// The reference is single-GPU; this is new design for BASELINE configs 3
// and 5.  Rank g owns global rows [row0, row0+rows) of all nine planes and keeps
// HG_HALO_ROWS ghost rows on each side.  One fused step consumes 6 ghost rows (flux 1 +
// thermal 2+2 + smooth 1; SURVEY.md §8a "dependency radii"), so ONE exchange per step
// suffices: after its step kernel a rank stores its new edge rows straight into the two
// neighbours' ghost rows through peer-mapped pointers (CUDA IPC between processes, plain
// pointers inside one process) and then publishes the step number in a flag word of
// EVERY rank.  The next step starts with a device-side wait until all ranks have
// published that number.  No host round trip, no collective: the data path is two
// nearest-neighbour stores plus n flag words.
//
// Why all ranks and not just the neighbours: the sediment back-trace is unbounded
// (sediment_transport.glsl:27-28), so the fused kernel's far-fetch path may read the
// PRE-step planes of any rank through its peer pointer.  That read is only safe while no
// rank is a step ahead (it would be overwriting those planes) or behind (still writing
// them); the all-rank flag wait is that guarantee.
#include <stdlib.h>
#include "hg_internal.cuh"

namespace {

struct PushSide {
    float* dst[HG_NPLANES];         // neighbour's planes (same set), local row 0; nullptr = no neighbour on this side
    size_t src_off, dst_off;        // element offsets of the first row to copy
};
typedef HgFlagArgs FlagArgs;
struct PushArgs {
    const float* src[HG_NPLANES];   // my planes (current read set), local row 0
    PushSide side[2];
    size_t n;                       // elements per plane and side (HG_HALO_ROWS * pitch)
    FlagArgs sig;                   // my flag word on every other rank
    unsigned gen;
    unsigned* done;                 // block counter (self-resetting)
};

// One kernel per step: float4 copy of HG_HALO_ROWS rows x 9 planes into each neighbour's ghost
// rows (blockIdx.y = plane, blockIdx.z = side); the last block to finish publishes the step
// number `gen` in this rank's flag word on every rank.
__global__ void __launch_bounds__(256) k_halo_push_signal(PushArgs A) {
    const int plane = blockIdx.y, sd = blockIdx.z;
    if (A.side[sd].dst[plane]) {
        const float4* s = reinterpret_cast<const float4*>(A.src[plane] + A.side[sd].src_off);
        float4* d = reinterpret_cast<float4*>(A.side[sd].dst[plane] + A.side[sd].dst_off);
        size_t n4 = A.n / 4;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) {
        unsigned total = gridDim.x * gridDim.y * gridDim.z;
        last = (atomicAdd(A.done, 1u) == total - 1);
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *A.done = 0u;
        __threadfence_system();
        if ((int)threadIdx.x < A.sig.n && A.sig.flag[threadIdx.x]) *reinterpret_cast<volatile unsigned*>(A.sig.flag[threadIdx.x]) = A.gen;
        __threadfence_system();
    }
}

// Bounded spin until every rank's word in MY flag page reached `gen`.  On timeout (limit_ns of wall time on the
// device's global timer; HG_HALO_TIMEOUT_S, default 60 s) it counts an error and raises the context's STICKY error
// word in mapped host memory instead of hanging the GPU: the host sees it without a synchronisation and every
// later hg_run / hg_dispatch_grid / hg_sync on the context fails with HG_ERR_STATE (the ghost rows are stale).
__global__ void k_halo_wait(FlagArgs F, unsigned gen, unsigned long long limit_ns, unsigned long long* err, volatile unsigned* sticky) {
    if ((int)threadIdx.x >= F.n || !F.flag[threadIdx.x]) return;
    const volatile unsigned* f = reinterpret_cast<const volatile unsigned*>(F.flag[threadIdx.x]);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*f - gen) < 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > limit_ns) {
            atomicAdd(err, 1ull);
            if (sticky) { *sticky = gen; __threadfence_system(); }
            return;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

// generation signal without a push (hg_slab_barrier(c, false))
__global__ void k_halo_signal(FlagArgs sig, unsigned gen) {
    __threadfence_system();
    if ((int)threadIdx.x < sig.n && sig.flag[threadIdx.x]) *reinterpret_cast<volatile unsigned*>(sig.flag[threadIdx.x]) = gen;
    __threadfence_system();
}

// flag word written by rank `from`, inside the flag page that follows the planes of `arena`
inline unsigned* flag_ptr(float* arena, size_t plane_elems, int from) {
    return reinterpret_cast<unsigned*>(arena + (size_t)2 * HG_NPLANES * plane_elems) + from * 32;   // 128 B apart
}
inline size_t plane_elems_of(const hg_ctx* c, int k) { return (size_t)(c->slabs.rows[k] + 2 * HG_HALO_ROWS) * c->g.pitch; }

}  // namespace

static int preload_step_kernels(hg_ctx* c) {
    cudaFuncAttributes a;
    HG_CUDA(cudaFuncGetAttributes(&a, k_halo_push_signal));
    HG_CUDA(cudaFuncGetAttributes(&a, k_halo_wait));
    HG_CUDA(cudaFuncGetAttributes(&a, k_halo_signal));
    int rc = hg_preload_fused_kernels();
    if (rc == HG_OK) rc = hg_preload_init_rain_kernels();
    if (rc == HG_OK) rc = hg_preload_context_kernels();
    if (rc == HG_OK && c->erosion_type == HG_PARTICLES) rc = hg_preload_particle_kernels();
    return rc;
}

// The sticky error word of k_halo_wait: mapped pinned host memory, so the host reads it without synchronising.
static int ensure_sticky(hg_ctx* c) {
    int rcp = preload_step_kernels(c);      // every connect: cheap once the kernels are loaded
    if (rcp) return rcp;
    if (c->h_sticky) return HG_OK;
    HG_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->h_sticky), sizeof(unsigned), cudaHostAllocMapped));
    *c->h_sticky = 0u;
    HG_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_sticky), c->h_sticky, 0));
    double s = 60.0;
    if (const char* e = getenv("HG_HALO_TIMEOUT_S")) { double v = atof(e); if (v > 0.0) s = v; }
    c->halo_timeout_ns = (unsigned long long)(s * 1e9);
    return HG_OK;
}

extern "C" int hg_slab_export_handle(hg_ctx* c, hg_slab_export* out) {
    HG_CHECK_CTX(c);
    if (!out) return HG_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    cudaIpcMemHandle_t h;
    HG_CUDA(cudaIpcGetMemHandle(&h, c->arena));
    static_assert(sizeof(h) <= HG_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(out->mem_handle, &h, sizeof(h));
    out->arena_bytes = c->arena_bytes;
    out->row0 = (uint32_t)c->g.row0; out->rows = (uint32_t)c->g.rows;
    out->map_w = (uint32_t)c->g.W; out->map_h = (uint32_t)c->g.H;
    out->device = c->device;
    return HG_OK;
}

static int check_layout(hg_ctx* c, int n, int me, const int* row0, const int* rows, const int* w, const int* h) {
    if (n < 1 || n > HG_MAX_SLABS || me < 0 || me >= n) { hg_set_error("bad slab table (n=%d, me=%d)", n, me); return HG_ERR_INVALID; }
    int next = 0;
    for (int k = 0; k < n; k++) {
        if (w[k] != c->g.W || h[k] != c->g.H || row0[k] != next || rows[k] < HG_HALO_ROWS) {
            hg_set_error("slab %d (rows [%d,%d) of %dx%d) does not tile a %dx%d map in order, or is thinner than the halo (%d rows)",
                         k, row0[k], row0[k] + rows[k], w[k], h[k], c->g.W, c->g.H, HG_HALO_ROWS);
            return HG_ERR_INVALID;
        }
        next += rows[k];
    }
    if (next != c->g.H || row0[me] != c->g.row0 || rows[me] != c->g.rows) { hg_set_error("slab table does not cover the map or entry %d is not this context", me); return HG_ERR_INVALID; }
    return HG_OK;
}

extern "C" int hg_slab_connect(hg_ctx* c, const hg_slab_export* all, int n, int me) {
    HG_CHECK_CTX(c);
    if (!all) return HG_ERR_INVALID;
    int row0[HG_MAX_SLABS], rows[HG_MAX_SLABS], w[HG_MAX_SLABS], h[HG_MAX_SLABS];
    for (int k = 0; k < n && k < HG_MAX_SLABS; k++) { row0[k] = (int)all[k].row0; rows[k] = (int)all[k].rows; w[k] = (int)all[k].map_w; h[k] = (int)all[k].map_h; }
    int rc = check_layout(c, n, me, row0, rows, w, h);
    if (rc) return rc;
    hg_slab_disconnect(c);
    for (int k = 0; k < n; k++) {
        c->slabs.row0[k] = row0[k]; c->slabs.rows[k] = rows[k];
        if (k == me) { c->slabs.arena[k] = c->arena; c->slab_ipc[k] = false; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, all[k].mem_handle, sizeof(hd));
        void* p = nullptr;
        HG_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        c->slabs.arena[k] = static_cast<float*>(p);
        c->slab_ipc[k] = true;
    }
    c->slabs.n = n; c->slabs.me = me;
    c->peers_connected = n > 1;
    return ensure_sticky(c);
}

extern "C" int hg_slab_connect_local(hg_ctx* c, hg_ctx* const* all, int n, int me) {
    HG_CHECK_CTX(c);
    if (!all) return HG_ERR_INVALID;
    int row0[HG_MAX_SLABS], rows[HG_MAX_SLABS], w[HG_MAX_SLABS], h[HG_MAX_SLABS];
    for (int k = 0; k < n && k < HG_MAX_SLABS; k++) {
        if (!all[k]) { hg_set_error("null slab %d", k); return HG_ERR_INVALID; }
        row0[k] = all[k]->g.row0; rows[k] = all[k]->g.rows; w[k] = all[k]->g.W; h[k] = all[k]->g.H;
    }
    int rc = check_layout(c, n, me, row0, rows, w, h);
    if (rc) return rc;
    if (all[me] != c) { hg_set_error("entry %d is not this context", me); return HG_ERR_INVALID; }
    hg_slab_disconnect(c);
    for (int k = 0; k < n; k++) {
        if (all[k]->device != c->device) {
            int can = 0;
            HG_CUDA(cudaDeviceCanAccessPeer(&can, c->device, all[k]->device));
            if (!can) { hg_set_error("device %d cannot access device %d", c->device, all[k]->device); return HG_ERR_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(all[k]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { hg_set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return HG_ERR_CUDA; }
            cudaGetLastError();
        }
        c->slabs.arena[k] = all[k]->arena;
        c->slabs.row0[k] = row0[k]; c->slabs.rows[k] = rows[k];
        c->slab_ipc[k] = false;
        if (c->erosion_type == HG_PARTICLES) {
            if (all[k]->erosion_type != HG_PARTICLES || all[k]->particle_count != c->particle_count || !all[k]->pa || !all[k]->p_own) {
                hg_set_error("slab %d is not a droplet slab with %u droplets", k, c->particle_count);
                return HG_ERR_INVALID;
            }
            c->peer_pa[k] = all[k]->pa; c->peer_parts[k] = all[k]->particles; c->peer_own[k] = all[k]->p_own;
            c->peer_p_ipc[k] = false;
        }
    }
    c->slabs.n = n; c->slabs.me = me;
    c->peers_connected = n > 1;
    if (c->erosion_type == HG_PARTICLES) {
        int rco = hg_particle_own_init(c);
        if (rco) return rco;
    }
    return ensure_sticky(c);
}

// The droplet mode's three extra allocations (texel images, droplet array, ownership bytes) between processes.
extern "C" int hg_slab_export_particles(hg_ctx* c, hg_slab_export_particles_t* out) {
    HG_CHECK_CTX(c);
    if (!out) return HG_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    if (c->erosion_type != HG_PARTICLES || !c->pa || !c->p_own) { hg_set_error("not a droplet slab"); return HG_ERR_STATE; }
    cudaIpcMemHandle_t h;
    HG_CUDA(cudaIpcGetMemHandle(&h, c->pa)); memcpy(out->images_handle, &h, sizeof(h));
    HG_CUDA(cudaIpcGetMemHandle(&h, c->particles)); memcpy(out->droplets_handle, &h, sizeof(h));
    HG_CUDA(cudaIpcGetMemHandle(&h, c->p_own)); memcpy(out->owners_handle, &h, sizeof(h));
    out->particle_count = c->particle_count;
    return HG_OK;
}
// after hg_slab_connect: all[0..n) in the same order
extern "C" int hg_slab_connect_particles(hg_ctx* c, const hg_slab_export_particles_t* all, int n, int me) {
    HG_CHECK_CTX(c);
    if (!all || n != c->slabs.n || me != c->slabs.me || c->erosion_type != HG_PARTICLES) { hg_set_error("hg_slab_connect_particles: call hg_slab_connect first, with the same table order"); return HG_ERR_STATE; }
    for (int k = 0; k < n; k++) {
        if (all[k].particle_count != c->particle_count) { hg_set_error("slab %d holds %u droplets, this one %u", k, all[k].particle_count, c->particle_count); return HG_ERR_INVALID; }
        if (k == me) { c->peer_pa[k] = c->pa; c->peer_parts[k] = c->particles; c->peer_own[k] = c->p_own; c->peer_p_ipc[k] = false; continue; }
        cudaIpcMemHandle_t h;
        void* p = nullptr;
        memcpy(&h, all[k].images_handle, sizeof(h)); HG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); c->peer_pa[k] = static_cast<float4*>(p);
        memcpy(&h, all[k].droplets_handle, sizeof(h)); HG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); c->peer_parts[k] = static_cast<hg_particle*>(p);
        memcpy(&h, all[k].owners_handle, sizeof(h)); HG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); c->peer_own[k] = static_cast<unsigned char*>(p);
        c->peer_p_ipc[k] = true;
    }
    return hg_particle_own_init(c);
}

void hg_slab_disconnect(hg_ctx* c) {
    for (int k = 0; k < c->slabs.n; k++)
        if (c->slab_ipc[k] && c->slabs.arena[k]) cudaIpcCloseMemHandle(c->slabs.arena[k]);
    for (int k = 0; k < HG_MAX_SLABS; k++)
        if (c->peer_p_ipc[k]) {
            if (c->peer_pa[k]) cudaIpcCloseMemHandle(c->peer_pa[k]);
            if (c->peer_parts[k]) cudaIpcCloseMemHandle(c->peer_parts[k]);
            if (c->peer_own[k]) cudaIpcCloseMemHandle(c->peer_own[k]);
        }
    memset(c->peer_pa, 0, sizeof(c->peer_pa)); memset(c->peer_parts, 0, sizeof(c->peer_parts));
    memset(c->peer_own, 0, sizeof(c->peer_own)); memset(c->peer_p_ipc, 0, sizeof(c->peer_p_ipc));
    memset(&c->slabs, 0, sizeof(c->slabs));
    memset(c->slab_ipc, 0, sizeof(c->slab_ipc));
    c->peers_connected = false;
    c->pending_gen = 0;
}

extern "C" int hg_slab_errors(hg_ctx* c, uint64_t* count) {
    HG_CHECK_CTX(c);
    unsigned long long v = 0;
    HG_CUDA(cudaMemcpyAsync(&v, c->d_counters + 1, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    HG_CUDA(cudaStreamSynchronize(c->stream));
    if (count) *count = v;
    return HG_OK;
}

int hg_slab_check_sticky(hg_ctx* c) {
    if (c->h_sticky && *reinterpret_cast<volatile unsigned*>(c->h_sticky) != 0u) {
        hg_set_error("a halo wait timed out at exchange generation %u (a rank fell more than %.0f s behind or died): the ghost rows of this slab are stale",
                     *c->h_sticky, (double)c->halo_timeout_ns * 1e-9);
        return HG_ERR_STATE;
    }
    return HG_OK;
}

// After a fused step.  When the step's own kernels already stored the edge rows into the neighbours' ghost rows and
// signalled the generation (hg_fused.cu: fused push), only the bookkeeping is left.
int hg_slab_exchange(hg_ctx* c) {
    if (c->fused_push_gen) {
        c->pending_gen = c->fused_push_gen;
        c->fused_push_gen = 0;
        return HG_OK;
    }
    return hg_slab_barrier(c, true);
}

int hg_slab_peer_planes(const hg_ctx* c, float* out[2][HG_NPLANES], int* mask) {
    const HgSlabTable& T = c->slabs;
    *mask = 0;
    for (int s = 0; s < 2; s++) {
        const int nb = T.me + (s == 0 ? -1 : 1);
        for (int p = 0; p < HG_NPLANES; p++) out[s][p] = nullptr;
        if (nb < 0 || nb >= T.n) continue;
        *mask |= 1 << s;
        const size_t nb_elems = plane_elems_of(c, nb);
        // my local row r (ghost rows included) of the first owned rows is the lower neighbour's local row r + rows[nb];
        // of the last owned rows (local rows [rows, rows + HALO)) the upper neighbour's local row r - rows
        const ptrdiff_t shift = s == 0 ? (ptrdiff_t)T.rows[nb] * c->g.pitch : -(ptrdiff_t)c->g.rows * c->g.pitch;
        for (int p = 0; p < HG_NPLANES; p++) {
            const int dst_set = c->ri[hg_field_of_plane(p)] ^ 1;      // a fused step writes the other set of every field
            out[s][p] = T.arena[nb] + ((size_t)dst_set * HG_NPLANES + p) * nb_elems + shift;
        }
    }
    return HG_OK;
}
void hg_slab_signal_args(const hg_ctx* c, HgFlagArgs* out) {
    const HgSlabTable& T = c->slabs;
    memset(out, 0, sizeof(*out));
    out->n = T.n;
    for (int k = 0; k < T.n; k++)
        if (k != T.me) out->flag[k] = flag_ptr(T.arena[k], plane_elems_of(c, k), T.me);
}
static int slab_generation(hg_ctx* c, int mode);
int hg_slab_barrier(hg_ctx* c, bool push) { return slab_generation(c, push ? 1 : 0); }
int hg_slab_push_images(hg_ctx* c) { return slab_generation(c, 2); }

// After a fused step (push = true): push my new edge rows to both neighbours and publish the generation on
// every rank; the wait until every rank has published it is enqueued in front of the next step (hg_slab_wait_pending).  push = false: only the
// generation signal + wait, an all-rank barrier on the device (after an in-place rain: a peer's far fetch of
// the next step must not read this rank's water before the rain has been added).
// mode 0: signal only; 1: push the nine planes; 2: push the droplet mode's heightmap and momentum images
static int slab_generation(hg_ctx* c, int mode) {
    const bool push = mode != 0;
    if (!c->peers_connected) return HG_OK;
    int rcw = hg_slab_wait_pending(c);      // generations are waited for in order
    if (rcw) return rcw;
    const HgSlabTable& T = c->slabs;
    c->step_flag++;
    const unsigned gen = c->step_flag;
    const size_t rowsz = (size_t)c->g.pitch;
    PushArgs A;
    memset(&A, 0, sizeof(A));
    // elements per row of a "plane": 1 float per cell, or 4 for a texel image (droplet mode: H image, M image)
    const size_t per = mode == 2 ? 4 : 1;
    if (mode == 2) {
        A.src[0] = reinterpret_cast<const float*>(hg_pa_h(c, 1));
        A.src[1] = reinterpret_cast<const float*>(hg_pa_m(c, 1));
    } else {
        for (int p = 0; p < HG_NPLANES; p++) A.src[p] = hg_plane(c, c->ri[hg_field_of_plane(p)], p);
    }
    A.n = (size_t)HG_HALO_ROWS * rowsz * per;
    for (int s = 0; push && s < 2; s++) {
        int nb = T.me + (s == 0 ? -1 : 1);
        if (nb < 0 || nb >= T.n) continue;
        size_t nb_elems = plane_elems_of(c, nb);
        if (mode == 2) {
            A.side[s].dst[0] = reinterpret_cast<float*>(c->peer_pa[nb] + (size_t)c->ri[0] * nb_elems);
            A.side[s].dst[1] = reinterpret_cast<float*>(c->peer_pa[nb] + (size_t)(2 + c->ri[2]) * nb_elems);
        } else {
            for (int p = 0; p < HG_NPLANES; p++)
                A.side[s].dst[p] = T.arena[nb] + ((size_t)c->ri[hg_field_of_plane(p)] * HG_NPLANES + p) * nb_elems;
        }
        if (s == 0) {   // my first owned rows -> lower neighbour's upper ghost rows
            A.side[s].src_off = (size_t)HG_HALO_ROWS * rowsz * per;
            A.side[s].dst_off = (size_t)(HG_HALO_ROWS + T.rows[nb]) * rowsz * per;
        } else {        // my last owned rows -> upper neighbour's lower ghost rows
            A.side[s].src_off = (size_t)c->g.rows * rowsz * per;
            A.side[s].dst_off = 0;
        }
    }
    A.sig.n = T.n;
    for (int k = 0; k < T.n; k++)
        if (k != T.me) A.sig.flag[k] = flag_ptr(T.arena[k], plane_elems_of(c, k), T.me);   // my word on rank k
    A.gen = gen;
    A.done = reinterpret_cast<unsigned*>(c->d_counters + 10);
    if (push) {
        size_t blocks = (A.n / 4 + 255) / 256;
        dim3 grid((unsigned)(blocks < 16 ? blocks : 16), HG_NPLANES, 2);
        k_halo_push_signal<<<grid, 256, 0, c->stream>>>(A);
    } else {
        k_halo_signal<<<1, 32, 0, c->stream>>>(A.sig, gen);
    }
    HG_LAUNCH_CHECK(c);
    // The wait for the other ranks' signals is NOT enqueued here but in front of the next kernel that touches the
    // planes (hg_slab_wait_pending, called by the rain and step launchers): nothing between two steps needs it, the
    // other ranks get the length of that gap to catch up, and a host thread that drives several slabs of one process
    // may issue blocking calls (allocations, checkpoint I/O) for slab B while slab A has signalled -- with the spin
    // kernel already on A's stream those calls would wait for a kernel that waits for B.
    c->pending_gen = gen;
    return HG_OK;
}

int hg_slab_wait_pending(hg_ctx* c) {
    if (!c->peers_connected || !c->pending_gen) return HG_OK;
    const HgSlabTable& T = c->slabs;
    FlagArgs Wt{};
    Wt.n = T.n;
    for (int k = 0; k < T.n; k++)
        if (k != T.me) Wt.flag[k] = flag_ptr(c->arena, c->g.plane_elems, k);      // rank k's word on me
    k_halo_wait<<<1, 32, 0, c->stream>>>(Wt, c->pending_gen, c->halo_timeout_ns, c->d_counters + 1, c->d_sticky);
    HG_LAUNCH_CHECK(c);
    c->pending_gen = 0;
    return HG_OK;
}

// Ghost rows after the owned rows were replaced from outside (hg_upload, hg_checkpoint_load): every rank pushes
// its edge rows to its neighbours and waits for theirs.  Collective: every rank of the slab table must call it
// (hg_checkpoint_load does; after plain uploads the caller does), in the same order relative to its steps.
extern "C" int hg_slab_refresh_halo(hg_ctx* c) {
    HG_CHECK_CTX(c);
    return hg_slab_barrier(c, true);
}
