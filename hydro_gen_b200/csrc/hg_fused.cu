// hg_fused.cu — the product path: one grid erosion step (Erosion::dispatch_grid,
// src/erosion.cpp:158-200: flux, erosion, sediment transport, thermal x2 layers,
// smoothing) as ONE kernel that reads the nine persistent planes once and writes them
// once (72 B/cell-step) instead of the reference's 8 dispatches and
// >= 512 B/cell-step.  The per-thread row iteration lives in hg_fused_body.cuh (shared
// with the CPU emulation in tests/host_emul); this file holds the kernel shells, the
// far-fetch fix-up kernel and the launcher.
//
// The back-trace is unbounded in the reference (sediment_transport.glsl:27-28).  When
// its 2x2 footprint lies within +-1 cell (> 99.9 % of cells in a normal run) S' is read
// from the ring.  The other cells are appended to a list and resolved after the main
// kernel by k_far_fixup, which recomputes u,v and S' at the four texels from the PRE-step
// planes (still intact: a step writes the other plane set; on another GPU's slab they are
// read through its peer pointer).  That gives exactly the reference's result without
// materialising S' and without stalling a CTA on a scattered gather.
//
// All arithmetic is the shared per-cell code of hg_cell.cuh: results are bit-identical to
// the PASSES schedule and to the CPU oracle.
//
// Kernels in this file: k_fused_ws (the default: hydraulic and thermal stages on two warp groups of
// one CTA, three CTAs per SM), k_fused_step (the single-group form, a tuning variant), k_far_fixup,
// and k_plan_segments (re-cuts the strips into row segments of equal forecast cost for the next
// step, from the durations the CTAs of this step reported).
#include <cuda.h>
#include <vector>
#include "hg_internal.cuh"
#include "hg_fused_body.cuh"
#include "hg_fused_body2.cuh"
#include "hg_fused_body3.cuh"
#include "hg_plan.cuh"

#ifndef HG_FREE_UNROLL
#define HG_FREE_UNROLL 1
#endif
constexpr int kFreeUnroll = HG_FREE_UNROLL;
// steady-state row loop of k_fused_ws per warp group (measured: anything above 1 leaves the instruction cache)
#ifndef HG_UNROLL_H
#define HG_UNROLL_H 1
#endif
#ifndef HG_UNROLL_T
#define HG_UNROLL_T 1
#endif
constexpr int kUnrollH = HG_UNROLL_H, kUnrollT = HG_UNROLL_T;
constexpr int HG_MAX_DEVICES = 64;
#ifndef HG_FUSED_DEFAULT_VARIANT
#define HG_FUSED_DEFAULT_VARIANT 5     // 5: one column per thread (k_fused_ws); 10: two columns per thread (k_fused_ws2)
#endif

namespace {

struct FusedArgs {
    const float* src[HG_NPLANES];
    float* dst[HG_NPLANES];
    int W, H, row0, rows, pitch;
    int seg, nstrips;
    int src_set[HG_NPLANES];             // plane set index of src (for peers)
    HgSlabTable slabs;                   // n >= 1; entry `me` is this slab
    unsigned* far_list;                  // local linear cell indices (row - row0) * W + x
    unsigned long long* far_count;       // this step's counter
    unsigned long long* far_count_next;  // zeroed by the fix-up kernel for the next step
    unsigned long long* far_total;       // statistics
    // connected slab with the push fused into the step (HgFusedK::peer): the fix-up repeats its stores of edge-row cells
    // into the neighbours' ghost rows, and its LAST block publishes the exchange generation on every rank (hg_slab.cu)
    float* peer[2][2];                   // [side][sediment rock, dirt], pre-offset like HgFusedK::peer
    int push_mask, push_lo_end, push_hi_begin;
    HgFlagArgs sig;                      // my flag word on every other rank; n == 0: no signal from this kernel
    unsigned gen;
    unsigned* done;                      // block counter (self-resetting)
    HgStepParams P;
};

// ------------------------------------------------------------------ far-fetch path
// Row gy of the pre-step plane set, wherever it lives: pointer to (plane 0, row gy, column 0)
// and the plane stride of that slab.  The own slab (ghost rows included) is tried first; a row
// outside it is read through the owning rank's peer pointer.  gy must be inside the map.
struct FarRow { const float* p; size_t pe; };
__device__ __forceinline__ FarRow far_row(const FusedArgs& A, int gy) {
    const HgSlabTable& T = A.slabs;
    int k = T.me;
    if (gy < T.row0[k] - HG_HALO_ROWS || gy >= T.row0[k] + T.rows[k] + HG_HALO_ROWS) {
        for (int j = 0; j < T.n; j++)
            if (gy >= T.row0[j] && gy < T.row0[j] + T.rows[j]) { k = j; break; }
    }
    FarRow r;
    r.pe = (size_t)(T.rows[k] + 2 * HG_HALO_ROWS) * A.pitch;
    r.p = T.arena[k] + (size_t)A.src_set[0] * HG_NPLANES * r.pe + (size_t)(gy - T.row0[k] + HG_HALO_ROWS) * A.pitch;
    return r;
}
// Stage A of any cell recomputed from the pre-step planes: S' and the velocity.  An out-of-map
// texel is texelFetch's 0.  All 25 loads are issued from clamped (always valid) addresses before
// any of them is used, and out-of-map neighbours are replaced afterwards, so one evaluation
// costs one memory round trip.  (All nine read planes sit in the same ping-pong set: align_sets.)
__device__ __forceinline__ void far_stage_a(const FusedArgs& A, int x, int gy, float* sr, float* sd, float* u, float* v) {
    const bool inmap = !(x < 0 || x > A.W - 1 || gy < 0 || gy > A.H - 1);
    const int cx = min(max(x, 0), A.W - 1), cy = min(max(gy, 0), A.H - 1);
    const bool hasL = cx > 0, hasR = cx < A.W - 1, hasB = cy > 0, hasT = cy < A.H - 1;
    const int xl = hasL ? cx - 1 : cx, xr = hasR ? cx + 1 : cx;
    const FarRow r0 = far_row(A, cy), rt = far_row(A, hasT ? cy + 1 : cy), rb = far_row(A, hasB ? cy - 1 : cy);
#define FL(row, plane, col) __ldcg((row).p + (size_t)(plane) * (row).pe + (col))
    const float rk0 = FL(r0, PL_ROCK, cx), dt0 = FL(r0, PL_DIRT, cx), w0 = FL(r0, PL_WATER, cx);
    float rkL = FL(r0, PL_ROCK, xl), dtL = FL(r0, PL_DIRT, xl), wL = FL(r0, PL_WATER, xl);
    float rkR = FL(r0, PL_ROCK, xr), dtR = FL(r0, PL_DIRT, xr), wR = FL(r0, PL_WATER, xr);
    float rkT = FL(rt, PL_ROCK, cx), dtT = FL(rt, PL_DIRT, cx), wT = FL(rt, PL_WATER, cx);
    float rkB = FL(rb, PL_ROCK, cx), dtB = FL(rb, PL_DIRT, cx), wB = FL(rb, PL_WATER, cx);
    const float fL = FL(r0, PL_FL, cx), fR = FL(r0, PL_FR, cx), fT = FL(r0, PL_FT, cx), fB = FL(r0, PL_FB, cx);
    float inL = FL(r0, PL_FR, xl), inR = FL(r0, PL_FL, xr), inT = FL(rt, PL_FB, cx), inB = FL(rb, PL_FT, cx);
    const float s0 = FL(r0, PL_SR, cx), s1 = FL(r0, PL_SD, cx);
#undef FL
    // H.a of the neighbours in the shader's association (r + g) + b; imageLoad outside the map: 0 / OOB height
    float aL = rkL + dtL + wL, aR = rkR + dtR + wR, aT = rkT + dtT + wT, aB = rkB + dtB + wB;
    if (!hasL) { rkL = 0.0f; dtL = 0.0f; aL = HG_OOB_HEIGHT; inL = 0.0f; }
    if (!hasR) { rkR = 0.0f; dtR = 0.0f; aR = HG_OOB_HEIGHT; inR = 0.0f; }
    if (!hasT) { rkT = 0.0f; dtT = 0.0f; aT = HG_OOB_HEIGHT; inT = 0.0f; }
    if (!hasB) { rkB = 0.0f; dtB = 0.0f; aB = HG_OOB_HEIGHT; inB = 0.0f; }
    HgFluxOut o = hg_flux_cell(A.P, cx, cy, A.W, A.H, rk0 + dt0 + w0, aL, aR, aT, aB, fL, fR, fT, fB, inL, inR, inT, inB, w0);
    HgEroOut e = hg_erosion_cell(A.P, rk0, dt0, s0, s1, o.u, o.v, o.vz, rkR, dtR, rkL, dtL, rkB, dtB, rkT, dtT);
    *sr = inmap ? e.sr : 0.0f; *sd = inmap ? e.sd : 0.0f; *u = inmap ? o.u : 0.0f; *v = inmap ? o.v : 0.0f;
}

// Four lanes per listed cell: the sediment pass of sediment_transport.glsl:66-93 with every
// texel recomputed from pre-step state.  All four lanes evaluate the cell itself (same
// addresses, one broadcast load each) to get its velocity and back-traced position, then lane q
// evaluates texel (px + (q & 1), py + (q >> 1)) and lane 0 gathers the four by shuffle: two
// memory round trips per cell instead of five dependent evaluations in one thread.
__global__ void __launch_bounds__(128) k_far_fixup(const __grid_constant__ FusedArgs A) {
    const unsigned long long n = *A.far_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *A.far_count_next = 0ull;
        if (n) atomicAdd(A.far_total, n);
    }
    const int q = threadIdx.x & 3;
    const unsigned long long stride = (unsigned long long)gridDim.x * (blockDim.x / 4);
    const unsigned long long n_up = (n + 7ull) / 8ull * 8ull;      // whole warps stay in the loop together (shuffles)
    for (unsigned long long e = blockIdx.x * (unsigned long long)(blockDim.x / 4) + threadIdx.x / 4; e < n_up; e += stride) {
        const bool live = e < n;
        unsigned li = live ? A.far_list[e] : 0u;
        int ly = (int)(li / (unsigned)A.W), x = (int)(li - (unsigned)ly * (unsigned)A.W);
        int gy = A.row0 + ly;
        float s0, s1, u, v;
        far_stage_a(A, x, gy, &s0, &s1, &u, &v);
        HgBack b = hg_backtrace(A.P, x, gy, A.W, A.H, u, v);
        float tr, td, du, dv;
        far_stage_a(A, b.px + (q & 1), b.py + (q >> 1), &tr, &td, &du, &dv);
        const int base = (threadIdx.x & 31) & ~3;
        const float t00r = __shfl_sync(0xffffffffu, tr, base), t10r = __shfl_sync(0xffffffffu, tr, base + 1);
        const float t01r = __shfl_sync(0xffffffffu, tr, base + 2), t11r = __shfl_sync(0xffffffffu, tr, base + 3);
        const float t00d = __shfl_sync(0xffffffffu, td, base), t10d = __shfl_sync(0xffffffffu, td, base + 1);
        const float t01d = __shfl_sync(0xffffffffu, td, base + 2), t11d = __shfl_sync(0xffffffffu, td, base + 3);
        if (live && q == 0) {
            size_t idx = (size_t)(ly + HG_HALO_ROWS) * A.pitch + x;
            const float vr = hg_bilerp(t00r, t10r, t01r, t11r, b.sx, b.sy), vd = hg_bilerp(t00d, t10d, t01d, t11d, b.sx, b.sy);
            A.dst[PL_SR][idx] = vr;
            A.dst[PL_SD][idx] = vd;
            if ((A.push_mask & 1) && gy < A.push_lo_end) { A.peer[0][0][idx] = vr; A.peer[0][1][idx] = vd; }
            if ((A.push_mask & 2) && gy >= A.push_hi_begin) { A.peer[1][0][idx] = vr; A.peer[1][1][idx] = vd; }
        }
    }
    if (A.sig.n) {
        // Every store of this step into a neighbour's ghost rows -- the step kernel's (complete: it ran before this kernel
        // on the same stream) and this kernel's -- is ordered before the flag words by the fence + counter + fence of
        // the last block, as in k_halo_push_signal.
        __threadfence_system();
        __syncthreads();
        __shared__ unsigned last;
        if (threadIdx.x == 0) last = (atomicAdd(A.done, 1u) == gridDim.x - 1);
        __syncthreads();
        if (last) {
            if (threadIdx.x == 0) *A.done = 0u;
            __threadfence_system();
            if ((int)threadIdx.x < A.sig.n && A.sig.flag[threadIdx.x]) *reinterpret_cast<volatile unsigned*>(A.sig.flag[threadIdx.x]) = A.gen;
            __threadfence_system();
        }
    }
}

// ------------------------------------------------------------------ main kernel
// TMA / mbarrier primitives (PTX ISA 8.x, sm_90+): one elected thread arms an mbarrier with
// the byte count of a box and issues cp.async.bulk.tensor; every thread waits on the phase.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HG_DONE_%=;\n"
        "bra HG_WAIT_%=;\n"
        "HG_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// The box start column must be a multiple of 4 floats: TMA faults ("illegal instruction") on a
// start address that is not 16-byte aligned, also with interleave and swizzle off.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// The same with shared-window addresses formed once outside the row loop (the generic-to-shared conversion of a pointer
// costs a special-register read and two uniform operations every time it is repeated).
__device__ __forceinline__ void mbar_expect_tx_a(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HG_WAITA_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HG_DONEA_%=;\n"
        "bra HG_WAITA_%=;\n"
        "HG_DONEA_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(unsigned dst, const CUtensorMap* map, unsigned bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

// Shared block: [raw ring: 2 slots x 9 planes x (NT+4) floats, each 128-byte aligned][row rings][2 mbarriers]
template <int NT> struct FusedSmem {
    static constexpr size_t RAW_BOX = (size_t)HGF_NPL * HGF_RAW_LD(NT);             // floats delivered per box
    static constexpr size_t RAW_SLOT = (RAW_BOX + 31) / 32 * 32;                    // slot stride, floats
    static constexpr size_t RINGS = 2 * RAW_SLOT;                                   // float offset of the row rings
    static constexpr size_t BARS = (RINGS + HgRings<NT>::TOTAL + 3) / 4 * 4;        // float offset of the mbarriers (16-byte aligned)
    static constexpr size_t BYTES = (BARS + 4) * sizeof(float);
};

template <int NT, int MINB, bool DROPS = false>
__global__ void __launch_bounds__(NT, MINB) k_fused_step(const __grid_constant__ HgFusedK K, const __grid_constant__ CUtensorMap tmap) {
    // Dynamic shared memory is the kernel's only shared allocation, so it starts at the CTA's window
    // base (1 KiB aligned); the TMA destination needs 128 bytes.  (Rounding the pointer up at run
    // time makes the compiler lose the address space and emit generic LD/ST instead of LDS/STS.)
    extern __shared__ __align__(128) float smb[];
    float* const sm = smb + FusedSmem<NT>::RINGS;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smb + FusedSmem<NT>::BARS);
    const int tid = threadIdx.x;
    const int strip = blockIdx.x % K.nstrips, segi = blockIdx.x / K.nstrips;
    const int x0 = strip * (NT - 2 * HGF_HX) - HGF_HX;
    const int x = x0 + tid;
    const bool xin = x >= 0 && x < K.W;
    const bool owned = tid >= HGF_HX && tid < NT - HGF_HX && x < K.W;
    const int gy0 = K.row0 + segi * K.seg;
    const int gy1 = min(gy0 + K.seg, K.row0 + K.rows);
    const HgFusedPlan pl = hg_fused_plan(gy0, gy1, K.H);
    const unsigned pitch = (unsigned)K.pitch;
    // element offset of (row i, column x) in a plane, advanced by one row per iteration
    unsigned off = (unsigned)(pl.i_begin - K.row0 + HG_HALO_ROWS) * pitch + (unsigned)x;
    const int ly0 = pl.i_begin - K.row0 + HG_HALO_ROWS;      // plane row of iteration i_begin
    constexpr unsigned BOX_BYTES = DROPS ? (unsigned)(HGF_RAW_LD(NT) * 4 * sizeof(float)) : (unsigned)(FusedSmem<NT>::RAW_BOX * sizeof(float));
    const int bx0 = x0 - 2;      // box start column: a multiple of 4 (TMA needs a 16-byte aligned start)
#define HG_TMA_ROW1(dst, bar, row) (DROPS ? tma_load_3d((dst), &tmap, (bar), 0, bx0, (row)) : tma_load_3d((dst), &tmap, (bar), bx0, (row), 0))
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[0], BOX_BYTES);
        HG_TMA_ROW1(smb, &bars[0], ly0);
    }
    __syncthreads();
    HgCol c;
    hg_fused_begin(c);
    int i = pl.i_begin;
    // iteration with relative index rel = i - i_begin: its raw row sits in slot rel & 1, phase (rel >> 1) & 1;
    // the row of the next iteration is requested first (its slot was last read before the previous barrier).
    // A deeper ring (3 slots, two rows ahead) was measured and is not faster.
#define HG_ROW(FREEFLAG)                                                                                             \
    {                                                                                                                \
        const int rel = i - pl.i_begin;                                                                              \
        if (tid == 0 && i < pl.i_end) {                                                                              \
            mbar_expect_tx(&bars[(rel + 1) & 1], BOX_BYTES);                                                         \
            HG_TMA_ROW1(smb + ((rel + 1) & 1) * FusedSmem<NT>::RAW_SLOT, &bars[(rel + 1) & 1], ly0 + rel + 1);         \
        }                                                                                                            \
        mbar_wait(&bars[rel & 1], (unsigned)(rel >> 1) & 1u);                                                        \
        hg_fused_iter<NT, FREEFLAG, HGF_ALL, DROPS>(c, sm, smb + (rel & 1) * FusedSmem<NT>::RAW_SLOT, K, tid, x, xin, owned, gy0, gy1, i, off); \
        __syncthreads();                                                                                             \
    }
    for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW(false)
#pragma unroll kFreeUnroll
    for (; i <= pl.free_hi; i++, off += pitch) HG_ROW(true)
    for (; i <= pl.i_end; i++, off += pitch) HG_ROW(false)
#undef HG_ROW
#undef HG_TMA_ROW1
}

// Warp-specialised form of the same step: a CTA of 2*NT threads, threads [0, NT) run the
// hydraulic stages (L, A, B) of the strip's NT columns and threads [NT, 2*NT) the thermal and
// smoothing stages (C..G) of the same columns, one row behind each other through the rings
// (hg_fused_body.cuh).  The two groups rebalance the CTA's registers with setmaxnreg (RH for a
// hydraulic thread, RT for a thermal one, (RH + RT) / 2 = the launch allocation), so three CTAs
// = 24 warps stay resident per SM where the single-group kernel holds 16: the step is bound by
// instruction issue and latency, not by HBM, and the extra warps are what hides the latency.
template <int RH> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RH)); }
template <int RT> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RT)); }
// The two groups meet at barrier 0 from different loops.  An inline-asm barrier is not a convergent
// operation for the compiler, which may leave lanes of a warp diverged in front of it (seen with
// compute-sanitizer synccheck after the guarded stores of stage B): reconverge the warp first, and use
// the form of the instruction that is defined for unaligned arrival (barrier.sync, not bar.sync =
// barrier.sync.aligned).
__device__ __forceinline__ void cta_barrier() {
    __syncwarp();
    asm volatile("barrier.sync 0;" ::: "memory");
}

// SWAP: the hydraulic group on the upper half of the CTA's warps (the SM's warp arbiter prefers the higher warp id).
template <int NT, int MINB, int RH, int RT, bool DROPS, int SWAP = 0>
__global__ void __launch_bounds__(2 * NT, MINB) k_fused_ws(const __grid_constant__ HgFusedK K, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smb[];
    float* const sm = smb + FusedSmem<NT>::RINGS;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smb + FusedSmem<NT>::BARS);
    // SWAP 2: the groups interleaved in pairs of warps (warps 0,1 | 4,5 hydraulic, 2,3 | 6,7 thermal), so that a scheduler
    // (warp id mod 4) only ever runs ONE of the two loops; needs RH == RT (setmaxnreg works on whole warpgroups)
    static_assert(SWAP != 2 || (RH == RT && NT % 64 == 0), "interleaved groups cannot rebalance registers");
    const bool hydro = SWAP == 2 ? ((threadIdx.x >> 6) & 1) == 0 : SWAP == 1 ? threadIdx.x >= NT : threadIdx.x < NT;
    const int tid = SWAP == 2 ? (int)(((threadIdx.x >> 7) << 6) + (threadIdx.x & 63)) : threadIdx.x < NT ? threadIdx.x : threadIdx.x - NT;
    int strip, gy0, gy1;
    if (K.plan) {           // balanced partition (hg_plan_*): this CTA's strip and rows come from the plan
        const HgPlanItem it = K.plan[blockIdx.x];
        strip = it.strip; gy0 = it.gy0; gy1 = it.gy1;
    } else {
        const int segi = blockIdx.x / K.nstrips;
        strip = blockIdx.x % K.nstrips;
        gy0 = K.row0 + segi * K.seg;
        gy1 = min(gy0 + K.seg, K.row0 + K.rows);
    }
    unsigned long long t_start = 0;
    if (K.cta_ns && hydro && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    const int x0 = strip * (NT - 2 * HGF_HX) - HGF_HX;
    const int x = x0 + tid;
    const bool xin = x >= 0 && x < K.W;
    const bool owned = tid >= HGF_HX && tid < NT - HGF_HX && x < K.W;
    const HgFusedPlan pl = hg_fused_plan_slab(gy0, gy1, K.H, K);
    const unsigned pitch = (unsigned)K.pitch;
    unsigned off = (unsigned)(pl.i_begin - K.row0 + HG_HALO_ROWS) * pitch + (unsigned)x;
    const int ly0 = pl.i_begin - K.row0 + HG_HALO_ROWS;
    // grid step: box = (NT+4 columns) x 1 row x 9 planes at (column, row, plane 0); droplet mode: box = 4 channels x
    // (NT+4 texels) x 1 row of the heightmap in texture layout at (channel 0, column, row)
    constexpr unsigned BOX_BYTES = DROPS ? (unsigned)(HGF_RAW_LD(NT) * 4 * sizeof(float)) : (unsigned)(FusedSmem<NT>::RAW_BOX * sizeof(float));
    const int bx0 = x0 - 2;
#define HG_TMA_ROW(dst, bar, row) (DROPS ? tma_load_3d((dst), &tmap, (bar), 0, bx0, (row)) : tma_load_3d((dst), &tmap, (bar), bx0, (row), 0))
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[0], BOX_BYTES);
        HG_TMA_ROW(smb, &bars[0], ly0);
    }
    __syncthreads();
    HgCol c;
    hg_fused_begin(c);
    int i = pl.i_begin;
#ifndef HG_NO_RING_ROTATE
    HgRingOff ro = hg_ring_off<NT>(i, tid);
#define HG_RO , &ro
#define HG_RO_NEXT hg_ring_off_next(ro);
#else
#define HG_RO
#define HG_RO_NEXT
#endif
    if (hydro) {
        if (RH < RT) reg_dec<RH>(); else if (RH > RT) reg_inc<RH>();      // droplet mode gives the hydraulic group the larger share
        unsigned bar_a = smem_u32(bars), raw_a = smem_u32(smb);      /* loop-invariant shared-window addresses, */
        asm volatile("" : "+r"(bar_a), "+r"(raw_a));                       /* opaque so that they are kept, not re-derived every row */
#define HG_ROW_H(FREEFLAG)                                                                                           \
    {                                                                                                                \
        const int rel = i - pl.i_begin;                                                                              \
        if (tid == 0 && i < pl.i_end) {                                                                              \
            const unsigned nb = bar_a + (((unsigned)rel + 1u) & 1u) * 8u;                                            \
            const unsigned nd = raw_a + (((unsigned)rel + 1u) & 1u) * (unsigned)(FusedSmem<NT>::RAW_SLOT * sizeof(float)); \
            mbar_expect_tx_a(nb, BOX_BYTES);                                                                         \
            if (DROPS) tma_load_3d_a(nd, &tmap, nb, 0, bx0, ly0 + rel + 1); else tma_load_3d_a(nd, &tmap, nb, bx0, ly0 + rel + 1, 0); \
        }                                                                                                            \
        mbar_wait_a(bar_a + ((unsigned)rel & 1u) * 8u, ((unsigned)rel >> 1) & 1u);                                   \
        hg_fused_iter<NT, FREEFLAG, HGF_HYDRO, DROPS>(c, sm, smb + (rel & 1) * FusedSmem<NT>::RAW_SLOT, K, tid, x, xin, owned, gy0, gy1, i, off HG_RO); \
        HG_RO_NEXT                                                                                                   \
        cta_barrier();                                                                                               \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#pragma unroll kUnrollH
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_H(true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#undef HG_ROW_H
        if (K.cta_ns && tid == 0) {       // the last barrier has passed: both groups are done with their rows
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            K.cta_ns[blockIdx.x] = (unsigned)(t_end - t_start);
        }
    } else {
        if (RH < RT) reg_inc<RT>(); else if (RH > RT) reg_dec<RT>();
#define HG_ROW_T(FREEFLAG)                                                                                           \
    {                                                                                                                \
        hg_fused_iter<NT, FREEFLAG, HGF_THERMAL, DROPS>(c, sm, smb, K, tid, x, xin, owned, gy0, gy1, i, off HG_RO);         \
        HG_RO_NEXT                                                                                                   \
        cta_barrier();                                                                                               \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_T(false)
#pragma unroll kUnrollT
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_T(true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_T(false)
#undef HG_ROW_T
#undef HG_TMA_ROW
    }
#undef HG_RO
#undef HG_RO_NEXT
}


// ------------------------------------------------------------------ three warp groups
// The row period of k_fused_ws is the latency of ONE thermal warp's row (~690 mostly dependent instructions between two
// CTA barriers), not the issue slots (75 % busy).  Thermal layer 1 reads layer 0's result only through the R1D ring, so
// the thermal stages split once more: a CTA of 3*NT threads, threads [0, NT) run L, A, B, threads [NT, 2*NT) run C, D
// (layer 0) and threads [2*NT, 3*NT) run E, F (layer 1); smoothing (G) goes to group SG.  Three CTAs = 36 warps per SM
// where k_fused_ws holds 24, each with a dependent chain half as long per row.  R0 / R1: registers of a layer-0 /
// layer-1 thread; RH + R0 + R1 <= 3 * the launch allocation.
template <int RL, int R> __device__ __forceinline__ void reg_set() {
    if (R < RL) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
    else if (R > RL) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
}
template <int NT, int MINB, int RH, int R0, int R1, bool DROPS, int SG>
__global__ void __launch_bounds__(3 * NT, MINB) k_fused_ws3(const __grid_constant__ HgFusedK K, const __grid_constant__ CUtensorMap tmap) {
    constexpr int RL = 65536 / (3 * NT * MINB) / 8 * 8;      // launch allocation per thread
    static_assert(RH + R0 + R1 <= 3 * RL && RH % 8 == 0 && R0 % 8 == 0 && R1 % 8 == 0, "register budget of the three warp groups");
    extern __shared__ __align__(128) float smb[];
    float* const sm = smb + FusedSmem<NT>::RINGS;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smb + FusedSmem<NT>::BARS);
    const int group = threadIdx.x / NT;       // warp-uniform
    const int tid = threadIdx.x - group * NT;
    int strip, gy0, gy1;
    if (K.plan) {
        const HgPlanItem it = K.plan[blockIdx.x];
        strip = it.strip; gy0 = it.gy0; gy1 = it.gy1;
    } else {
        const int segi = blockIdx.x / K.nstrips;
        strip = blockIdx.x % K.nstrips;
        gy0 = K.row0 + segi * K.seg;
        gy1 = min(gy0 + K.seg, K.row0 + K.rows);
    }
    unsigned long long t_start = 0;
    if (K.cta_ns && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    const int x0 = strip * (NT - 2 * HGF_HX) - HGF_HX;
    const int x = x0 + tid;
    const bool xin = x >= 0 && x < K.W;
    const bool owned = tid >= HGF_HX && tid < NT - HGF_HX && x < K.W;
    const HgFusedPlan pl = hg_fused_plan(gy0, gy1, K.H);
    const unsigned pitch = (unsigned)K.pitch;
    unsigned off = (unsigned)(pl.i_begin - K.row0 + HG_HALO_ROWS) * pitch + (unsigned)x;
    const int ly0 = pl.i_begin - K.row0 + HG_HALO_ROWS;
    constexpr unsigned BOX_BYTES = DROPS ? (unsigned)(HGF_RAW_LD(NT) * 4 * sizeof(float)) : (unsigned)(FusedSmem<NT>::RAW_BOX * sizeof(float));
    const int bx0 = x0 - 2;
#define HG_TMA_ROW(dst, bar, row) (DROPS ? tma_load_3d((dst), &tmap, (bar), 0, bx0, (row)) : tma_load_3d((dst), &tmap, (bar), bx0, (row), 0))
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[0], BOX_BYTES);
        HG_TMA_ROW(smb, &bars[0], ly0);
    }
    __syncthreads();
    HgCol c;
    hg_fused_begin(c);
    int i = pl.i_begin;
    if (group == 0) {
        reg_set<RL, RH>();
#define HG_ROW_H(FREEFLAG)                                                                                           \
    {                                                                                                                \
        const int rel = i - pl.i_begin;                                                                              \
        if (tid == 0 && i < pl.i_end) {                                                                              \
            mbar_expect_tx(&bars[(rel + 1) & 1], BOX_BYTES);                                                         \
            HG_TMA_ROW(smb + ((rel + 1) & 1) * FusedSmem<NT>::RAW_SLOT, &bars[(rel + 1) & 1], ly0 + rel + 1);       \
        }                                                                                                            \
        mbar_wait(&bars[rel & 1], (unsigned)(rel >> 1) & 1u);                                                        \
        hg_fused_iter<NT, FREEFLAG, HGF_HYDRO, DROPS, SG>(c, sm, smb + (rel & 1) * FusedSmem<NT>::RAW_SLOT, K, tid, x, xin, owned, gy0, gy1, i, off); \
        cta_barrier();                                                                                               \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#pragma unroll 1
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_H(true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#undef HG_ROW_H
        if (K.cta_ns && threadIdx.x == 0) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            K.cta_ns[blockIdx.x] = (unsigned)(t_end - t_start);
        }
    } else if (group == 1) {
        reg_set<RL, R0>();
#define HG_ROW_T(G, FREEFLAG)                                                                                        \
    {                                                                                                                \
        hg_fused_iter<NT, FREEFLAG, G, DROPS, SG>(c, sm, smb, K, tid, x, xin, owned, gy0, gy1, i, off);              \
        cta_barrier();                                                                                               \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_T(HGF_THERMAL0, false)
#pragma unroll 1
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_T(HGF_THERMAL0, true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_T(HGF_THERMAL0, false)
    } else {
        reg_set<RL, R1>();
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_T(HGF_THERMAL1, false)
#pragma unroll 1
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_T(HGF_THERMAL1, true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_T(HGF_THERMAL1, false)
#undef HG_ROW_T
#undef HG_TMA_ROW
    }
}

// ------------------------------------------------------------------ queued thermal outflow (hg_fused_body3.cuh)
// CTA = NT hydraulic threads + NT thermal threads + NSVC service warps.  The thermal threads only test their cell and
// queue the marked ones; the service warps evaluate thermal_erosion.glsl:59-115 for the queued cells of the whole CTA, 32
// per pass, one iteration later.  One CTA barrier per row as before.
template <int NT> struct FusedSmemQ {
    static constexpr size_t RAW_BOX = (size_t)HGF_NPL * HGF_RAW_LD(NT);
    static constexpr size_t RAW_SLOT = (RAW_BOX + 31) / 32 * 32;
    static constexpr size_t RINGS = 2 * RAW_SLOT;
    static constexpr size_t BARS = (RINGS + HgRingsQ<NT>::TOTAL + 3) / 4 * 4;
    static constexpr size_t BYTES = (BARS + 4) * sizeof(float);
};

template <int NT, int MINB, int NSVC, bool DROPS>
__global__ void __launch_bounds__(2 * NT + 32 * NSVC, MINB) k_fused_q(const __grid_constant__ HgFusedK K, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smb[];
    float* const sm = smb + FusedSmemQ<NT>::RINGS;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smb + FusedSmemQ<NT>::BARS);
    const int group = threadIdx.x < NT ? HGF_HYDRO : threadIdx.x < 2 * NT ? HGF_THERMAL : HGQ_SERVICE;      // warp-uniform
    const int tid = group == HGF_HYDRO ? threadIdx.x : group == HGF_THERMAL ? threadIdx.x - NT : threadIdx.x - 2 * NT;
    int strip, gy0, gy1;
    if (K.plan) {
        const HgPlanItem it = K.plan[blockIdx.x];
        strip = it.strip; gy0 = it.gy0; gy1 = it.gy1;
    } else {
        const int segi = blockIdx.x / K.nstrips;
        strip = blockIdx.x % K.nstrips;
        gy0 = K.row0 + segi * K.seg;
        gy1 = min(gy0 + K.seg, K.row0 + K.rows);
    }
    unsigned long long t_start = 0;
    if (K.cta_ns && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    const int x0 = strip * (NT - 2 * HGF_HX) - HGF_HX;
    const int x = x0 + tid;
    const bool xin = x >= 0 && x < K.W;
    const bool owned = tid >= HGF_HX && tid < NT - HGF_HX && x < K.W;
    const HgFusedPlan pl = hg_fusedq_plan(gy0, gy1, K.H);
    const unsigned pitch = (unsigned)K.pitch;
    unsigned off = (unsigned)(pl.i_begin - K.row0 + HG_HALO_ROWS) * pitch + (unsigned)x;
    const int ly0 = pl.i_begin - K.row0 + HG_HALO_ROWS;
    constexpr unsigned BOX_BYTES = DROPS ? (unsigned)(HGF_RAW_LD(NT) * 4 * sizeof(float)) : (unsigned)(FusedSmemQ<NT>::RAW_BOX * sizeof(float));
    const int bx0 = x0 - 2;
#define HG_TMA_ROW(dst, bar, row) (DROPS ? tma_load_3d((dst), &tmap, (bar), 0, bx0, (row)) : tma_load_3d((dst), &tmap, (bar), bx0, (row), 0))
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[0], BOX_BYTES);
        HG_TMA_ROW(smb, &bars[0], ly0);
    }
    if (threadIdx.x < 4) reinterpret_cast<unsigned*>(reinterpret_cast<char*>(sm) + HgRingsQ<NT>::QCNT)[threadIdx.x] = 0u;
    __syncthreads();
    int i = pl.i_begin;
    int m3 = (i - 3 + 3 * (1 << 20)) % 3;       // (i - 3) mod 3, i may be negative
    if (group == HGF_HYDRO) {
        HgColQ c;
        hg_colq_init(c);
#define HG_ROW_H(FREEFLAG)                                                                                           \
    {                                                                                                                \
        const int rel = i - pl.i_begin;                                                                              \
        if (tid == 0 && i < pl.i_end) {                                                                              \
            mbar_expect_tx(&bars[(rel + 1) & 1], BOX_BYTES);                                                         \
            HG_TMA_ROW(smb + ((rel + 1) & 1) * FusedSmemQ<NT>::RAW_SLOT, &bars[(rel + 1) & 1], ly0 + rel + 1);      \
        }                                                                                                            \
        mbar_wait(&bars[rel & 1], (unsigned)(rel >> 1) & 1u);                                                        \
        hg_fusedq_iter<NT, FREEFLAG, HGF_HYDRO, DROPS>(c, sm, smb + (rel & 1) * FusedSmemQ<NT>::RAW_SLOT, K, tid, x, xin, owned, gy0, gy1, i, m3, off); \
        cta_barrier();                                                                                               \
        m3 = m3 == 2 ? 0 : m3 + 1;                                                                                   \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#pragma unroll 1
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_H(true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#undef HG_ROW_H
        if (K.cta_ns && threadIdx.x == 0) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            K.cta_ns[blockIdx.x] = (unsigned)(t_end - t_start);
        }
    } else if (group == HGF_THERMAL) {
        HgColQ c;
        hg_colq_init(c);
#define HG_ROW_T(FREEFLAG)                                                                                           \
    {                                                                                                                \
        hg_fusedq_iter<NT, FREEFLAG, HGF_THERMAL, DROPS>(c, sm, smb, K, tid, x, xin, owned, gy0, gy1, i, m3, off);   \
        cta_barrier();                                                                                               \
        m3 = m3 == 2 ? 0 : m3 + 1;                                                                                   \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_T(false)
#pragma unroll 1
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_T(true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_T(false)
#undef HG_ROW_T
    } else {
        // service warps: the cells queued during the previous iteration, 32 per pass, alternating between the warps
        typedef HgRingsQ<NT> R;
        char* const smc = reinterpret_cast<char*>(sm);
        volatile unsigned* const qcnt = reinterpret_cast<volatile unsigned*>(smc + R::QCNT);
        const unsigned short* const qbuf = reinterpret_cast<const unsigned short*>(smc + R::QUEUE);
#pragma unroll 1
        for (; i <= pl.i_end; i++) {
            const unsigned n = qcnt[(i - 1) & 3];
            for (unsigned b = (unsigned)tid; b < n; b += 32u * NSVC)
                hg_fusedq_serve<NT>(sm, K, i, m3, qbuf[((i - 1) & 1) * (2 * NT) + b]);
            if (tid == 0) qcnt[(i + 1) & 3] = 0u;
            cta_barrier();
            m3 = m3 == 2 ? 0 : m3 + 1;
        }
    }
#undef HG_TMA_ROW
}

// ------------------------------------------------------------------ two columns per thread (hg_fused_body2.cuh)
// Same warp-specialised shell; a CTA of 2*NT threads owns a PAIR of adjacent strips and every thread carries column t
// of both strips in the two lanes of packed fp32 arithmetic.  One TMA box per row covers both strips
// (2*NT - 8 columns x 9 planes).
template <int NT> struct FusedSmem2 {
    static constexpr size_t RAW_BOX = (size_t)HGF_NPL * HGF2_RAW_LD(NT);
    static constexpr size_t RAW_SLOT = (RAW_BOX + 31) / 32 * 32;
    static constexpr size_t RINGS = 2 * RAW_SLOT;
    static constexpr size_t BARS = (RINGS + HgRings2<NT>::TOTAL + 3) / 4 * 4;
    static constexpr size_t BYTES = (BARS + 4) * sizeof(float);
};

template <int NT, int MINB, int RH, int RT>
__global__ void __launch_bounds__(2 * NT, MINB) k_fused_ws2(const __grid_constant__ HgFusedK K, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smb[];
    float* const sm = smb + FusedSmem2<NT>::RINGS;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smb + FusedSmem2<NT>::BARS);
    const bool hydro = threadIdx.x < NT;
    const int tid = hydro ? threadIdx.x : threadIdx.x - NT;
    int strip, gy0, gy1;      // strip = index of the strip PAIR
    if (K.plan) {
        const HgPlanItem it = K.plan[blockIdx.x];
        strip = it.strip; gy0 = it.gy0; gy1 = it.gy1;
    } else {
        const int segi = blockIdx.x / K.nstrips;
        strip = blockIdx.x % K.nstrips;
        gy0 = K.row0 + segi * K.seg;
        gy1 = min(gy0 + K.seg, K.row0 + K.rows);
    }
    unsigned long long t_start = 0;
    if (K.cta_ns && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    constexpr int HALF = HGF2_HALF(NT);
    const int x0 = strip * (2 * HALF) - HGF_HX;
    const HgLanes L = hg_lanes(x0 + tid, HALF, tid, NT, K.W);
    const HgFusedPlan pl = hg_fused_plan(gy0, gy1, K.H);
    const unsigned pitch = (unsigned)K.pitch;
    unsigned off = (unsigned)(pl.i_begin - K.row0 + HG_HALO_ROWS) * pitch + (unsigned)(x0 + tid);
    const int ly0 = pl.i_begin - K.row0 + HG_HALO_ROWS;
    constexpr unsigned BOX_BYTES = (unsigned)(FusedSmem2<NT>::RAW_BOX * sizeof(float));
    const int bx0 = x0 - 2;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[0], BOX_BYTES);
        tma_load_3d(smb, &tmap, &bars[0], bx0, ly0, 0);
    }
    __syncthreads();
    HgCol2 c;
    hg_col2_init(c);
    int i = pl.i_begin;
    if (hydro) {
        if (RH != RT) reg_dec<RH>();
#define HG_ROW_H(FREEFLAG)                                                                                           \
    {                                                                                                                \
        const int rel = i - pl.i_begin;                                                                              \
        if (tid == 0 && i < pl.i_end) {                                                                              \
            mbar_expect_tx(&bars[(rel + 1) & 1], BOX_BYTES);                                                         \
            tma_load_3d(smb + ((rel + 1) & 1) * FusedSmem2<NT>::RAW_SLOT, &tmap, &bars[(rel + 1) & 1], bx0, ly0 + rel + 1, 0); \
        }                                                                                                            \
        mbar_wait(&bars[rel & 1], (unsigned)(rel >> 1) & 1u);                                                        \
        hg_fused_iter2<NT, FREEFLAG, HGF_HYDRO>(c, sm, smb + (rel & 1) * FusedSmem2<NT>::RAW_SLOT, K, tid, L, gy0, gy1, i, off); \
        cta_barrier();                                                                                               \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#pragma unroll 1
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_H(true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_H(false)
#undef HG_ROW_H
        if (K.cta_ns && threadIdx.x == 0) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            K.cta_ns[blockIdx.x] = (unsigned)(t_end - t_start);
        }
    } else {
        if (RH != RT) reg_inc<RT>();
#define HG_ROW_T(FREEFLAG)                                                                                           \
    {                                                                                                                \
        hg_fused_iter2<NT, FREEFLAG, HGF_THERMAL>(c, sm, smb, K, tid, L, gy0, gy1, i, off);                          \
        cta_barrier();                                                                                               \
    }
        for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) HG_ROW_T(false)
#pragma unroll 1
        for (; i <= pl.free_hi; i++, off += pitch) HG_ROW_T(true)
        for (; i <= pl.i_end; i++, off += pitch) HG_ROW_T(false)
#undef HG_ROW_T
    }
}

// ------------------------------------------------------------------ balanced partition
// All CTAs of a step are resident at once (one wave, 3 per SM), so the step ends when the SLOWEST CTA
// ends.  With equal segments of rows that was 0.69 ms at 4096^2 while the mean CTA took 0.55 ms: the
// cost of a row depends on the terrain (the thermal outflow path is taken only where a cell is marked)
// and on the strip (the ones overhanging the map edge).  Terrain changes slowly, so the durations
// the CTAs of one step report are a good forecast for the next: this kernel re-cuts every strip into
// row segments of equal FORECAST cost, handing a strip a share of the n_cta segments proportional to
// its total.  One thread per strip; the cost inside an old segment is taken as uniform.  Any partition
// gives the same bits (segments only decide who computes a row), so a poor forecast costs time only.
struct PlanArgs {
    const HgPlanItem* old_plan; const unsigned* cta_ns; HgPlanItem* new_plan;
    int n_cta, nstrips, row0, rows, min_rows;
};
// One CTA.  The arithmetic is hg_plan.cuh (shared with the CPU tests); here only the bookkeeping: where each
// strip's old segments start, the per-strip cost sums, and one thread per strip for the cuts.
__global__ void __launch_bounds__(512) k_plan_segments(PlanArgs A) {
    __shared__ float strip_cost[256], frac[256], total_s;
    __shared__ int strip_first[257], new_n[256], new_first[257], left_s;
    const int t = threadIdx.x;
    // old segments are grouped by strip, in row order: a strip starts where the strip index changes
    for (int b = t; b < A.n_cta; b += blockDim.x)
        if (b == 0 || A.old_plan[b].strip != A.old_plan[b - 1].strip) strip_first[A.old_plan[b].strip] = b;
    if (t == 0) strip_first[A.nstrips] = A.n_cta;
    __syncthreads();
    if (t < A.nstrips) {
        float c = 0.0f;
        for (int k = strip_first[t]; k < strip_first[t + 1]; k++) c += hg_plan_cost(A.cta_ns[k]);
        strip_cost[t] = c;
    }
    __syncthreads();
    const int cap = max(1, A.rows / A.min_rows);
    if (t == 0) {
        float total = 0.0f; 
        for (int k = 0; k < A.nstrips; k++) total += strip_cost[k];
        total_s = total;
    }
    __syncthreads();
    if (t < A.nstrips) hg_plan_share(A.n_cta, strip_cost[t], total_s, cap, &new_n[t], &frac[t]);
    __syncthreads();
    if (t == 0) {
        int g = 0;
        for (int k = 0; k < A.nstrips; k++) g += new_n[k];
        left_s = A.n_cta - g;
    }
    __syncthreads();
    const int bonus = t < A.nstrips ? hg_plan_bonus(t, A.nstrips, frac, left_s, cap, new_n[t]) : 0;
    __syncthreads();
    if (t < A.nstrips) new_n[t] += bonus;
    __syncthreads();
    if (t == 0) {
        hg_plan_repair(A.n_cta, A.nstrips, strip_cost, cap, new_n);
        int b = 0;
        for (int k = 0; k < A.nstrips; k++) { new_first[k] = b; b += new_n[k]; }
        new_first[A.nstrips] = b;
    }
    __syncthreads();
    if (t < A.nstrips)
        hg_plan_cut_strip(t, new_n[t], A.old_plan + strip_first[t], A.cta_ns + strip_first[t], strip_first[t + 1] - strip_first[t],
                          A.row0, A.rows, A.min_rows, A.new_plan + new_first[t]);
}

}  // namespace

// 3-D tensor map over one ping-pong set of the arena: (column, plane row, plane), box (NT+4) x 1 x 9.
// The box may not exceed 256 columns, so NT <= 252.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tmap(hg_ctx* c, int set, int box_cols, CUtensorMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        HG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { hg_set_error("cuTensorMapEncodeTiled is not available in this driver"); return HG_ERR_CUDA; }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t dims[3] = {(cuuint64_t)c->g.W, (cuuint64_t)c->g.rows_alloc, (cuuint64_t)HG_NPLANES};
    cuuint64_t strides[2] = {(cuuint64_t)c->g.pitch * sizeof(float), (cuuint64_t)c->g.plane_elems * sizeof(float)};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, 1u, (cuuint32_t)HG_NPLANES};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, hg_plane(c, set, 0), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hg_set_error("cuTensorMapEncodeTiled failed (%d) for a %dx%d slab, box %d", (int)r, c->g.W, c->g.rows_alloc, box_cols); return HG_ERR_CUDA; }
    return HG_OK;
}

// 3-D tensor map over one heightmap image in texture layout (droplet mode): (channel, column, row), box 4 x (NT+4) x 1.
static int make_tmap_aos(hg_ctx* c, const float4* image, int box_cols, CUtensorMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        HG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { hg_set_error("cuTensorMapEncodeTiled is not available in this driver"); return HG_ERR_CUDA; }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t dims[3] = {4u, (cuuint64_t)c->g.W, (cuuint64_t)c->g.rows_alloc};
    cuuint64_t strides[2] = {4 * sizeof(float), (cuuint64_t)c->g.pitch * 4 * sizeof(float)};
    cuuint32_t box[3] = {4u, (cuuint32_t)box_cols, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float4*>(image), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hg_set_error("cuTensorMapEncodeTiled failed (%d) for a %dx%d texel image, box %d", (int)r, c->g.W, c->g.rows_alloc, box_cols); return HG_ERR_CUDA; }
    return HG_OK;
}

// NT threads per CTA; seg rows per CTA.
template <int NT, int MINB, bool DROPS = false>
static int launch_main(hg_ctx* c, const HgFusedK& K0, int seg, int src_set) {
    HgFusedK K = K0;
    K.nstrips = (c->g.W + (NT - 2 * HGF_HX) - 1) / (NT - 2 * HGF_HX);
    K.seg = seg;
    int nseg = (c->g.rows + seg - 1) / seg;
    constexpr size_t smem = FusedSmem<NT>::BYTES;
    static bool attr_set[HG_MAX_DEVICES] = {};      // the attribute is per device (one process may drive several GPUs)
    if (c->device >= HG_MAX_DEVICES || !attr_set[c->device]) {
        HG_CUDA(cudaFuncSetAttribute(k_fused_step<NT, MINB, DROPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->device < HG_MAX_DEVICES) attr_set[c->device] = true;
    }
    alignas(64) CUtensorMap tmap;
    int rc = DROPS ? make_tmap_aos(c, reinterpret_cast<const float4*>(K.ha_src), HGF_RAW_LD(NT), &tmap) : make_tmap(c, src_set, HGF_RAW_LD(NT), &tmap);
    if (rc) return rc;
    if (c->prof_ev0) HG_CUDA(cudaEventRecord(c->prof_ev0, c->stream));
    k_fused_step<NT, MINB, DROPS><<<K.nstrips * nseg, NT, smem, c->stream>>>(K, tmap);
    HG_LAUNCH_CHECK(c);
    if (c->prof_ev1) HG_CUDA(cudaEventRecord(c->prof_ev1, c->stream));
    return HG_OK;
}

template <int NT, int MINB, int RH, int RT, bool DROPS = false, int SWAP = 0>
static int launch_ws(hg_ctx* c, const HgFusedK& K0, int seg, int src_set) {
    static_assert((RH + RT) / 2 * 2 * NT * MINB <= 65536 && RH % 8 == 0 && RT % 8 == 0, "register budget of the two warp groups");
    HgFusedK K = K0;
    K.nstrips = (c->g.W + (NT - 2 * HGF_HX) - 1) / (NT - 2 * HGF_HX);
    K.seg = seg;
    int nseg = (c->g.rows + seg - 1) / seg;
    constexpr size_t smem = FusedSmem<NT>::BYTES;
    static bool attr_set[HG_MAX_DEVICES] = {};
    if (c->device >= HG_MAX_DEVICES || !attr_set[c->device]) {
        HG_CUDA(cudaFuncSetAttribute(k_fused_ws<NT, MINB, RH, RT, DROPS, SWAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->device < HG_MAX_DEVICES) attr_set[c->device] = true;
    }
    alignas(64) CUtensorMap tmap;
    int rc = DROPS ? make_tmap_aos(c, reinterpret_cast<const float4*>(K.ha_src), HGF_RAW_LD(NT), &tmap) : make_tmap(c, src_set, HGF_RAW_LD(NT), &tmap);
    if (rc) return rc;
    if (c->prof_ev0) HG_CUDA(cudaEventRecord(c->prof_ev0, c->stream));
    k_fused_ws<NT, MINB, RH, RT, DROPS, SWAP><<<K.plan ? c->plan_n : K.nstrips * nseg, 2 * NT, smem, c->stream>>>(K, tmap);
    HG_LAUNCH_CHECK(c);
    if (c->prof_ev1) HG_CUDA(cudaEventRecord(c->prof_ev1, c->stream));
    return HG_OK;
}


template <int NT, int MINB, int RH, int R0, int R1, bool DROPS = false, int SG = HGF_THERMAL0>
static int launch_ws3(hg_ctx* c, const HgFusedK& K0, int seg, int src_set) {
    HgFusedK K = K0;
    K.nstrips = (c->g.W + (NT - 2 * HGF_HX) - 1) / (NT - 2 * HGF_HX);
    K.seg = seg;
    int nseg = (c->g.rows + seg - 1) / seg;
    constexpr size_t smem = FusedSmem<NT>::BYTES;
    static bool attr_set[HG_MAX_DEVICES] = {};
    if (c->device >= HG_MAX_DEVICES || !attr_set[c->device]) {
        HG_CUDA(cudaFuncSetAttribute(k_fused_ws3<NT, MINB, RH, R0, R1, DROPS, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->device < HG_MAX_DEVICES) attr_set[c->device] = true;
    }
    alignas(64) CUtensorMap tmap;
    int rc = DROPS ? make_tmap_aos(c, reinterpret_cast<const float4*>(K.ha_src), HGF_RAW_LD(NT), &tmap) : make_tmap(c, src_set, HGF_RAW_LD(NT), &tmap);
    if (rc) return rc;
    if (c->prof_ev0) HG_CUDA(cudaEventRecord(c->prof_ev0, c->stream));
    k_fused_ws3<NT, MINB, RH, R0, R1, DROPS, SG><<<K.plan ? c->plan_n : K.nstrips * nseg, 3 * NT, smem, c->stream>>>(K, tmap);
    HG_LAUNCH_CHECK(c);
    if (c->prof_ev1) HG_CUDA(cudaEventRecord(c->prof_ev1, c->stream));
    return HG_OK;
}

template <int NT, int MINB, int NSVC, bool DROPS = false>
static int launch_q(hg_ctx* c, const HgFusedK& K0, int seg, int src_set) {
    HgFusedK K = K0;
    K.nstrips = (c->g.W + (NT - 2 * HGF_HX) - 1) / (NT - 2 * HGF_HX);
    K.seg = seg;
    int nseg = (c->g.rows + seg - 1) / seg;
    constexpr size_t smem = FusedSmemQ<NT>::BYTES;
    static bool attr_set[HG_MAX_DEVICES] = {};
    if (c->device >= HG_MAX_DEVICES || !attr_set[c->device]) {
        HG_CUDA(cudaFuncSetAttribute(k_fused_q<NT, MINB, NSVC, DROPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->device < HG_MAX_DEVICES) attr_set[c->device] = true;
    }
    alignas(64) CUtensorMap tmap;
    int rc = DROPS ? make_tmap_aos(c, reinterpret_cast<const float4*>(K.ha_src), HGF_RAW_LD(NT), &tmap) : make_tmap(c, src_set, HGF_RAW_LD(NT), &tmap);
    if (rc) return rc;
    if (c->prof_ev0) HG_CUDA(cudaEventRecord(c->prof_ev0, c->stream));
    k_fused_q<NT, MINB, NSVC, DROPS><<<K.plan ? c->plan_n : K.nstrips * nseg, 2 * NT + 32 * NSVC, smem, c->stream>>>(K, tmap);
    HG_LAUNCH_CHECK(c);
    if (c->prof_ev1) HG_CUDA(cudaEventRecord(c->prof_ev1, c->stream));
    return HG_OK;
}

// two columns per thread: the strips of the launch are strip PAIRS of 2 * (NT - 12) owned columns
template <int NT, int MINB, int RH, int RT>
static int launch_ws2(hg_ctx* c, const HgFusedK& K0, int seg, int src_set) {
    static_assert((RH + RT) / 2 * 2 * NT * MINB <= 65536 && RH % 8 == 0 && RT % 8 == 0, "register budget of the two warp groups");
    HgFusedK K = K0;
    K.nstrips = (c->g.W + 2 * HGF2_HALF(NT) - 1) / (2 * HGF2_HALF(NT));
    K.seg = seg;
    int nseg = (c->g.rows + seg - 1) / seg;
    constexpr size_t smem = FusedSmem2<NT>::BYTES;
    static bool attr_set[HG_MAX_DEVICES] = {};
    if (c->device >= HG_MAX_DEVICES || !attr_set[c->device]) {
        HG_CUDA(cudaFuncSetAttribute(k_fused_ws2<NT, MINB, RH, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->device < HG_MAX_DEVICES) attr_set[c->device] = true;
    }
    alignas(64) CUtensorMap tmap;
    int rc = make_tmap(c, src_set, HGF2_RAW_LD(NT), &tmap);
    if (rc) return rc;
    if (c->prof_ev0) HG_CUDA(cudaEventRecord(c->prof_ev0, c->stream));
    k_fused_ws2<NT, MINB, RH, RT><<<K.plan ? c->plan_n : K.nstrips * nseg, 2 * NT, smem, c->stream>>>(K, tmap);
    HG_LAUNCH_CHECK(c);
    if (c->prof_ev1) HG_CUDA(cudaEventRecord(c->prof_ev1, c->stream));
    return HG_OK;
}

// The TMA box spans all nine planes, so the read planes of H, F and S must sit in the same
// ping-pong set.  They always do on the FUSED schedule (all three flip together, rain is in
// place); after PASSES dispatches they may not, and the odd ones are copied across once.
static int align_sets(hg_ctx* c) {
    const size_t pb = c->g.plane_elems * sizeof(float);
    if (c->ri[1] != c->ri[0]) {
        HG_CUDA(cudaMemcpyAsync(hg_plane(c, c->ri[0], PL_FL), hg_plane(c, c->ri[1], PL_FL), 4 * pb, cudaMemcpyDeviceToDevice, c->stream));
        c->ri[1] = c->ri[0];
    }
    if (c->ri[3] != c->ri[0]) {
        HG_CUDA(cudaMemcpyAsync(hg_plane(c, c->ri[0], PL_SR), hg_plane(c, c->ri[3], PL_SR), 2 * pb, cudaMemcpyDeviceToDevice, c->stream));
        c->ri[3] = c->ri[0];
    }
    return HG_OK;
}

// Under CUDA's lazy module loading the FIRST launch of a kernel loads its code, which can wait for kernels that are
// running -- and a slab's stream may hold a spinning halo wait whose release needs another slab's launch from the same
// host thread.  Connected contexts therefore load every kernel of the step path up front (hg_slab_connect*).
int hg_preload_fused_kernels(void) {
    cudaFuncAttributes a;
    HG_CUDA(cudaFuncGetAttributes(&a, k_fused_ws<128, 3, 72, 88, false>));
    HG_CUDA(cudaFuncGetAttributes(&a, k_fused_ws<128, 3, 72, 88, true>));
    HG_CUDA(cudaFuncGetAttributes(&a, k_fused_ws2<128, 2, 120, 136>));
    HG_CUDA(cudaFuncGetAttributes(&a, k_far_fixup));
    HG_CUDA(cudaFuncGetAttributes(&a, k_plan_segments));
    HG_CUDA(cudaFuncSetAttribute(k_fused_ws<128, 3, 72, 88, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedSmem<128>::BYTES));
    HG_CUDA(cudaFuncSetAttribute(k_fused_ws<128, 3, 72, 88, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedSmem<128>::BYTES));
    return HG_OK;
}

// drops = false: Erosion::dispatch_grid.  drops = true: the grid part of Erosion::dispatch_particle
// (src/erosion.cpp:146-155: thermal x2 + smoothing with the momentum map) on the same kernel, the hydraulic warp
// group only feeding (rock, dirt) to the thermal one; reads H and the momentum map, writes the other sets and H.a.
static int launch_fused(hg_ctx* c, bool drops);
int hg_launch_fused_step(hg_ctx* c) { return launch_fused(c, false); }
int hg_launch_fused_thermal_smooth_particle(hg_ctx* c) { return launch_fused(c, true); }

static int launch_fused(hg_ctx* c, bool drops) {
    if (c->g.plane_elems >= (size_t)1 << 32) { hg_set_error("slab too large for 32-bit plane offsets (%zu elements)", c->g.plane_elems); return HG_ERR_INVALID; }
    {   // ghost rows and peer planes of the previous exchange generation must have landed
        int rcw = hg_slab_wait_pending(c);
        if (rcw) return rcw;
    }
    if (!drops) {
        int rca = align_sets(c);
        if (rca) return rca;
    }
    FusedArgs A;
    memset(&A, 0, sizeof(A));
    HgFusedK K;
    memset(&K, 0, sizeof(K));
    for (int p = 0; p < HG_NPLANES; p++) {
        A.src[p] = K.src[p] = hg_cur(c, p, 1);
        A.dst[p] = K.dst[p] = hg_cur(c, p, 0);
        A.src_set[p] = c->ri[hg_field_of_plane(p)];
    }
    A.W = K.W = c->g.W; A.H = K.H = c->g.H; A.row0 = K.row0 = c->g.row0; A.rows = K.rows = c->g.rows; A.pitch = K.pitch = c->g.pitch;
    if (c->slabs.n > 0) {
        A.slabs = c->slabs;
    } else {
        A.slabs.n = 1; A.slabs.me = 0;
        A.slabs.arena[0] = c->arena; A.slabs.row0[0] = c->g.row0; A.slabs.rows[0] = c->g.rows;
    }
    if (!drops && !c->far_list) {   // one entry per owned cell: correct even if every back-trace is far
        HG_CUDA(cudaMalloc(&c->far_list, (size_t)c->g.rows * c->g.W * sizeof(unsigned)));
    }
    A.far_list = K.far_list = c->far_list;
    A.far_count = K.far_count = c->d_counters + 8 + (c->far_parity & 1);
    A.far_count_next = c->d_counters + 8 + ((c->far_parity + 1) & 1);
    A.far_total = c->d_counters;
    c->far_parity ^= 1;
    A.P = K.P = c->sp;
    // Connected grid slab on the default kernel: the step kernel and the fix-up store the edge rows straight into the
    // neighbours' ghost rows and the fix-up's last block signals the generation -- no separate push kernel
    // (hg_slab_exchange then only books the generation).  HG_FUSED_PUSH=0: the separate k_halo_push_signal.
    bool fused_push = false;
    if (!drops && c->peers_connected && !c->no_fused_push && (c->tune_variant < 0 || c->tune_variant == 5) && HG_FUSED_DEFAULT_VARIANT == 5) {
        float* pp[2][HG_NPLANES];
        int mask = 0;
        int rcp = hg_slab_peer_planes(c, pp, &mask);
        if (rcp) return rcp;
        for (int sd = 0; sd < 2; sd++) {
            for (int p = 0; p < HG_NPLANES; p++) K.peer[sd][p] = pp[sd][p];
            A.peer[sd][0] = pp[sd][PL_SR]; A.peer[sd][1] = pp[sd][PL_SD];
        }
        A.push_mask = K.push_mask = mask;
        A.push_lo_end = K.push_lo_end = c->g.row0 + HG_HALO_ROWS;
        A.push_hi_begin = K.push_hi_begin = c->g.row0 + c->g.rows - HG_HALO_ROWS;
        c->step_flag++;
        A.gen = c->step_flag;
        hg_slab_signal_args(c, &A.sig);
        A.done = reinterpret_cast<unsigned*>(c->d_counters + 11);
        fused_push = true;
    }
    if (drops) {
        // heightmap and momentum map in texture layout (hg_particle_layout): read images -> write images
        int rcl = hg_particle_layout(c, true);
        if (rcl) return rcl;
        K.ha_src = reinterpret_cast<const HgF4*>(hg_pa_h(c, 1)); K.ha_dst = reinterpret_cast<HgF4*>(hg_pa_h(c, 0));
        K.ma_src = reinterpret_cast<const HgF4*>(hg_pa_m(c, 1)); K.ma_dst = reinterpret_cast<HgF4*>(hg_pa_m(c, 0));
    }
    // CTA shape (threads, resident CTAs per SM); HG_FUSED_VARIANT / HG_FUSED_SEG override (tuning aids)
    // variants 5..: warp-specialised (k_fused_ws), 2 warp groups per CTA
    // variants 10..: two columns per thread (k_fused_ws2): strip pairs, 2 CTAs of 256 threads per SM
    // variants 13..: three warp groups per CTA (k_fused_ws3)
    // variants 18..: queued thermal outflow with service warps (k_fused_q)
    // variants 21..23: k_fused_ws with the groups swapped / 192-column strips
    // variants 25, 26: k_fused_ws without register rebalancing (80 / 80), groups interleaved in warp pairs / in order
    constexpr int NVAR = 27;
    static const int nt_of[NVAR] = {128, 128, 192, 224, 224, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 192, 192, 192, 128, 128};
    static const int res_of[NVAR] = {4, 3, 2, 2, 1, 3, 2, 3, 4, 4, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 2, 3, 3};
    static const int wpc_of[NVAR] = {4, 4, 6, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 12, 12, 12, 12, 12, 9, 10, 12, 8, 12, 12, 13, 8, 8};
    int v = c->tune_variant >= 0 && c->tune_variant < NVAR ? c->tune_variant : HG_FUSED_DEFAULT_VARIANT;
    if (drops) v = c->tune_drops_variant == 1 ? 3 : c->tune_drops_variant == 2 ? 13 : c->tune_drops_variant == 3 ? 18 : 5;      // HG_DROPS_VARIANT=1: one warp group of 224 threads runs every stage; 2: three groups
    const bool two_lane = v >= 10 && v < 13;
    const bool three_groups = v >= 13;      // (and every later multi-group kernel: they all take the balanced partition)      // (and the queued kernels: every multi-group kernel takes the balanced partition)
    const bool queued = v >= 18;
    const int NT = nt_of[v];
    const int strip_w = two_lane ? 2 * HGF2_HALF(NT) : NT - 2 * HGF_HX;      // owned columns per CTA
    int nstrips = (c->g.W + strip_w - 1) / strip_w;
    // Rows per CTA.  A CTA runs seg + 17 row iterations (pipeline fill), about 8 of them of the
    // slower non-FREE kind.  An SM's time for a wave of w resident warps was measured as roughly
    // proportional to 10 + 0.375 w per row iteration (8 warps reach 61 % of the throughput of 16).
    // Pick the segment count that minimises the sum over the waves of this grid.
    int seg = c->tune_seg;
    if (seg <= 0) {
        const int res = res_of[v], wpc = wpc_of[v];
        double best = 1e30;
        for (int nseg = 1; nseg <= c->g.rows / 16 + 1 && nseg <= 4096; nseg++) {
            int sg = (c->g.rows + nseg - 1) / nseg;
            long long ctas = (long long)nstrips * ((c->g.rows + sg - 1) / sg);
            long long per_sm = (ctas + 147) / 148;                 // CTAs the busiest SM runs
            long long full = per_sm / res, last = per_sm % res;    // full waves + a partial one
            double iters = sg + HGF_HX + HGF_LAG_G + 8;
            double cost = iters * (full * (10.0 + 0.375 * res * wpc) + (last ? 10.0 + 0.375 * last * wpc : 0.0));
            if (cost < best) { best = cost; seg = sg; }
        }
    }
    // Balanced partition (default for the warp-specialised kernel when the slab is big enough to fill the GPU):
    // n_cta = 3 CTAs on every SM, cut per strip by the forecast of k_plan_segments.  HG_FUSED_SEG or
    // HG_FUSED_BALANCE=0 keep the uniform segments.
    bool balanced = false;
    PlanArgs plan_args{};
    // (droplet slabs keep uniform segments: the first use of a plan allocates and synchronises, which a host thread
    // driving several slabs of one process must not do between two exchange generations)
    if ((v == 5 || two_lane || three_groups) && c->tune_seg <= 0 && !c->no_balance && !(drops && c->peers_connected)) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        // Only with at least six segments per strip: with fewer (16384 columns: 3.1 per strip) one segment more or
        // less is too coarse a step, and moving only the cuts inside a strip was measured 5 % SLOWER than uniform
        // segments there (CTAs of 5000 rows already average the terrain; the forecast then mostly carries noise).
        const int min_rows = 48;
        const int n_cta = res_of[v] * sms;
        if (n_cta / nstrips >= 6 && (long long)nstrips * (c->g.rows / min_rows) >= 2LL * n_cta) {
            if (!c->plan[0]) {      // first use: uniform cut into n_cta pieces (strip k gets n_cta/nstrips, the first few one more)
                HG_CUDA(cudaMalloc(&c->plan[0], 2 * n_cta * sizeof(HgPlanItem)));
                c->plan[1] = c->plan[0] + n_cta;
                HG_CUDA(cudaMalloc(&c->cta_ns, n_cta * sizeof(unsigned)));
                std::vector<HgPlanItem> h;
                for (int k = 0; k < nstrips; k++) {
                    const int n = n_cta / nstrips + (k < n_cta % nstrips ? 1 : 0);
                    for (int m = 0; m < n; m++) {
                        HgPlanItem it;
                        it.strip = k; it.pad = 0;
                        it.gy0 = c->g.row0 + (int)((long long)c->g.rows * m / n);
                        it.gy1 = c->g.row0 + (int)((long long)c->g.rows * (m + 1) / n);
                        h.push_back(it);
                    }
                }
                HG_CUDA(cudaMemcpyAsync(c->plan[0], h.data(), h.size() * sizeof(HgPlanItem), cudaMemcpyHostToDevice, c->stream));
                HG_CUDA(cudaStreamSynchronize(c->stream));      // h goes out of scope
                c->plan_cur = 0; c->plan_n = n_cta; c->plan_valid = false;
            }
            if (c->plan_valid) {    // the cut for this step was made on the side stream while the previous step finished
                HG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_plan, 0));
                c->plan_valid = false;
            }
            K.plan = c->plan[c->plan_cur];
            K.cta_ns = c->cta_ns;
            balanced = true;
            plan_args = PlanArgs{c->plan[c->plan_cur], c->cta_ns, c->plan[c->plan_cur ^ 1], n_cta, nstrips, c->g.row0, c->g.rows, min_rows};
        }
    }
    int rc;
    switch (v) {
    case 0: rc = launch_main<128, 4>(c, K, seg, c->ri[0]); break;
    case 1: rc = launch_main<128, 3>(c, K, seg, c->ri[0]); break;
    case 2: rc = launch_main<192, 2>(c, K, seg, c->ri[0]); break;
    case 3: rc = drops ? launch_main<224, 3, true>(c, K, seg, c->ri[0]) : launch_main<224, 2>(c, K, seg, c->ri[0]); break;
    case 4: rc = launch_main<224, 1>(c, K, seg, c->ri[0]); break;
    case 5: rc = drops ? launch_ws<128, 3, 72, 88, true>(c, K, seg, c->ri[0]) : launch_ws<128, 3, 72, 88>(c, K, seg, c->ri[0]); break;
    case 6: rc = launch_ws<128, 2, 96, 128>(c, K, seg, c->ri[0]); break;
    case 7: rc = launch_ws<128, 3, 64, 96>(c, K, seg, c->ri[0]); break;
    case 8: rc = launch_ws<128, 4, 64, 64>(c, K, seg, c->ri[0]); break;
    case 9: rc = launch_ws<128, 4, 56, 72>(c, K, seg, c->ri[0]); break;
    case 10: rc = launch_ws2<128, 2, 120, 136>(c, K, seg, c->ri[0]); break;
    case 11: rc = launch_ws2<128, 2, 128, 128>(c, K, seg, c->ri[0]); break;
    case 12: rc = launch_ws2<128, 2, 112, 144>(c, K, seg, c->ri[0]); break;
    case 13: rc = drops ? launch_ws3<128, 3, 56, 56, 56, true>(c, K, seg, c->ri[0]) : launch_ws3<128, 3, 72, 48, 48, false, HGF_THERMAL0>(c, K, seg, c->ri[0]); break;
    case 14: rc = launch_ws3<128, 3, 64, 48, 56, false, HGF_THERMAL1>(c, K, seg, c->ri[0]); break;
    case 15: rc = launch_ws3<128, 3, 64, 56, 48, false, HGF_THERMAL0>(c, K, seg, c->ri[0]); break;
    case 16: rc = launch_ws3<128, 3, 72, 48, 48, false, HGF_THERMAL1>(c, K, seg, c->ri[0]); break;
    case 17: rc = launch_ws3<128, 3, 64, 48, 56, false, HGF_THERMAL0>(c, K, seg, c->ri[0]); break;
    case 18: rc = drops ? launch_q<128, 3, 1, true>(c, K, seg, c->ri[0]) : launch_q<128, 3, 1>(c, K, seg, c->ri[0]); break;
    case 19: rc = launch_q<128, 3, 2>(c, K, seg, c->ri[0]); break;
    case 20: rc = launch_q<128, 3, 4>(c, K, seg, c->ri[0]); break;
    case 21: rc = launch_ws<128, 3, 72, 88, false, 1>(c, K, seg, c->ri[0]); break;
    case 22: rc = launch_ws<192, 2, 80, 80>(c, K, seg, c->ri[0]); break;      // (setmaxnreg works on whole warpgroups of 4 warps: a 6 + 6 warp CTA cannot rebalance)
    case 23: rc = launch_ws<192, 2, 80, 80, false, 1>(c, K, seg, c->ri[0]); break;
    case 24: rc = launch_q<192, 2, 1>(c, K, seg, c->ri[0]); break;
    case 25: rc = launch_ws<128, 3, 80, 80, false, 2>(c, K, seg, c->ri[0]); break;
    default: rc = launch_ws<128, 3, 80, 80>(c, K, seg, c->ri[0]); break;
    }
    if (rc) return rc;
    if (balanced) {
        // durations of this step -> the next step's cut, on a side stream beside the fix-up kernel (both are tiny
        // and latency-bound; the next step kernel waits for ev_plan)
        if (!c->plan_stream) {
            HG_CUDA(cudaStreamCreateWithFlags(&c->plan_stream, cudaStreamNonBlocking));
            HG_CUDA(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
            HG_CUDA(cudaEventCreateWithFlags(&c->ev_plan, cudaEventDisableTiming));
        }
        HG_CUDA(cudaEventRecord(c->ev_main, c->stream));
        HG_CUDA(cudaStreamWaitEvent(c->plan_stream, c->ev_main, 0));
        k_plan_segments<<<1, 512, 0, c->plan_stream>>>(plan_args);
        HG_LAUNCH_CHECK(c);
        HG_CUDA(cudaEventRecord(c->ev_plan, c->plan_stream));
        c->plan_cur ^= 1;
        c->plan_valid = true;
    }
    if (drops) {
        c->ri[0] ^= 1; c->ri[2] ^= 1;      // heightmap and momentum map were written into their other textures
        return HG_OK;
    }
    k_far_fixup<<<148 * 2, 128, 0, c->stream>>>(A);
    HG_LAUNCH_CHECK(c);
    if (fused_push) c->fused_push_gen = A.gen;      // hg_slab_exchange: already pushed and signalled
    for (int f = 0; f < 4; f++) if (f != 2) c->ri[f] ^= 1;   // H, F, S flip once per fused step; V is not stored
    return HG_OK;
}
