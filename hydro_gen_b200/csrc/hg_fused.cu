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
#include "hg_internal.cuh"
#include "hg_fused_body.cuh"

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
    HgStepParams P;
};

// ------------------------------------------------------------------ far-fetch path
__device__ __forceinline__ const float* far_plane(const FusedArgs& A, int plane, int gy, size_t* idx_row) {
    const HgSlabTable& T = A.slabs;
    int k = T.me;
    if (gy < T.row0[k] - HG_HALO_ROWS || gy >= T.row0[k] + T.rows[k] + HG_HALO_ROWS) {
        for (int j = 0; j < T.n; j++)
            if (gy >= T.row0[j] && gy < T.row0[j] + T.rows[j]) { k = j; break; }
    }
    size_t pe = (size_t)(T.rows[k] + 2 * HG_HALO_ROWS) * A.pitch;
    *idx_row = (size_t)(gy - T.row0[k] + HG_HALO_ROWS) * A.pitch;
    return T.arena[k] + ((size_t)A.src_set[plane] * HG_NPLANES + plane) * pe;
}
__device__ __forceinline__ float far_ld(const FusedArgs& A, int plane, int x, int gy, float oobv) {
    if (x < 0 || x > A.W - 1 || gy < 0 || gy > A.H - 1) return oobv;
    size_t r;
    const float* p = far_plane(A, plane, gy, &r);
    return __ldcg(p + r + x);
}
// Stage A of any in-map cell recomputed from the pre-step planes: S' and the velocity.
// An out-of-map texel is texelFetch's 0.
__device__ __noinline__ void far_stage_a(const FusedArgs& A, int x, int gy, float* sr, float* sd, float* u, float* v) {
    if (x < 0 || x > A.W - 1 || gy < 0 || gy > A.H - 1) { *sr = 0.0f; *sd = 0.0f; *u = 0.0f; *v = 0.0f; return; }
    float rk[5], dt[5], at[5];   // own, L, R, T, B
    const int ox[5] = {0, -1, 1, 0, 0}, oy[5] = {0, 0, 0, 1, -1};
    float water = 0.0f;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int xx = x + ox[k], yy = gy + oy[k];
        bool in = !(xx < 0 || xx > A.W - 1 || yy < 0 || yy > A.H - 1);
        float r = far_ld(A, PL_ROCK, xx, yy, 0.0f), d = far_ld(A, PL_DIRT, xx, yy, 0.0f), w = far_ld(A, PL_WATER, xx, yy, 0.0f);
        rk[k] = r; dt[k] = d;
        at[k] = in ? r + d + w : HG_OOB_HEIGHT;
        if (k == 0) water = w;
    }
    HgFluxOut o = hg_flux_cell(A.P, x, gy, A.W, A.H, at[0], at[1], at[2], at[3], at[4],
        far_ld(A, PL_FL, x, gy, 0.0f), far_ld(A, PL_FR, x, gy, 0.0f), far_ld(A, PL_FT, x, gy, 0.0f), far_ld(A, PL_FB, x, gy, 0.0f),
        far_ld(A, PL_FR, x - 1, gy, 0.0f), far_ld(A, PL_FL, x + 1, gy, 0.0f),
        far_ld(A, PL_FB, x, gy + 1, 0.0f), far_ld(A, PL_FT, x, gy - 1, 0.0f), water);
    HgEroOut e = hg_erosion_cell(A.P, rk[0], dt[0], far_ld(A, PL_SR, x, gy, 0.0f), far_ld(A, PL_SD, x, gy, 0.0f),
        o.u, o.v, o.vz, rk[2], dt[2], rk[1], dt[1], rk[4], dt[4], rk[3], dt[3]);
    *sr = e.sr; *sd = e.sd; *u = o.u; *v = o.v;
}

// One thread per listed cell: the sediment pass of sediment_transport.glsl:66-93 with every
// texel recomputed from pre-step state.
__global__ void __launch_bounds__(128) k_far_fixup(const __grid_constant__ FusedArgs A) {
    const unsigned long long n = *A.far_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *A.far_count_next = 0ull;
        if (n) atomicAdd(A.far_total, n);
    }
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned li = A.far_list[e];
        int ly = (int)(li / (unsigned)A.W), x = (int)(li - (unsigned)ly * (unsigned)A.W);
        int gy = A.row0 + ly;
        float s0, s1, u, v;
        far_stage_a(A, x, gy, &s0, &s1, &u, &v);
        HgBack b = hg_backtrace(A.P, x, gy, A.W, A.H, u, v);
        float t00r, t00d, t10r, t10d, t01r, t01d, t11r, t11d, du, dv;
        far_stage_a(A, b.px, b.py, &t00r, &t00d, &du, &dv);
        far_stage_a(A, b.px + 1, b.py, &t10r, &t10d, &du, &dv);
        far_stage_a(A, b.px, b.py + 1, &t01r, &t01d, &du, &dv);
        far_stage_a(A, b.px + 1, b.py + 1, &t11r, &t11d, &du, &dv);
        size_t idx = (size_t)(ly + HG_HALO_ROWS) * A.pitch + x;
        A.dst[PL_SR][idx] = hg_bilerp(t00r, t10r, t01r, t11r, b.sx, b.sy);
        A.dst[PL_SD][idx] = hg_bilerp(t00d, t10d, t01d, t11d, b.sx, b.sy);
    }
}

// ------------------------------------------------------------------ main kernel
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_fused_step(const __grid_constant__ HgFusedK K) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int strip = blockIdx.x % K.nstrips, segi = blockIdx.x / K.nstrips;
    const int x = strip * (NT - 2 * HGF_HX) - HGF_HX + tid;
    const bool xin = x >= 0 && x < K.W;
    const bool owned = tid >= HGF_HX && tid < NT - HGF_HX && x < K.W;
    const int gy0 = K.row0 + segi * K.seg;
    const int gy1 = min(gy0 + K.seg, K.row0 + K.rows);
    const HgFusedPlan pl = hg_fused_plan(gy0, gy1, K.H);
    const unsigned pitch = (unsigned)K.pitch;
    // element offset of (row i, column x) in a plane, advanced by one row per iteration
    unsigned off = (unsigned)(pl.i_begin - K.row0 + HG_HALO_ROWS) * pitch + (unsigned)x;
    HgCol c;
    hg_fused_begin(c, K, xin, pl.i_begin, off);
    int i = pl.i_begin;
    for (; i < pl.free_lo && i <= pl.i_end; i++, off += pitch) {
        hg_fused_iter<NT, false>(c, sm, K, tid, x, xin, owned, gy0, gy1, i, off);
        __syncthreads();
    }
#pragma unroll 1
    for (; i <= pl.free_hi; i++, off += pitch) {
        hg_fused_iter<NT, true>(c, sm, K, tid, x, xin, owned, gy0, gy1, i, off);
        __syncthreads();
    }
    for (; i <= pl.i_end; i++, off += pitch) {
        hg_fused_iter<NT, false>(c, sm, K, tid, x, xin, owned, gy0, gy1, i, off);
        __syncthreads();
    }
}

}  // namespace

// NT threads per CTA; seg rows per CTA.  Longer segments amortise the 17-row pipeline fill and
// the generic (non-FREE) iterations; enough CTAs must remain to fill 148 SMs x resident CTAs.
template <int NT, int MINB>
static int launch_main(hg_ctx* c, const HgFusedK& K0, int seg) {
    HgFusedK K = K0;
    K.nstrips = (c->g.W + (NT - 2 * HGF_HX) - 1) / (NT - 2 * HGF_HX);
    K.seg = seg;
    int nseg = (c->g.rows + seg - 1) / seg;
    constexpr size_t smem = (size_t)HgRings<NT>::TOTAL * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        HG_CUDA(cudaFuncSetAttribute(k_fused_step<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    if (c->prof_ev0) HG_CUDA(cudaEventRecord(c->prof_ev0, c->stream));
    k_fused_step<NT, MINB><<<K.nstrips * nseg, NT, smem, c->stream>>>(K);
    HG_LAUNCH_CHECK(c);
    if (c->prof_ev1) HG_CUDA(cudaEventRecord(c->prof_ev1, c->stream));
    return HG_OK;
}

int hg_launch_fused_step(hg_ctx* c) {
    if (c->g.plane_elems >= (size_t)1 << 32) { hg_set_error("slab too large for 32-bit plane offsets (%zu elements)", c->g.plane_elems); return HG_ERR_INVALID; }
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
    if (!c->far_list) {   // one entry per owned cell: correct even if every back-trace is far
        HG_CUDA(cudaMalloc(&c->far_list, (size_t)c->g.rows * c->g.W * sizeof(unsigned)));
    }
    A.far_list = K.far_list = c->far_list;
    A.far_count = K.far_count = c->d_counters + 8 + (c->far_parity & 1);
    A.far_count_next = c->d_counters + 8 + ((c->far_parity + 1) & 1);
    A.far_total = c->d_counters;
    c->far_parity ^= 1;
    A.P = K.P = c->sp;
    // CTA shape (threads, resident CTAs per SM) and rows per CTA; HG_FUSED_VARIANT / HG_FUSED_SEG override (tuning aids)
    static const int nt_of[] = {128, 128, 192, 256, 256};
    static const int res_of[] = {4, 3, 2, 2, 1};
    int v = c->tune_variant >= 0 && c->tune_variant < 5 ? c->tune_variant : 1;
    const int NT = nt_of[v];
    int nstrips = (c->g.W + (NT - 2 * HGF_HX) - 1) / (NT - 2 * HGF_HX);
    int seg = c->tune_seg > 0 ? c->tune_seg : 512;
    if (c->tune_seg <= 0)   // as long as possible while the grid still fills the resident slots about twice
        while (seg > 32 && (long long)nstrips * ((c->g.rows + seg - 1) / seg) < 148 * res_of[v] * 2) seg /= 2;
    int rc;
    switch (v) {
    case 0: rc = launch_main<128, 4>(c, K, seg); break;
    case 1: rc = launch_main<128, 3>(c, K, seg); break;
    case 2: rc = launch_main<192, 2>(c, K, seg); break;
    case 3: rc = launch_main<256, 2>(c, K, seg); break;
    default: rc = launch_main<256, 1>(c, K, seg); break;
    }
    if (rc) return rc;
    k_far_fixup<<<148 * 2, 128, 0, c->stream>>>(A);
    HG_LAUNCH_CHECK(c);
    for (int f = 0; f < 4; f++) if (f != 2) c->ri[f] ^= 1;   // H, F, S flip once per fused step; V is not stored
    return HG_OK;
}
