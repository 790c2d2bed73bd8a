// hg_fused.cu — the product path: one grid erosion step (Erosion::dispatch_grid,
// src/erosion.cpp:158-200: flux, erosion, sediment transport, thermal x2 layers,
// smoothing) as ONE kernel that reads the nine persistent planes once and writes them
// once (72 B/cell-step) instead of the reference's 8 dispatches and
// >= 512 B/cell-step.
//
// Row-marching software pipeline.  A CTA of NT threads owns a strip of NT-12 columns
// (6 halo columns each side, recomputed) and a segment of rows; thread t holds column
// x0-6+t and marches in +y.  At iteration i every stage works on its own lagged row:
//
//   L(i)    raw row i arrives (prefetched a whole iteration earlier); H.a=(r+g)+b
//   A(i-1)  hydro_flux + hydro_erosion (+ evaporation)  -> F', water' to HBM; rockE, dirtE, S', u, v
//   B(i-3)  sediment back-trace + bilinear gather of S' -> sediment' to HBM
//   C(i-3)  thermal outflow, layer 0 (rock)             -> 8 outflows
//   D(i-5)  thermal transport, layer 0                  -> rock1
//   E(i-7)  thermal outflow, layer 1 (rock1 + dirtE)    -> 8 outflows
//   F(i-9)  thermal transport, layer 1                  -> dirt2
//   G(i-11) smoothing                                   -> rock', dirt' to HBM
//
// A thread keeps its own column's history in registers; values of the x+-1 columns come
// through small shared-memory row rings written one iteration earlier, so ONE
// __syncthreads per row is enough and the seven stages of an iteration are independent
// instruction streams (ILP instead of occupancy).  V, H.a, TC and TD never touch HBM.
// Stage bodies are branch-free (out-of-map cells compute on sentinels and are masked by
// selects); ring slots are per-iteration byte offsets so every shared access is one
// LDS/STS with an immediate.
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

namespace {

constexpr int HX = 6;                    // halo columns per side
constexpr int LAG_G = 11;                // rows between L and G

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

// shared-memory row rings, in rows of NT floats
enum {
    R_XA = 0,    // H.a           [2]
    R_XR = 2,    // rock (pre)    [2]
    R_XD = 4,    // dirt (pre)    [2]
    R_XFL = 6,   // fL (pre)      [2]
    R_XFR = 8,   // fR (pre)      [2]
    R_RE = 10,   // rockE         [2]
    R_DE = 12,   // dirtE         [8]
    R_SR = 20,   // S' rock-sed   [4]
    R_SD = 24,   // S' dirt-sed   [4]
    R_O0 = 28,   // layer-0 outflow R,L,RT,LT,RB,LB  [6][2]
    R_O1 = 40,   // layer-1 outflow                   [6][2]
    R_R1 = 52,   // rock1         [8]
    R_D2 = 60,   // dirt2         [4]
    R_TOTAL = 64
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
__global__ void __launch_bounds__(NT, MINB) k_fused_step(const __grid_constant__ FusedArgs A) {
    extern __shared__ float sm_raw[];
    const HgStepParams& P = A.P;
    const int tid = threadIdx.x;
    // thread's own element of ring row 0; one float of padding before row 0 and after the last
    // row lets the x-1 / x+1 reads of the edge threads stay inside the allocation (their
    // results are never consumed)
    float* const st = sm_raw + 1 + tid;
#define SM(row, off) st[(row) * NT + (off)]
    const int strip = blockIdx.x % A.nstrips, segi = blockIdx.x / A.nstrips;
    const int x = strip * (NT - 2 * HX) - HX + tid;
    const bool xin = x >= 0 && x < A.W;
    const bool owned = tid >= HX && tid < NT - HX && x < A.W;
    const int gy0 = A.row0 + segi * A.seg;
    const int gy1 = min(gy0 + A.seg, A.row0 + A.rows);
    const int W = A.W, H = A.H;

    // ---- per-thread rolling state (own column) ----
    float rk0 = 0, rk1 = 0, rk2 = 0, dt0 = 0, dt1 = 0, dt2 = 0;       // rock, dirt rows i-2,i-1,i
    float at0 = HG_OOB_HEIGHT, at1 = HG_OOB_HEIGHT, at2 = HG_OOB_HEIGHT;   // H.a
    float w1 = 0, w2 = 0;                                                // water rows i-1, i
    float f1L = 0, f1R = 0, f1T = 0, f1B = 0, f2L = 0, f2R = 0, f2T = 0, f2B = 0, f0T = 0;
    float s1r = 0, s1d = 0, s2r = 0, s2d = 0;
    float u_d1 = 0, v_d1 = 0, u_d2 = 0, v_d2 = 0;                        // velocity delayed 1, 2 iterations
    float e00 = 0, e01 = 0, e02 = 0, e10 = 0, e11 = 0, e12 = 0, e20 = 0, e21 = 0, e22 = 0;   // rockE rows i-4..i-2
    float so0_d1 = 0, so0_d2 = 0, T0_d1 = 0, T0_d2 = 0, T0_d3 = 0, B0_d1 = 0;
    float nR0_d1 = 0, nL0_d1 = 0, nRT0_d1 = 0, nRT0_d2 = 0, nLT0_d1 = 0, nLT0_d2 = 0;
    float p00 = 0, p01 = 0, p02 = 0, p10 = 0, p11 = 0, p12 = 0, p20 = 0, p21 = 0, p22 = 0;   // rock1 rows i-8..i-6
    float q00 = 0, q01 = 0, q02 = 0, q10 = 0, q11 = 0, q12 = 0, q20 = 0, q21 = 0, q22 = 0;   // dirtE rows i-8..i-6
    float so1_d1 = 0, so1_d2 = 0, T1_d1 = 0, T1_d2 = 0, T1_d3 = 0, B1_d1 = 0;
    float nR1_d1 = 0, nL1_d1 = 0, nRT1_d1 = 0, nRT1_d2 = 0, nLT1_d1 = 0, nLT1_d2 = 0;
    float g_r0 = 0, g_r1 = 0, g_r2 = 0, g_d0 = 0, g_d1 = 0, g_d2 = 0;   // own column of rock1 / dirt2 rows i-12..i-10
    float pf[HG_NPLANES];

    const int i_begin = gy0 - HX, i_end = gy1 + LAG_G - 1;   // inclusive
    // element offset of (row i, column x) in a plane, advanced by one row per iteration
    long long gidx = (long long)(i_begin - A.row0 + HG_HALO_ROWS) * A.pitch + x;
    {
        const bool ld = xin && i_begin >= 0 && i_begin < H;
#pragma unroll
        for (int p = 0; p < HG_NPLANES; p++) pf[p] = ld ? __ldg(A.src[p] + gidx) : 0.0f;
    }

    for (int i = i_begin; i <= i_end; i++, gidx += A.pitch) {
        // ring slot offsets of this iteration (warp-uniform)
        const int p0 = (i & 1) * NT, p1 = NT - p0;          // slot of row i / of rows i-1, i-3, ...
        const int a3 = ((i - 1) & 3) * NT, a7 = ((i - 1) & 7) * NT;
        const int e7 = ((i - 6) & 7) * NT, d7 = ((i - 5) & 7) * NT, f3 = ((i - 9) & 3) * NT;
        const int g7 = ((i - 10) & 7) * NT, g3 = ((i - 10) & 3) * NT;

        // ------------------------------------------------------------ L(i)
        rk0 = rk1; rk1 = rk2; dt0 = dt1; dt1 = dt2; at0 = at1; at1 = at2; w1 = w2;
        f0T = f1T; f1L = f2L; f1R = f2R; f1T = f2T; f1B = f2B; s1r = s2r; s1d = s2d;
        rk2 = pf[PL_ROCK]; dt2 = pf[PL_DIRT]; w2 = pf[PL_WATER];
        f2L = pf[PL_FL]; f2R = pf[PL_FR]; f2T = pf[PL_FT]; f2B = pf[PL_FB];
        s2r = pf[PL_SR]; s2d = pf[PL_SD];
        at2 = (xin && i >= 0 && i < H) ? rk2 + dt2 + w2 : HG_OOB_HEIGHT;
        SM(R_XA, p0) = at2;
        SM(R_XR, p0) = rk2;
        SM(R_XD, p0) = dt2;
        SM(R_XFL, p0) = f2L;
        SM(R_XFR, p0) = f2R;
        // prefetch raw row i+1 (consumed next iteration)
        {
            const int gy = i + 1;
            const bool ld = xin && gy >= 0 && gy < H && gy < gy1 + HX;
            const long long nidx = gidx + A.pitch;
#pragma unroll
            for (int p = 0; p < HG_NPLANES; p++) pf[p] = ld ? __ldg(A.src[p] + nidx) : 0.0f;
        }

        // ------------------------------------------------------------ A(i-1)
        float u_new = 0.0f, v_new = 0.0f;
        {
            const int ya = i - 1;
            if (ya >= gy0 - 5 && ya < gy1 + 5) {
                const bool in = xin && ya >= 0 && ya < H;
                float aL = SM(R_XA, p1 - 1), aR = SM(R_XA, p1 + 1);
                float rL = SM(R_XR, p1 - 1), rR = SM(R_XR, p1 + 1);
                float gL = SM(R_XD, p1 - 1), gR = SM(R_XD, p1 + 1);
                float inL = SM(R_XFR, p1 - 1), inR = SM(R_XFL, p1 + 1);
                HgFluxOut o = hg_flux_cell(P, x, ya, W, H, at1, aL, aR, at2, at0,
                                           f1L, f1R, f1T, f1B, inL, inR, f2B, f0T, w1);
                HgEroOut e = hg_erosion_cell(P, rk1, dt1, s1r, s1d, o.u, o.v, o.vz,
                                             rR, gR, rL, gL, rk0, dt0, rk2, dt2);
                u_new = o.u; v_new = o.v;
                if (owned && in && ya >= gy0 && ya < gy1) {
                    const long long idx = gidx - A.pitch;
                    A.dst[PL_FL][idx] = o.fL; A.dst[PL_FR][idx] = o.fR;
                    A.dst[PL_FT][idx] = o.fT; A.dst[PL_FB][idx] = o.fB;
                    A.dst[PL_WATER][idx] = o.water * P.evap;     // sediment_transport.glsl:75
                }
                SM(R_RE, p1) = in ? e.rock : HG_OOB_HEIGHT;
                SM(R_DE, a7) = in ? e.dirt : HG_OOB_HEIGHT;
                SM(R_SR, a3) = in ? e.sr : 0.0f;
                SM(R_SD, a3) = in ? e.sd : 0.0f;
            }
        }

        // ------------------------------------------------------------ B(i-3)
        {
            const int yb = i - 3;
            if (yb >= gy0 && yb < gy1) {
                HgBack b = hg_backtrace(P, x, yb, W, H, u_d2, v_d2);
                const int dx = b.px - x, dy = b.py - yb;
                const bool fast = dx >= -1 && dx <= 0 && dy >= -1 && dy <= 0;
                const int cdx = fast ? dx : 0;
                const int r0 = ((fast ? b.py : yb) & 3) * NT, r1 = (((fast ? b.py : yb) + 1) & 3) * NT;
                float sr = hg_bilerp(SM(R_SR, r0 + cdx), SM(R_SR, r0 + cdx + 1), SM(R_SR, r1 + cdx), SM(R_SR, r1 + cdx + 1), b.sx, b.sy);
                float sd = hg_bilerp(SM(R_SD, r0 + cdx), SM(R_SD, r0 + cdx + 1), SM(R_SD, r1 + cdx), SM(R_SD, r1 + cdx + 1), b.sx, b.sy);
                if (owned) {
                    const long long idx = gidx - 3 * (long long)A.pitch;
                    if (fast) {
                        A.dst[PL_SR][idx] = sr;
                        A.dst[PL_SD][idx] = sd;
                    } else {
                        unsigned long long slot = atomicAdd(A.far_count, 1ull);
                        A.far_list[slot] = (unsigned)(yb - A.row0) * (unsigned)W + (unsigned)x;
                    }
                }
            }
        }

        // ------------------------------------------------------------ C(i-3), D(i-5)
        {
            const float rockE_d = e01;     // rockE of row i-5 leaves the window now; D needs it
            e00 = e10; e01 = e11; e02 = e12; e10 = e20; e11 = e21; e12 = e22;
            e20 = SM(R_RE, p0 - 1); e21 = SM(R_RE, p0); e22 = SM(R_RE, p0 + 1);
            const int yc = i - 3;
            float so0 = 0.0f, T0 = 0.0f, B0 = 0.0f;
            if (yc >= gy0 - 4 && yc < gy1 + 4) {
                const bool in = xin && yc >= 0 && yc < H;
                float out[8], d_h[8];
                // L R T B LT RT LB RB; window rows: 0 = y-1, 1 = y, 2 = y+1
                d_h[0] = 0.0f; d_h[0] += e11 - e10;
                d_h[1] = 0.0f; d_h[1] += e11 - e12;
                d_h[2] = 0.0f; d_h[2] += e11 - e21;
                d_h[3] = 0.0f; d_h[3] += e11 - e01;
                d_h[4] = 0.0f; d_h[4] += e11 - e20;
                d_h[5] = 0.0f; d_h[5] += e11 - e22;
                d_h[6] = 0.0f; d_h[6] += e11 - e00;
                d_h[7] = 0.0f; d_h[7] += e11 - e02;
                if (!in) {
#pragma unroll
                    for (int k = 0; k < 8; k++) d_h[k] = -1.0f;   // an out-of-map cell has no outflow
                }
                so0 = hg_thermal_outflow(P, 0, e11, d_h, out);
                T0 = out[2]; B0 = out[3];
                SM(R_O0 + 0, p1) = out[1];    // R
                SM(R_O0 + 2, p1) = out[0];    // L
                SM(R_O0 + 4, p1) = out[5];    // RT
                SM(R_O0 + 6, p1) = out[4];    // LT
                SM(R_O0 + 8, p1) = out[7];    // RB
                SM(R_O0 + 10, p1) = out[6];   // LB
            }
            // D(i-5): neighbours' outflow of row i-4 (written last iteration)
            const int yd = i - 5;
            const float nR = SM(R_O0 + 0, p0 - 1), nL = SM(R_O0 + 2, p0 + 1);
            const float nRT = SM(R_O0 + 4, p0 - 1), nLT = SM(R_O0 + 6, p0 + 1);
            const float nRB = SM(R_O0 + 8, p0 - 1), nLB = SM(R_O0 + 10, p0 + 1);
            if (yd >= gy0 - 3 && yd < gy1 + 3) {
                const bool in = xin && yd >= 0 && yd < H;
                float delta = hg_thermal_delta(so0_d2, nR0_d1, nL0_d1, B0_d1, T0_d3, nRB, nLB, nRT0_d2, nLT0_d2);
                SM(R_R1, d7) = in ? rockE_d + delta : HG_OOB_HEIGHT;
            }
            so0_d2 = so0_d1; so0_d1 = so0;
            T0_d3 = T0_d2; T0_d2 = T0_d1; T0_d1 = T0;
            B0_d1 = B0;
            nR0_d1 = nR; nL0_d1 = nL;
            nRT0_d2 = nRT0_d1; nRT0_d1 = nRT; nLT0_d2 = nLT0_d1; nLT0_d1 = nLT;
        }

        // ------------------------------------------------------------ E(i-7), F(i-9)
        {
            const float dirtE_d = q01;     // dirtE of row i-9
            p00 = p10; p01 = p11; p02 = p12; p10 = p20; p11 = p21; p12 = p22;
            q00 = q10; q01 = q11; q02 = q12; q10 = q20; q11 = q21; q12 = q22;
            p20 = SM(R_R1, e7 - 1); p21 = SM(R_R1, e7); p22 = SM(R_R1, e7 + 1);
            q20 = SM(R_DE, e7 - 1); q21 = SM(R_DE, e7); q22 = SM(R_DE, e7 + 1);
            const int ye = i - 7;
            float so1 = 0.0f, T1 = 0.0f, B1 = 0.0f;
            if (ye >= gy0 - 2 && ye < gy1 + 2) {
                const bool in = xin && ye >= 0 && ye < H;
                float out[8], d_h[8];
                d_h[0] = 0.0f; d_h[0] += p11 - p10; d_h[0] += q11 - q10;
                d_h[1] = 0.0f; d_h[1] += p11 - p12; d_h[1] += q11 - q12;
                d_h[2] = 0.0f; d_h[2] += p11 - p21; d_h[2] += q11 - q21;
                d_h[3] = 0.0f; d_h[3] += p11 - p01; d_h[3] += q11 - q01;
                d_h[4] = 0.0f; d_h[4] += p11 - p20; d_h[4] += q11 - q20;
                d_h[5] = 0.0f; d_h[5] += p11 - p22; d_h[5] += q11 - q22;
                d_h[6] = 0.0f; d_h[6] += p11 - p00; d_h[6] += q11 - q00;
                d_h[7] = 0.0f; d_h[7] += p11 - p02; d_h[7] += q11 - q02;
                if (!in) {
#pragma unroll
                    for (int k = 0; k < 8; k++) d_h[k] = -1.0f;
                }
                so1 = hg_thermal_outflow(P, 1, q11, d_h, out);
                T1 = out[2]; B1 = out[3];
                SM(R_O1 + 0, p1) = out[1];
                SM(R_O1 + 2, p1) = out[0];
                SM(R_O1 + 4, p1) = out[5];
                SM(R_O1 + 6, p1) = out[4];
                SM(R_O1 + 8, p1) = out[7];
                SM(R_O1 + 10, p1) = out[6];
            }
            const int yf = i - 9;
            const float nR = SM(R_O1 + 0, p0 - 1), nL = SM(R_O1 + 2, p0 + 1);
            const float nRT = SM(R_O1 + 4, p0 - 1), nLT = SM(R_O1 + 6, p0 + 1);
            const float nRB = SM(R_O1 + 8, p0 - 1), nLB = SM(R_O1 + 10, p0 + 1);
            if (yf >= gy0 - 1 && yf < gy1 + 1) {
                const bool in = xin && yf >= 0 && yf < H;
                float delta = hg_thermal_delta(so1_d2, nR1_d1, nL1_d1, B1_d1, T1_d3, nRB, nLB, nRT1_d2, nLT1_d2);
                SM(R_D2, f3) = in ? dirtE_d + delta : HG_OOB_HEIGHT;
            }
            so1_d2 = so1_d1; so1_d1 = so1;
            T1_d3 = T1_d2; T1_d2 = T1_d1; T1_d1 = T1;
            B1_d1 = B1;
            nR1_d1 = nR; nL1_d1 = nL;
            nRT1_d2 = nRT1_d1; nRT1_d1 = nRT; nLT1_d2 = nLT1_d1; nLT1_d1 = nLT;
        }

        // ------------------------------------------------------------ G(i-11)
        {
            // own column of rock1 / dirt2: rows i-12, i-11, i-10 (row i-10 was written last iteration)
            g_r0 = g_r1; g_r1 = g_r2; g_r2 = SM(R_R1, g7);
            g_d0 = g_d1; g_d1 = g_d2; g_d2 = SM(R_D2, g3);
            const int yg = i - LAG_G;
            if (yg >= gy0 && yg < gy1) {
                const int s7 = ((yg) & 7) * NT, s3 = ((yg) & 3) * NT;
                float rock = g_r1, dirt = g_d1;
                float sr_ = rock, sd_ = dirt;
                hg_smooth_cell(P, sr_, sd_, SM(R_R1, s7 - 1), SM(R_D2, s3 - 1), SM(R_R1, s7 + 1), SM(R_D2, s3 + 1),
                               g_r2, g_d2, g_r0, g_d0);
                const bool border = (x == 0 || yg == 0 || x == W - 1 || yg == H - 1);
                if (owned) {
                    const long long idx = gidx - LAG_G * (long long)A.pitch;
                    A.dst[PL_ROCK][idx] = border ? rock : sr_;
                    A.dst[PL_DIRT][idx] = border ? dirt : sd_;
                }
            }
        }

        u_d2 = u_d1; v_d2 = v_d1; u_d1 = u_new; v_d1 = v_new;
        __syncthreads();
    }
#undef SM
}

}  // namespace

int hg_launch_fused_step(hg_ctx* c) {
    constexpr int NT = 128;
    FusedArgs A;
    memset(&A, 0, sizeof(A));
    for (int p = 0; p < HG_NPLANES; p++) {
        A.src[p] = hg_cur(c, p, 1);
        A.dst[p] = hg_cur(c, p, 0);
        A.src_set[p] = c->ri[hg_field_of_plane(p)];
    }
    A.W = c->g.W; A.H = c->g.H; A.row0 = c->g.row0; A.rows = c->g.rows; A.pitch = c->g.pitch;
    A.nstrips = (c->g.W + (NT - 2 * HX) - 1) / (NT - 2 * HX);
    // rows per CTA: enough CTAs to fill 148 SMs a few times over, long enough to amortise the 17-row pipeline fill
    int seg = 128;
    while (seg > 32 && (long long)A.nstrips * ((c->g.rows + seg - 1) / seg) < 148 * 4) seg /= 2;
    A.seg = seg;
    if (c->slabs.n > 0) {
        A.slabs = c->slabs;
    } else {
        A.slabs.n = 1; A.slabs.me = 0;
        A.slabs.arena[0] = c->arena; A.slabs.row0[0] = c->g.row0; A.slabs.rows[0] = c->g.rows;
    }
    if (!c->far_list) {   // one entry per owned cell: correct even if every back-trace is far
        HG_CUDA(cudaMalloc(&c->far_list, (size_t)c->g.rows * c->g.W * sizeof(unsigned)));
    }
    A.far_list = c->far_list;
    A.far_count = c->d_counters + 8 + (c->far_parity & 1);
    A.far_count_next = c->d_counters + 8 + ((c->far_parity + 1) & 1);
    A.far_total = c->d_counters;
    c->far_parity ^= 1;
    A.P = c->sp;
    int nseg = (c->g.rows + seg - 1) / seg;
    size_t smem = ((size_t)R_TOTAL * NT + 2) * sizeof(float);
    static_assert(((size_t)R_TOTAL * NT + 2) * sizeof(float) <= 48 * 1024, "raise the dynamic shared memory limit for larger CTAs");
    if (c->prof_ev0) HG_CUDA(cudaEventRecord(c->prof_ev0, c->stream));
    k_fused_step<NT, 4><<<A.nstrips * nseg, NT, smem, c->stream>>>(A);
    HG_LAUNCH_CHECK(c);
    if (c->prof_ev1) HG_CUDA(cudaEventRecord(c->prof_ev1, c->stream));
    k_far_fixup<<<148 * 2, 128, 0, c->stream>>>(A);
    HG_LAUNCH_CHECK(c);
    for (int f = 0; f < 4; f++) if (f != 2) c->ri[f] ^= 1;   // H, F, S flip once per fused step; V is not stored
    return HG_OK;
}
