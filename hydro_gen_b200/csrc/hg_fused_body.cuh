// hg_fused_body.cuh — one row iteration of the fused grid erosion step, for one thread.
//
// Erosion::dispatch_grid (src/erosion.cpp:158-200: flux, erosion, sediment transport,
// thermal x2 layers, smoothing) as a row-marching software pipeline.  A CTA of NT threads
// owns a strip of NT-12 columns (6 halo columns each side, recomputed) and a segment of
// rows; thread t holds column x0-6+t and marches in +y.  At iteration i every stage works
// on its own lagged row:
//
//   L(i)    raw row i (TMA-staged in shared memory one iteration earlier); H.a = (r+g)+b
//   A(i-1)  hydro_flux + hydro_erosion (+ evaporation)  -> F', water' to HBM; rockE, dirtE, S', u, v
//   B(i-3)  sediment back-trace + bilinear gather of S' -> sediment' to HBM
//   C(i-3)  thermal outflow, layer 0 (rock)             -> 8 outflows
//   D(i-5)  thermal transport, layer 0                  -> rock1
//   E(i-7)  thermal outflow, layer 1 (rock1 + dirtE)    -> 8 outflows
//   F(i-9)  thermal transport, layer 1                  -> dirt2
//   G(i-11) smoothing                                   -> rock', dirt' to HBM
//
// A thread keeps the history of the raw row, the velocity and the outflow sums of its own
// column in registers (HgCol); values of the x+-1 columns come through shared-memory row rings
// written at least one iteration earlier, so ONE barrier per row is enough and the seven stages
// of an iteration are independent instruction streams.  The 3x3 windows of the thermal and
// smoothing stages (rockE; rock1/dirtE; rock1/dirt2) are NOT held in registers: their rings are
// four rows deep and the nine texels are re-read every row, which costs no more instructions than
// rotating a register window (3 loads + 6 moves) and frees 33 registers per thread.  Ring rows are indexed by (absolute row mod N), N a power of two.
// FREE=true is the steady state: every stage's row lies inside its active range and
// strictly inside the map in y, so no range or y-border test is left; pipeline fill/drain
// and the rows next to the map border use FREE=false, which tests everything.  The
// steady-state body must stay inside the 32 KB L1.5 instruction cache: a 6x unrolled
// version with compile-time ring slots executed 35 % fewer instructions and ran slower
// (profiles/r01b: no_instruction stalls 0.3 -> 2.6 per issue).
//
// Exchanges are packed so one LDS/STS moves what a neighbour needs:
//   XQ  float4 (H.a, rock, dirt, fR)   XL float fL          raw row            N=2
//   RD  float2 (rockE, dirtE)                               after stage A      N=4
//   SS  float2 (S'.rock, S'.dirt)                           after stage A      N=4
//   OxR float4 (R, RT, RB, -)  OxL float4 (L, LT, LB, -)    thermal outflow    N=2, per layer
//   R1D float2 (rock1, dirtE)                               after stage D      N=4
//   G2  float2 (rock1, dirt2)                               after stage F      N=4
//
// Every hand-off between the hydraulic stages (L, A, B) and the thermal/smoothing stages (C..G)
// already goes through a ring with a lag of at least one row: rockE via RE, dirtE via DE.  The
// body can therefore be split between two warp groups of one CTA that share the rings and the
// one barrier per row (GROUP = HGF_HYDRO / HGF_THERMAL, k_fused_ws): each group carries only its
// own half of the column history, which halves the registers a thread needs and lets half as
// many more warps stay resident.  GROUP = HGF_ALL is the single-group form (k_fused_step).
//
// This header is plain C++ apart from the HGF_* macros, so tests/host_emul runs the very
// same body thread by thread on the CPU (-m "not gpu") against the oracle.
#pragma once
#include "hg_cell.cuh"

#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define HGF_LDG(p) __ldg(p)
#define HGF_LDG4(p) hg_ldg_f4(p)
#define HGF_ATOMIC_INC64(p) atomicAdd((p), 1ull)
#else
#define HGF_LDG(p) (*(p))
#define HGF_LDG4(p) (*(p))
#define HGF_ATOMIC_INC64(p) ((*(p))++)
#endif

// which stages a warp group runs: everything / L, A, B / C..G / C, D (thermal layer 0) / E, F (thermal layer 1).
// Layer 1 reads layer 0's result only through the R1D ring, so the thermal stages split between two groups the same
// way the hydraulic and thermal stages do (k_fused_ws3: three warp groups per CTA).
constexpr int HGF_ALL = 0, HGF_HYDRO = 1, HGF_THERMAL = 2, HGF_THERMAL0 = 3, HGF_THERMAL1 = 4;
#ifndef HGF_SMOOTH_GROUP
#define HGF_SMOOTH_GROUP HGF_THERMAL
#endif
constexpr int HGF_HX = 6;        // halo columns per side
constexpr int HGF_LAG_G = 11;    // rows between L and G
constexpr int HGF_NPL = 9;       // rock dirt water fL fR fT fB sr sd (HgPlane order)
#define HGF_RAW_LD(NT) ((NT) + 4)  // columns per plane row in the TMA-staged raw block

struct alignas(16) HgF4 { float x, y, z, w; };
struct alignas(8) HgF2 { float x, y; };
#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
__device__ __forceinline__ HgF4 hg_ldg_f4(const HgF4* p) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); HgF4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
#endif
#if defined(__CUDACC__)
static_assert(sizeof(HgF4) == 16 && sizeof(HgF2) == 8, "packed ring elements");
#endif

// Ring offsets in BYTES inside the CTA's shared block; every ring row has NT+2 elements (one pad
// element each side so tid-1 / tid+1 of the edge threads stay inside).  A ring is [rows][E]
// elements, so the address of (row slot s, element el) is ring + (s*E + el) * sizeof(element):
// one per-iteration pointer per (element size, slot) serves every ring of that shape, and each
// access is that pointer plus a compile-time offset.
template <int NT> struct HgRings {
    static constexpr int E = NT + 2;
    static constexpr int XQ = 0;                     // float4 [2][E]
    static constexpr int O0R = XQ + 2 * E * 16;      // float4 [2][E]
    static constexpr int O0L = O0R + 2 * E * 16;
    static constexpr int O1R = O0L + 2 * E * 16;
    static constexpr int O1L = O1R + 2 * E * 16;
    static constexpr int SS = O1L + 2 * E * 16;      // float2 [4][E]
    static constexpr int RD = SS + 4 * E * 8;        // float2 [4][E]
    static constexpr int R1D = RD + 4 * E * 8;       // float2 [4][E]
    static constexpr int G2 = R1D + 4 * E * 8;       // float2 [4][E]
    static constexpr int XL = G2 + 4 * E * 8;        // float  [2][E]
    static constexpr int TOTAL_BYTES = XL + 2 * E * 4;
    static constexpr int TOTAL = TOTAL_BYTES / 4;    // 74 * E floats
};

// (a.x - b.x) + (a.y - b.y).  Device: the two subtractions as one FADD2 (add.rn.f32x2 with a negated operand: each lane
// rounded like the scalar instruction, DESIGN.md §3.1 "two columns per thread"), then the scalar sum.
HG_FN float hg_pair_diff_sum(const HgF2& a, const HgF2& b) {
#if defined(__CUDACC__) && defined(__CUDA_ARCH__) && !defined(HG_NO_PACKED_DIFF)
    const float2 d = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return d.x + d.y;
#else
    return (a.x - b.x) + (a.y - b.y);
#endif
}

// hg_smooth_cell on the (rock, dirt) pairs the G2 ring holds.  Device: the eight differences and the two five-point sums
// as packed additions (same operations, same association, each lane rounded like the scalar instruction); everything
// else is hg_smooth_cell's code.  Host: hg_smooth_cell itself.
HG_FN void hg_smooth_pair(const HgStepParams& P, float& rock, float& dirt, const HgF2& own, const HgF2& l, const HgF2& r, const HgF2& t, const HgF2& b) {
#if defined(__CUDACC__) && defined(__CUDA_ARCH__) && !defined(HG_NO_PACKED_DIFF)
    const float2 o = make_float2(own.x, own.y);
    const float2 dl = __fadd2_rn(o, make_float2(-l.x, -l.y)), dr = __fadd2_rn(o, make_float2(-r.x, -r.y));
    const float2 dt = __fadd2_rn(o, make_float2(-t.x, -t.y)), db = __fadd2_rn(o, make_float2(-b.x, -b.y));
    const float dlr = dl.x, drr = dr.x, dtr = dt.x, dbr = db.x;
    const float dlg = dl.y + dlr, drg = dr.y + drr, dtg = dt.y + dtr, dbg = db.y + dbr;
    const float g_hdiff = fabsf((dlg + drg + dtg + dbg) / 4.0f);
    const float r_hdiff = fabsf((dlr + drr + dtr + dbr) / 4.0f);
    const float2 xc = __fmul2_rn(make_float2(dlr, dlg), make_float2(drr, drg)), yc = __fmul2_rn(make_float2(dtr, dtg), make_float2(dbr, dbg));
    const float xcr = xc.x, xcg = xc.y, ycr = yc.x, ycg = yc.y;
    // terr + l + r + t + b for both layers at once
    const float2 s5 = __fadd2_rn(__fadd2_rn(__fadd2_rn(__fadd2_rn(o, make_float2(l.x, l.y)), make_float2(r.x, r.y)), make_float2(t.x, t.y)), make_float2(b.x, b.y));
    float terr_r = own.x, terr_g = own.y;
    if ((((-dlr) > r_hdiff || (-drr) > r_hdiff) && xcr > 0.0f) || (((-dtr) > r_hdiff || (-dbr) > r_hdiff) && ycr > 0.0f)) terr_r = hg_div5(s5.x);
    if ((((-dlg) > g_hdiff || (-drg) > g_hdiff) && xcg > 0.0f) || (((-dtg) > g_hdiff || (-dbg) > g_hdiff) && ycg > 0.0f)) terr_g = hg_div5(s5.y);
    const float m = P.smooth_mul, m1 = 1.0f - m;
    const float2 pa = __fmul2_rn(make_float2(m, m), make_float2(terr_r, terr_g)), pb = __fmul2_rn(make_float2(m1, m1), o);
    rock = __fadd_rn(pa.x, pb.x);      // scalar sums: a packed add of packed products would be fused by ptxas
    dirt = __fadd_rn(pa.y, pb.y);
#else
    rock = own.x; dirt = own.y;
    hg_smooth_cell(P, rock, dirt, l.x, l.y, r.x, r.y, t.x, t.y, b.x, b.y);
#endif
}

struct HgPlanItem { int strip, gy0, gy1, pad; };

struct HgFusedK {
    const float* src[HGF_NPL];
    float* dst[HGF_NPL];
    int W, H, row0, rows, pitch;
    int seg, nstrips;
    unsigned* far_list;                  // local linear cell indices (row - row0) * W + x
    unsigned long long* far_count;       // this step's counter
    // balanced partition (k_fused_ws): CTA b works on strip plan[b].x, rows [plan[b].y, plan[b].z) and reports its
    // duration in cta_ns[b]; null = uniform segments of `seg` rows
    const HgPlanItem* plan;
    unsigned* cta_ns;
    // droplet mode (DROPS): heightmap and momentum map in the reference's texture layout (one 16-byte texel per cell:
    // (rock, dirt, water, total) and (mx, my, acc_x, acc_y)), the layout the droplet kernels gather from and scatter to
    const HgF4* ha_src;
    HgF4* ha_dst;
    const HgF4* ma_src;
    HgF4* ma_dst;
    // row slabs on several GPUs (hg_slab.cu): the rows a neighbour keeps as ghost rows are stored a second time, straight
    // into that neighbour's planes through its peer pointer (NVLink), by the thread that produces them -- the step kernel
    // IS the halo push.  push_mask bit 0: rows < push_lo_end also go to peer[0] (the slab below), bit 1: rows >=
    // push_hi_begin to peer[1] (the slab above); the pointers are pre-offset so that this slab's element index applies.
    float* peer[2][HGF_NPL];
    int push_mask, push_lo_end, push_hi_begin;
    HgStepParams P;
};

// second store of a result row into the neighbours' ghost rows (generic iterations only: the FREE range of a connected
// slab keeps clear of the edge rows, hg_fused_plan_slab)
#define HGF_PEER_STORE(K, p, row, idx, val)                                                                  \
    {                                                                                                        \
        if (((K).push_mask & 1) && (row) < (K).push_lo_end) (K).peer[0][p][idx] = (val);                     \
        if (((K).push_mask & 2) && (row) >= (K).push_hi_begin) (K).peer[1][p][idx] = (val);                  \
    }

// per-thread rolling state (one column)
struct HgCol {
    float rk0, rk1, rk2, dt0, dt1, dt2;          // rock, dirt rows i-2, i-1, i
    float at0, at1, at2;                         // H.a
    float w1, w2;                                // water rows i-1, i
    float f1L, f1R, f1T, f1B, f2L, f2R, f2T, f2B, f0T;
    float s1r, s1d, s2r, s2d;
    float u_d1, v_d1, u_d2, v_d2;                // velocity delayed 1, 2 iterations
    float e_old, de_old;                         // own (rockE, dirtE) of row i-5 (stage D)
    float so0_d1, so0_d2, T0_d1, T0_d2, T0_d3, B0_d1;
    float nR0_d1, nL0_d1, nRT0_d1, nRT0_d2, nLT0_d1, nLT0_d2;
    float p_old, q_old;                          // own (rock1, dirtE) of row i-9 (stage F)
    float so1_d1, so1_d2, T1_d1, T1_d2, T1_d3, B1_d1;
    float nR1_d1, nL1_d1, nRT1_d1, nRT1_d2, nLT1_d1, nLT1_d2;
    float pf_w, pf_m0, pf_m1, pf_m2, pf_m3;      // droplet mode: water and momentum map of the NEXT smoothing row, fetched one iteration ahead
};

HG_FN void hg_col_init(HgCol& c) {
    float* f = reinterpret_cast<float*>(&c);
    for (int k = 0; k < (int)(sizeof(HgCol) / sizeof(float)); k++) f[k] = 0.0f;
    c.at0 = c.at1 = c.at2 = HG_OOB_HEIGHT;
}

// Byte offsets (inside the ring block) of this thread's element in the ring rows of iteration i, carried from one iteration
// to the next and ROTATED instead of being recomputed from i (a 4-cycle and two swaps: 11 moves instead of ~30 integer
// operations per row and group).  a16 / a4: rows i, i-1 of the two-row float4 / float rings; b8: rows i .. i-3 of the four-row
// float2 rings.
struct HgRingOff { int a16_0, a16_1, a4_0, a4_1, b8_0, b8_1, b8_2, b8_3; };
template <int NT> HG_FN HgRingOff hg_ring_off(int i, int tid) {
    typedef HgRings<NT> R;
    const int e = tid + 1;
    const int u0 = (i & 1) * R::E + e, u1 = (R::E + 2 * e) - u0;
    HgRingOff o;
    o.a16_0 = u0 * 16; o.a16_1 = u1 * 16; o.a4_0 = u0 * 4; o.a4_1 = u1 * 4;
    o.b8_0 = ((i & 3) * R::E + e) * 8; o.b8_1 = (((i - 1) & 3) * R::E + e) * 8;
    o.b8_2 = (((i - 2) & 3) * R::E + e) * 8; o.b8_3 = (((i - 3) & 3) * R::E + e) * 8;
    return o;
}
HG_FN void hg_ring_off_next(HgRingOff& o) {      // i -> i + 1: row i+1 takes the slot of row i-1 (two rows) / row i-3 (four rows)
    int t = o.a16_0; o.a16_0 = o.a16_1; o.a16_1 = t;
    t = o.a4_0; o.a4_0 = o.a4_1; o.a4_1 = t;
    t = o.b8_3; o.b8_3 = o.b8_2; o.b8_2 = o.b8_1; o.b8_1 = o.b8_0; o.b8_0 = t;
}

// One iteration.  sm: the CTA's ring block; tid: thread index; x: global column of this
// thread; xin/owned: column inside the map / inside the strip proper; gy0, gy1: the CTA's
// row segment; off: element offset of (row i, column x) inside a plane.
// FREE: see the header.
template <int NT, bool FREE, int GROUP = HGF_ALL, bool DROPS = false, int SG = HGF_SMOOTH_GROUP>
HG_FN void hg_fused_iter(HgCol& c, float* sm, const float* raw, const HgFusedK& K, const int tid, const int x, const bool xin, const bool owned,
                         const int gy0, const int gy1, const int i, const unsigned off, const HgRingOff* ro = nullptr) {
    typedef HgRings<NT> R;
    const HgStepParams& P = K.P;
    const int W = K.W, H = K.H;
    const unsigned pitch = (unsigned)K.pitch;
    const int e = tid + 1;   // element index inside a ring row
    constexpr bool RUN_H = GROUP == HGF_ALL || GROUP == HGF_HYDRO;
    constexpr bool RUN_CD = GROUP == HGF_ALL || GROUP == HGF_THERMAL || GROUP == HGF_THERMAL0;
    constexpr bool RUN_EF = GROUP == HGF_ALL || GROUP == HGF_THERMAL || GROUP == HGF_THERMAL1;
    char* const smc = reinterpret_cast<char*>(sm);
    // Element index (slot * E + e) of the ring row that holds absolute row i - j: two-row rings
    // (j = 0, 1) and four-row rings (j = 0..3).  Row i - k lives in slot (i - k) mod n.
    const HgRingOff rq = ro ? *ro : hg_ring_off<NT>(i, tid);      // carried and rotated by the kernel shells, computed by the CPU emulation
    char* const a16_0 = smc + rq.a16_0; char* const a16_1 = smc + rq.a16_1;     // float4 rings
    char* const a4_0 = smc + rq.a4_0; char* const a4_1 = smc + rq.a4_1;         // float ring (XL)
    char* const b8_0 = smc + rq.b8_0; char* const b8_1 = smc + rq.b8_1; char* const b8_2 = smc + rq.b8_2; char* const b8_3 = smc + rq.b8_3;   // float2 rings
    // ring: byte offset (HgRings); k: the row is i - k; d: element offset relative to this thread's
#define A16(k) (((k) & 1) ? a16_1 : a16_0)
#define A4(k) (((k) & 1) ? a4_1 : a4_0)
#define B8(k) (((k) & 3) == 0 ? b8_0 : ((k) & 3) == 1 ? b8_1 : ((k) & 3) == 2 ? b8_2 : b8_3)
#define Q4(ring, k, d) (*reinterpret_cast<HgF4*>(A16(k) + (ring) + (d) * 16))
#define Q1(ring, k, d) (*reinterpret_cast<float*>(A4(k) + (ring) + (d) * 4))
#define Q2(ring, k, d) (*reinterpret_cast<HgF2*>(B8(k) + (ring) + (d) * 8))
#define Q2X(ring, k, d) (reinterpret_cast<const float*>(B8(k) + (ring) + (d) * 8)[0])   /* .x only: a 4-byte load */

    if (RUN_H && DROPS) {
    // ------------------------------------------------------------ droplet mode: no hydraulics
    // Erosion::dispatch_particle (src/erosion.cpp:146-155) runs thermal x2 + smoothing on the heightmap the droplets
    // left: this group only feeds the thermal group with (rock, dirt) of row i-1, as stage A does with (rockE, dirtE).
    c.rk1 = c.rk2; c.dt1 = c.dt2;
    {
        // raw row i: NT+4 heightmap texels (rock, dirt, water, total), staged by one TMA box load
        const HgF4 t = reinterpret_cast<const HgF4*>(raw)[tid + 2];
        c.rk2 = t.x; c.dt2 = t.y;
    }
    {
        const int ya = i - 1;
        if (FREE || (ya >= gy0 - 5 && ya < gy1 + 5)) {
            const bool in = xin && (FREE || (ya >= 0 && ya < H));
            HgF2 rd; rd.x = in ? c.rk1 : HG_OOB_HEIGHT; rd.y = in ? c.dt1 : HG_OOB_HEIGHT;
            Q2(R::RD, 1, 0) = rd;
        }
    }
    }
    if (RUN_H && !DROPS) {
    // ------------------------------------------------------------ L(i)
    c.rk0 = c.rk1; c.rk1 = c.rk2; c.dt0 = c.dt1; c.dt1 = c.dt2; c.at0 = c.at1; c.at1 = c.at2; c.w1 = c.w2;
    c.f0T = c.f1T; c.f1L = c.f2L; c.f1R = c.f2R; c.f1T = c.f2T; c.f1B = c.f2B; c.s1r = c.s2r; c.s1d = c.s2d;
    // raw row i: nine planes x (NT+4) columns staged in shared memory by one TMA box load; the box
    // starts 2 columns left of thread 0 because TMA needs a 16-byte aligned start.  Out-of-map
    // columns and rows arrive as zeros (TMA bounds fill / never-written ghost rows).
    {
        const float* rw = raw + tid + 2;
        constexpr int LD = HGF_RAW_LD(NT);
        c.rk2 = rw[0 * LD]; c.dt2 = rw[1 * LD]; c.w2 = rw[2 * LD];
        c.f2L = rw[3 * LD]; c.f2R = rw[4 * LD]; c.f2T = rw[5 * LD]; c.f2B = rw[6 * LD];
        c.s2r = rw[7 * LD]; c.s2d = rw[8 * LD];
    }
    c.at2 = (xin && (FREE || (i >= 0 && i < H))) ? c.rk2 + c.dt2 + c.w2 : HG_OOB_HEIGHT;
    {
        HgF4 q; q.x = c.at2; q.y = c.rk2; q.z = c.dt2; q.w = c.f2R;
        Q4(R::XQ, 0, 0) = q;
        Q1(R::XL, 0, 0) = c.f2L;
    }

    // ------------------------------------------------------------ A(i-1)
    float u_new = 0.0f, v_new = 0.0f;
    {
        const int ya = i - 1;
        if (FREE || (ya >= gy0 - 5 && ya < gy1 + 5)) {
            const bool in = xin && (FREE || (ya >= 0 && ya < H));
            const HgF4 ql = Q4(R::XQ, 1, -1);     // left neighbour: H.a, rock, dirt, fR
            const HgF4 qr = Q4(R::XQ, 1, 1);      // right neighbour: H.a, rock, dirt
            const float inR = Q1(R::XL, 1, 1);    // right neighbour's fL
            // FREE rows are strictly inside the map in y: y = 1, H = 4 folds the y-border tests away
            HgFluxOut o = hg_flux_cell(P, x, FREE ? 1 : ya, W, FREE ? 4 : H, c.at1, ql.x, qr.x, c.at2, c.at0,
                                       c.f1L, c.f1R, c.f1T, c.f1B, ql.w, inR, c.f2B, c.f0T, c.w1);
#ifdef HG_EXP_NO_ERO
            HgEroOut er; er.rock = c.rk1 + qr.y * 0.001f; er.dirt = c.dt1 + ql.z * 0.001f; er.sr = c.s1r + o.u * 0.001f; er.sd = c.s1d + o.vz * 0.001f;
#else
            HgEroOut er = hg_erosion_cell(P, c.rk1, c.dt1, c.s1r, c.s1d, o.u, o.v, o.vz,
                                          qr.y, qr.z, ql.y, ql.z, c.rk0, c.dt0, c.rk2, c.dt2);
#endif
            u_new = o.u; v_new = o.v;
            if (owned && in && (FREE || (ya >= gy0 && ya < gy1))) {
                const unsigned idx = off - pitch;
                K.dst[3][idx] = o.fL; K.dst[4][idx] = o.fR;
                K.dst[5][idx] = o.fT; K.dst[6][idx] = o.fB;
                K.dst[2][idx] = o.water * P.evap;     // sediment_transport.glsl:75
                if (!FREE && K.push_mask) {
                    HGF_PEER_STORE(K, 3, ya, idx, o.fL) HGF_PEER_STORE(K, 4, ya, idx, o.fR)
                    HGF_PEER_STORE(K, 5, ya, idx, o.fT) HGF_PEER_STORE(K, 6, ya, idx, o.fB)
                    HGF_PEER_STORE(K, 2, ya, idx, o.water * P.evap)
                }
            }
            HgF2 rd; rd.x = in ? er.rock : HG_OOB_HEIGHT; rd.y = in ? er.dirt : HG_OOB_HEIGHT;
            Q2(R::RD, 1, 0) = rd;
            HgF2 s; s.x = in ? er.sr : 0.0f; s.y = in ? er.sd : 0.0f;
            Q2(R::SS, 1, 0) = s;
        }
    }

    // ------------------------------------------------------------ B(i-3)
    {
        const int yb = i - 3;
#ifndef HG_EXP_NO_B
        if (FREE || (yb >= gy0 && yb < gy1)) {
            HgBack b = hg_backtrace(P, x, yb, W, H, c.u_d2, c.v_d2);
            const int dx = b.px - x, dy = b.py - yb;
            const bool fast = dx >= -1 && dx <= 0 && dy >= -1 && dy <= 0;
            const int cdx = fast ? dx : 0;
            const bool up = fast && dy == -1;          // footprint rows (yb-1, yb) instead of (yb, yb+1)
            const char* const r0 = (up ? B8(4) : B8(3)) + R::SS + cdx * 8;
            const char* const r1 = (up ? B8(3) : B8(2)) + R::SS + cdx * 8;
            const HgF2 t00 = reinterpret_cast<const HgF2*>(r0)[0], t10 = reinterpret_cast<const HgF2*>(r0)[1];
            const HgF2 t01 = reinterpret_cast<const HgF2*>(r1)[0], t11 = reinterpret_cast<const HgF2*>(r1)[1];
            float sr, sd;
            hg_bilerp2(t00.x, t00.y, t10.x, t10.y, t01.x, t01.y, t11.x, t11.y, b.sx, b.sy, sr, sd);
            if (owned) {
                const unsigned idx = off - 3u * pitch;
                if (fast) {
                    K.dst[7][idx] = sr;
                    K.dst[8][idx] = sd;
                    if (!FREE && K.push_mask) { HGF_PEER_STORE(K, 7, yb, idx, sr) HGF_PEER_STORE(K, 8, yb, idx, sd) }
                } else {
                    unsigned long long slot = HGF_ATOMIC_INC64(K.far_count);
                    K.far_list[slot] = (unsigned)(yb - K.row0) * (unsigned)W + (unsigned)x;
                }
            }
        }
#endif
    }

    c.u_d2 = c.u_d1; c.v_d2 = c.v_d1; c.u_d1 = u_new; c.v_d1 = v_new;
    }   // RUN_H

    // ------------------------------------------------------------ C(i-3), D(i-5)
    if (RUN_CD) {
        // rockE window: rows i-4 (y-1), i-3 (y), i-2 (y+1), columns x-1..x+1, straight from the ring
        // (stage A writes row i-1 into the fourth slot meanwhile)
        const HgF2 rd01 = Q2(R::RD, 4, 0);
        const float e00 = Q2X(R::RD, 4, -1), e01 = rd01.x, e02 = Q2X(R::RD, 4, 1);
        const float e10 = Q2X(R::RD, 3, -1), e11 = Q2X(R::RD, 3, 0), e12 = Q2X(R::RD, 3, 1);
        const float e20 = Q2X(R::RD, 2, -1), e21 = Q2X(R::RD, 2, 0), e22 = Q2X(R::RD, 2, 1);
        const float rockE_d = c.e_old, dirtE_d0 = c.de_old;   // own (rockE, dirtE) of row i-5: last iteration's row y-1; D needs them
        c.e_old = rd01.x; c.de_old = rd01.y;
        const int yc = i - 3;
        float so0 = 0.0f, T0 = 0.0f, B0 = 0.0f;
        if (FREE || (yc >= gy0 - 4 && yc < gy1 + 4)) {
            const bool in = xin && (FREE || (yc >= 0 && yc < H));
            float out[8], d_h[8];
            // L R T B LT RT LB RB; window rows: 0 = y-1, 1 = y, 2 = y+1.  The shader's "0 +" is
            // dropped: it can only turn a -0 difference into +0, and a zero d_h is never marked.
            d_h[0] = e11 - e10; d_h[1] = e11 - e12; d_h[2] = e11 - e21; d_h[3] = e11 - e01;
            d_h[4] = e11 - e20; d_h[5] = e11 - e22; d_h[6] = e11 - e00; d_h[7] = e11 - e02;
            so0 = hg_thermal_outflow(P, 0, e11, d_h, out, in);   // an out-of-map cell has no outflow
            T0 = out[2]; B0 = out[3];
            HgF4 tr; tr.x = out[1]; tr.y = out[5]; tr.z = out[7]; tr.w = 0.0f;    // R, RT, RB
            HgF4 tl; tl.x = out[0]; tl.y = out[4]; tl.z = out[6]; tl.w = 0.0f;    // L, LT, LB
            Q4(R::O0R, 3, 0) = tr;
            Q4(R::O0L, 3, 0) = tl;
        }
        // D(i-5): neighbours' outflow of row i-4 (written last iteration)
        const int yd = i - 5;
        const HgF4 nl = Q4(R::O0R, 4, -1);     // left neighbour's R, RT, RB
        const HgF4 nr = Q4(R::O0L, 4, 1);      // right neighbour's L, LT, LB
        if (FREE || (yd >= gy0 - 3 && yd < gy1 + 3)) {
            const bool in = xin && (FREE || (yd >= 0 && yd < H));
            float delta = hg_thermal_delta(c.so0_d2, c.nR0_d1, c.nL0_d1, c.B0_d1, c.T0_d3, nl.z, nr.z, c.nRT0_d2, c.nLT0_d2);
            HgF2 w; w.x = in ? rockE_d + delta : HG_OOB_HEIGHT; w.y = dirtE_d0;
            Q2(R::R1D, 5, 0) = w;
        }
        c.so0_d2 = c.so0_d1; c.so0_d1 = so0;
        c.T0_d3 = c.T0_d2; c.T0_d2 = c.T0_d1; c.T0_d1 = T0;
        c.B0_d1 = B0;
        c.nR0_d1 = nl.x; c.nL0_d1 = nr.x;
        c.nRT0_d2 = c.nRT0_d1; c.nRT0_d1 = nl.y; c.nLT0_d2 = c.nLT0_d1; c.nLT0_d1 = nr.y;
    }

    // ------------------------------------------------------------ E(i-7), F(i-9)
    if (RUN_EF) {
        // (rock1, dirtE) window: rows i-8 (y-1), i-7 (y), i-6 (y+1), columns x-1..x+1, from the ring
        // (stage D writes row i-5 into the fourth slot meanwhile)
        const HgF2 w00 = Q2(R::R1D, 8, -1), w01 = Q2(R::R1D, 8, 0), w02 = Q2(R::R1D, 8, 1);
        const HgF2 w10 = Q2(R::R1D, 7, -1), w11 = Q2(R::R1D, 7, 0), w12 = Q2(R::R1D, 7, 1);
        const HgF2 w20 = Q2(R::R1D, 6, -1), w21 = Q2(R::R1D, 6, 0), w22 = Q2(R::R1D, 6, 1);
        const float rock1_d = c.p_old, dirtE_d = c.q_old;     // own (rock1, dirtE) of row i-9: last iteration's row y-1
        c.p_old = w01.x; c.q_old = w01.y;
        const int ye = i - 7;
        float so1 = 0.0f, T1 = 0.0f, B1 = 0.0f;
        if (FREE || (ye >= gy0 - 2 && ye < gy1 + 2)) {
            const bool in = xin && (FREE || (ye >= 0 && ye < H));
            float out[8], d_h[8];
            // (rock1 - n.rock1) + (dirtE - n.dirtE): the two differences are one packed subtraction on the pair the ring holds
            d_h[0] = hg_pair_diff_sum(w11, w10); d_h[1] = hg_pair_diff_sum(w11, w12);
            d_h[2] = hg_pair_diff_sum(w11, w21); d_h[3] = hg_pair_diff_sum(w11, w01);
            d_h[4] = hg_pair_diff_sum(w11, w20); d_h[5] = hg_pair_diff_sum(w11, w22);
            d_h[6] = hg_pair_diff_sum(w11, w00); d_h[7] = hg_pair_diff_sum(w11, w02);
            so1 = hg_thermal_outflow(P, 1, w11.y, d_h, out, in);
            T1 = out[2]; B1 = out[3];
            HgF4 tr; tr.x = out[1]; tr.y = out[5]; tr.z = out[7]; tr.w = 0.0f;
            HgF4 tl; tl.x = out[0]; tl.y = out[4]; tl.z = out[6]; tl.w = 0.0f;
            Q4(R::O1R, 7, 0) = tr;
            Q4(R::O1L, 7, 0) = tl;
        }
        const int yf = i - 9;
        const HgF4 nl = Q4(R::O1R, 8, -1);
        const HgF4 nr = Q4(R::O1L, 8, 1);
        if (FREE || (yf >= gy0 - 1 && yf < gy1 + 1)) {
            const bool in = xin && (FREE || (yf >= 0 && yf < H));
            float delta = hg_thermal_delta(c.so1_d2, c.nR1_d1, c.nL1_d1, c.B1_d1, c.T1_d3, nl.z, nr.z, c.nRT1_d2, c.nLT1_d2);
            HgF2 w; w.x = rock1_d; w.y = in ? dirtE_d + delta : HG_OOB_HEIGHT;
            Q2(R::G2, 9, 0) = w;
        }
        c.so1_d2 = c.so1_d1; c.so1_d1 = so1;
        c.T1_d3 = c.T1_d2; c.T1_d2 = c.T1_d1; c.T1_d1 = T1;
        c.B1_d1 = B1;
        c.nR1_d1 = nl.x; c.nL1_d1 = nr.x;
        c.nRT1_d2 = c.nRT1_d1; c.nRT1_d1 = nl.y; c.nLT1_d2 = c.nLT1_d1; c.nLT1_d1 = nr.y;
    }

    // ------------------------------------------------------------ G(i-11)
    // (Layer 1 on the hydraulic group in droplet mode -- it reads layer 0's result only through the R1D ring -- evens the
    // instruction counts of the two groups, 520 / 310 per row instead of 206 / 599, and was measured 5 % SLOWER:
    // dispatch_particle 3.32 -> 3.51 ms at 8192^2.)
    // Smoothing reads only the G2 ring, so either group could run it.  The hydraulic group executes
    // 20 % fewer instructions per row than the thermal group, but moving G there was measured 6 %
    // slower (its instruction stream is the latency-heavy one: divisions, sqrt, the TMA wait).
    // Droplet mode: the hydraulic group only feeds rows to the thermal group, so it takes the smoothing stage
    // (and the momentum map) off the thermal group's hands.
    if (GROUP == HGF_ALL || GROUP == (DROPS ? HGF_HYDRO : SG)) {
        // own column of (rock1, dirt2): rows i-12 (y-1), i-11 (y), i-10 (y+1); stage F writes row i-9 meanwhile
        const int yg = i - HGF_LAG_G;
        if (FREE || (yg >= gy0 && yg < gy1)) {
            const HgF2 l = Q2(R::G2, 11, -1), r = Q2(R::G2, 11, 1);
            const HgF2 dn = Q2(R::G2, 12, 0), own = Q2(R::G2, 11, 0), up = Q2(R::G2, 10, 0);
            float rock = own.x, dirt = own.y;
            float sr_, sd_;
            hg_smooth_pair(P, sr_, sd_, own, l, r, up, dn);
            const bool border = (x == 0 || x == W - 1 || (!FREE && (yg == 0 || yg == H - 1)));
            if (owned) {
                const unsigned idx = off - (unsigned)HGF_LAG_G * pitch;
                if (DROPS) {      // smoothing.glsl:77-101 with the momentum map bound: momentum relaxation, display-water decay, H.a
                    float water = c.pf_w;
                    float mx = 0.0f, my = 0.0f, mz = 0.0f, mw = 0.0f;      // the border keeps its terrain; its momentum texel is defined as 0 (oracle: smooth_pass)
                    if (!border && P.particle_count != 0) {
                        mx = c.pf_m0; my = c.pf_m1; mz = c.pf_m2; mw = c.pf_m3;
                        hg_smooth_momentum(P, mx, my, mz, mw, water);
                    }
                    // H.a: the border texel keeps what thermal_transport.glsl:63 left, an interior one gets smoothing.glsl:101: both (r + g) + b
                    HgF4 h; h.x = border ? rock : sr_; h.y = border ? dirt : sd_; h.z = water; h.w = h.x + h.y + water;
                    HgF4 m; m.x = mx; m.y = my; m.z = mz; m.w = mw;
                    K.ha_dst[idx] = h;
                    K.ma_dst[idx] = m;
                } else {
                    K.dst[0][idx] = border ? rock : sr_;
                    K.dst[1][idx] = border ? dirt : sd_;
                    if (!FREE && K.push_mask) { HGF_PEER_STORE(K, 0, yg, idx, border ? rock : sr_) HGF_PEER_STORE(K, 1, yg, idx, border ? dirt : sd_) }
                }
            }
        }
        // The water and momentum texels are not staged through shared memory; fetching the next row's texels now keeps
        // the global-load latency off the row's critical path (the loads were 2 stall cycles per issued instruction).
        if (DROPS && owned) {
            const int yn = yg + 1;
            if ((FREE || yn >= gy0) && yn < gy1) {
                const unsigned idn = off - (unsigned)(HGF_LAG_G - 1) * pitch;
                c.pf_w = HGF_LDG(&K.ha_src[idn].z);
                const HgF4 m = HGF_LDG4(K.ma_src + idn);
                c.pf_m0 = m.x; c.pf_m1 = m.y; c.pf_m2 = m.z; c.pf_m3 = m.w;
            }
        }
    }
#undef A16
#undef A4
#undef B8
#undef Q4
#undef Q1
#undef Q2
#undef Q2X
}

// The rows of segment [gy0, gy1) for which FREE iterations are legal: every stage row is
// inside its active range and strictly inside the map.  Inclusive bounds on i.
HG_FN void hg_fused_free_range(int gy0, int gy1, int H, int* lo, int* hi) {
    int l = gy0 + HGF_LAG_G; if (l < 12) l = 12;
    int h = gy1 + 2; if (h > H - 2) h = H - 2;
    *lo = l; *hi = h;
}

// Iteration plan of one CTA: generic iterations [i_begin, free_lo), FREE iterations
// [free_lo, free_hi], then generic iterations up to i_end inclusive.
struct HgFusedPlan { int i_begin, i_end, free_lo, free_hi; };
HG_FN HgFusedPlan hg_fused_plan(int gy0, int gy1, int H) {
    HgFusedPlan p;
    p.i_begin = gy0 - HGF_HX;
    p.i_end = gy1 + HGF_LAG_G - 1;
    hg_fused_free_range(gy0, gy1, H, &p.free_lo, &p.free_hi);
    if (p.free_hi < p.free_lo) { p.free_lo = p.i_end + 1; p.free_hi = p.i_end; }
    return p;
}

// Connected slab: FREE iterations carry no peer stores, so they keep clear of the rows the neighbours hold as ghost rows
// (iteration i produces rows i-11 .. i-1).
HG_FN HgFusedPlan hg_fused_plan_slab(int gy0, int gy1, int H, const HgFusedK& K) {
    HgFusedPlan p = hg_fused_plan(gy0, gy1, H);
    if (K.push_mask) {
        if ((K.push_mask & 1) && p.free_lo < K.push_lo_end + HGF_LAG_G) p.free_lo = K.push_lo_end + HGF_LAG_G;
        if ((K.push_mask & 2) && p.free_hi > K.push_hi_begin) p.free_hi = K.push_hi_begin;
        if (p.free_hi < p.free_lo) { p.free_lo = p.i_end + 1; p.free_hi = p.i_end; }
    }
    return p;
}

// state before the first iteration: zeroed history
HG_FN void hg_fused_begin(HgCol& c) { hg_col_init(c); }
