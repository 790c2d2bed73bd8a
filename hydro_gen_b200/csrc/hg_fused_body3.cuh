// hg_fused_body3.cuh — the fused grid erosion step with the thermal OUTFLOW path taken off the row pipeline
// (k_fused_q).  Same arithmetic, same rings idea and the same one barrier per row as hg_fused_body.cuh; what changes
// is who evaluates thermal_erosion.glsl:59-115 for a cell that has a neighbour below the talus angle ("marked").
//
// Measured on k_fused_ws (profiles/r02h_*): a thermal warp enters that path in 49 % / 63 % of its rows (layer 0 / 1)
// with on average 4.2 of 32 lanes active, ~130 instructions each time: 35 % of the thermal group's instructions run at
// 13 % lane occupancy, and because the CTA meets at a barrier every row, the row period is set by the warp that has both
// layers marked (518 instructions against 256 for a warp with none and 402 for a hydraulic warp).
//
// Here a thermal thread only TESTS its cell (the d_h maxima against the two thresholds) and, when it is marked, pushes
// its column index on a CTA-wide queue in shared memory.  One iteration later a SERVICE warp pops the queue 32 cells at a
// time -- both layers mixed, lanes full whatever the terrain -- re-reads each cell's 3x3 window from the ring it came
// from, runs hg_thermal_outflow and stores the eight outflows and their negated sum into the outflow rings, which the
// thermal threads zero-filled when they tested.  The transport stages read them one iteration after that, so the lags
// between stages grow by one row per layer:
//
//   L(i)  A(i-1)  B(i-3)                       hydraulic group, unchanged
//   C(i-3)   test layer 0, zero-fill O0 row i-3, push            [service, iteration i+1: O0 row i-3]
//   D(i-6)   transport layer 0 (fresh outflow row i-5)  -> R1D row i-6
//   E(i-8)   test layer 1, zero-fill O1 row i-8, push            [service, iteration i+1: O1 row i-8]
//   F(i-11)  transport layer 1 (fresh outflow row i-10) -> G2 row i-11
//   G(i-13)  smoothing
//
// Rings (bytes per CTA at NT = 128: 66.7 KB with the raw-row staging, three CTAs per SM):
//   XQ  float4 [2]  XL float [2]  SS float2 [4]  G2 float2 [4]          as before
//   RD  float2 [8]  (rockE, dirtE): written by A, read by C (rows i-4..i-2), by the service (i-5..i-3), by D (own, i-6)
//   R1D float2 [8]  (rock1, dirtE): written by D, read by E (i-9..i-7), by the service (i-10..i-8), by F (own, i-11)
//   OxR float4 [3]  (R, RT, RB, T)   OxL float4 [3]  (L, LT, LB, B)   SOx float [3]  negated outflow sum; per layer
//   QUEUE uint16 [2][2*NT] (column | layer << 15), QCNT uint32 [4]
// The three outflow-ring slots rotate: at iteration i slot(i-3) is zero-filled (layer 1: slot(i-8) = slot(i-5)),
// slot(i-4) (layer 1: slot(i-9) = slot(i-3)) is being served, slot(i-5) (layer 1: slot(i-10) = slot(i-4)) is read.
//
// Bit-exactness: the test a thermal thread makes and the one hg_thermal_outflow repeats on the service warp see the
// same d_h (same ring values, same operations), so a cell is served iff it is marked; an unmarked cell's outflows
// are the zeros of the fill.  Order inside the queue is free: every item writes only its own ring elements.
//
// Plain C++ apart from the HGF_* / HGQ_* macros: tests/host_emul runs this body thread by thread on the CPU.
#pragma once
#include "hg_fused_body.cuh"

constexpr int HGQ_LAG_G = 13;      // rows between L and G
constexpr int HGQ_SERVICE = 5;     // warp-group id of the service warps (HGF_HYDRO / HGF_THERMAL for the others)

template <int NT> struct HgRingsQ {
    static constexpr int E = NT + 2;
    static constexpr int XQ = 0;                          // float4 [2][E]
    static constexpr int O0R = XQ + 2 * E * 16;           // float4 [3][E]
    static constexpr int O0L = O0R + 3 * E * 16;
    static constexpr int O1R = O0L + 3 * E * 16;
    static constexpr int O1L = O1R + 3 * E * 16;
    static constexpr int SS = O1L + 3 * E * 16;           // float2 [4][E]
    static constexpr int G2 = SS + 4 * E * 8;             // float2 [4][E]
    static constexpr int RD = G2 + 4 * E * 8;             // float2 [8][E]
    static constexpr int R1D = RD + 8 * E * 8;            // float2 [8][E]
    static constexpr int XL = R1D + 8 * E * 8;            // float [2][E]
    static constexpr int SO0 = XL + 2 * E * 4;            // float [3][E]
    static constexpr int SO1 = SO0 + 3 * E * 4;
    static constexpr int QUEUE = (SO1 + 3 * E * 4 + 15) / 16 * 16;   // uint16 [2][2*NT]
    static constexpr int QCNT = QUEUE + 2 * 2 * NT * 2;   // uint32 [4]
    static constexpr int TOTAL_BYTES = QCNT + 16;
    static constexpr int TOTAL = TOTAL_BYTES / 4;
};

// per-thread rolling state (one column); a group only touches its own part
struct HgColQ {
    // hydraulic
    float rk0, rk1, rk2, dt0, dt1, dt2;
    float at0, at1, at2;
    float w1, w2;
    float f1L, f1R, f1T, f1B, f2L, f2R, f2T, f2B, f0T;
    float s1r, s1d, s2r, s2d;
    float u_d1, v_d1, u_d2, v_d2;
    // thermal: outflow values of earlier rows, per layer (fresh row = the one served last iteration)
    float so0_d1, nR0_d1, nL0_d1, T0_d1, T0_d2, nRT0_d1, nRT0_d2, nLT0_d1, nLT0_d2;
    float so1_d1, nR1_d1, nL1_d1, T1_d1, T1_d2, nRT1_d1, nRT1_d2, nLT1_d1, nLT1_d2;
    float pf_w, pf_m0, pf_m1, pf_m2, pf_m3;      // droplet mode, as in HgCol
};
HG_FN void hg_colq_init(HgColQ& c) {
    float* f = reinterpret_cast<float*>(&c);
    for (int k = 0; k < (int)(sizeof(HgColQ) / sizeof(float)); k++) f[k] = 0.0f;
    c.at0 = c.at1 = c.at2 = HG_OOB_HEIGHT;
}

// queue push.  Device: one shared-memory atomic per warp and layer (the lanes of a warp take consecutive slots);
// host emulation: sequential.
#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define HGQ_PUSH(marked, cntp, qp, item)                                                        \
    {                                                                                           \
        const unsigned m_ = __ballot_sync(0xffffffffu, (marked));                               \
        if (m_) {                                                                               \
            const unsigned lane_ = threadIdx.x & 31u;                                           \
            unsigned base_ = 0;                                                                 \
            if (lane_ == 0) base_ = atomicAdd((cntp), (unsigned)__popc(m_));                    \
            base_ = __shfl_sync(0xffffffffu, base_, 0);                                         \
            if (marked) (qp)[base_ + __popc(m_ & ((1u << lane_) - 1u))] = (unsigned short)(item); \
        }                                                                                       \
    }
#else
#define HGQ_PUSH(marked, cntp, qp, item) { if (marked) { (qp)[*(cntp)] = (unsigned short)(item); *(cntp) += 1u; } }
#endif

// The marking test of hg_thermal_outflow alone (the same expression on the same values).
HG_FN bool hg_thermal_marked(const HgStepParams& P, int layer, const float d_h[8], bool live) {
    const float thc = P.th_mark[layer][0], thd = P.th_mark[layer][1];
    const float mc = fmaxf(fmaxf(d_h[0], d_h[1]), fmaxf(d_h[2], d_h[3]));
    const float md = fmaxf(fmaxf(d_h[4], d_h[5]), fmaxf(d_h[6], d_h[7]));
    return live && (mc >= thc || md >= thd);
}

// One iteration of a hydraulic (GROUP = HGF_HYDRO) or thermal (GROUP = HGF_THERMAL) thread.  m3 = (i - 3) mod 3.
template <int NT, bool FREE, int GROUP, bool DROPS = false>
HG_FN void hg_fusedq_iter(HgColQ& c, float* sm, const float* raw, const HgFusedK& K, const int tid, const int x, const bool xin, const bool owned,
                          const int gy0, const int gy1, const int i, const int m3, const unsigned off) {
    typedef HgRingsQ<NT> R;
    const HgStepParams& P = K.P;
    const int W = K.W, H = K.H;
    const unsigned pitch = (unsigned)K.pitch;
    const int e = tid + 1;
    char* const smc = reinterpret_cast<char*>(sm);
    // element offset (slot * E) of the ring row that holds absolute row i - k
#define S2(k) (((i - (k)) & 1) * R::E)
#define S4(k) (((i - (k)) & 3) * R::E)
#define S8(k) (((i - (k)) & 7) * R::E)
#define Q4(ring, k, d) (*reinterpret_cast<HgF4*>(smc + (ring) + (S2(k) + e + (d)) * 16))
#define Q1(ring, k, d) (*reinterpret_cast<float*>(smc + (ring) + (S2(k) + e + (d)) * 4))
#define Q2(ring, k, d) (*reinterpret_cast<HgF2*>(smc + (ring) + (S4(k) + e + (d)) * 8))
#define Q8(ring, k, d) (*reinterpret_cast<HgF2*>(smc + (ring) + (S8(k) + e + (d)) * 8))
#define Q8X(ring, k, d) (reinterpret_cast<const float*>(smc + (ring) + (S8(k) + e + (d)) * 8)[0])
    // outflow-ring slots: zero-filled / read, per layer (header)
    const int o_a = m3 * R::E, o_b = (m3 == 0 ? 2 : m3 - 1) * R::E, o_c = (m3 == 2 ? 0 : m3 + 1) * R::E;   // slot(i-3), slot(i-4), slot(i-5)
#define O4(ring, o, d) (*reinterpret_cast<HgF4*>(smc + (ring) + ((o) + e + (d)) * 16))
#define O4W(ring, o) (reinterpret_cast<const float*>(smc + (ring) + ((o) + e) * 16)[3])
#define O1(ring, o) (*reinterpret_cast<float*>(smc + (ring) + ((o) + e) * 4))
    unsigned* const qcnt = reinterpret_cast<unsigned*>(smc + R::QCNT) + (i & 3);
    unsigned short* const qbuf = reinterpret_cast<unsigned short*>(smc + R::QUEUE) + (i & 1) * (2 * NT);

    if (GROUP == HGF_HYDRO && DROPS) {
    // ------------------------------------------------------------ droplet mode: no hydraulics (hg_fused_body.cuh)
    c.rk1 = c.rk2; c.dt1 = c.dt2;
    {
        const HgF4 t = reinterpret_cast<const HgF4*>(raw)[tid + 2];
        c.rk2 = t.x; c.dt2 = t.y;
    }
    {
        const int ya = i - 1;
        if (FREE || (ya >= gy0 - 5 && ya < gy1 + 5)) {
            const bool in = xin && (FREE || (ya >= 0 && ya < H));
            HgF2 rd; rd.x = in ? c.rk1 : HG_OOB_HEIGHT; rd.y = in ? c.dt1 : HG_OOB_HEIGHT;
            Q8(R::RD, 1, 0) = rd;
        }
    }
    }
    if (GROUP == HGF_HYDRO && !DROPS) {
    // ------------------------------------------------------------ L(i)
    c.rk0 = c.rk1; c.rk1 = c.rk2; c.dt0 = c.dt1; c.dt1 = c.dt2; c.at0 = c.at1; c.at1 = c.at2; c.w1 = c.w2;
    c.f0T = c.f1T; c.f1L = c.f2L; c.f1R = c.f2R; c.f1T = c.f2T; c.f1B = c.f2B; c.s1r = c.s2r; c.s1d = c.s2d;
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
            const HgF4 ql = Q4(R::XQ, 1, -1);
            const HgF4 qr = Q4(R::XQ, 1, 1);
            const float inR = Q1(R::XL, 1, 1);
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
                K.dst[2][idx] = o.water * P.evap;
            }
            HgF2 rd; rd.x = in ? er.rock : HG_OOB_HEIGHT; rd.y = in ? er.dirt : HG_OOB_HEIGHT;
            Q8(R::RD, 1, 0) = rd;
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
            const bool up = fast && dy == -1;
            const char* const r0 = smc + R::SS + ((up ? S4(4) : S4(3)) + e + cdx) * 8;
            const char* const r1 = smc + R::SS + ((up ? S4(3) : S4(2)) + e + cdx) * 8;
            const HgF2 t00 = reinterpret_cast<const HgF2*>(r0)[0], t10 = reinterpret_cast<const HgF2*>(r0)[1];
            const HgF2 t01 = reinterpret_cast<const HgF2*>(r1)[0], t11 = reinterpret_cast<const HgF2*>(r1)[1];
            float sr = hg_bilerp(t00.x, t10.x, t01.x, t11.x, b.sx, b.sy);
            float sd = hg_bilerp(t00.y, t10.y, t01.y, t11.y, b.sx, b.sy);
            if (owned) {
                const unsigned idx = off - 3u * pitch;
                if (fast) {
                    K.dst[7][idx] = sr;
                    K.dst[8][idx] = sd;
                } else {
                    unsigned long long slot = HGF_ATOMIC_INC64(K.far_count);
                    K.far_list[slot] = (unsigned)(yb - K.row0) * (unsigned)W + (unsigned)x;
                }
            }
        }
#endif
    }
    c.u_d2 = c.u_d1; c.v_d2 = c.v_d1; c.u_d1 = u_new; c.v_d1 = v_new;
    }   // hydraulic

    if (GROUP == HGF_THERMAL) {
    // ------------------------------------------------------------ C(i-3): test layer 0
    {
        const int yc = i - 3;
        if (FREE || (yc >= gy0 - 4 && yc < gy1 + 4)) {
            const bool in = xin && (FREE || (yc >= 0 && yc < H));
            const float e00 = Q8X(R::RD, 4, -1), e01 = Q8X(R::RD, 4, 0), e02 = Q8X(R::RD, 4, 1);
            const float e10 = Q8X(R::RD, 3, -1), e11 = Q8X(R::RD, 3, 0), e12 = Q8X(R::RD, 3, 1);
            const float e20 = Q8X(R::RD, 2, -1), e21 = Q8X(R::RD, 2, 0), e22 = Q8X(R::RD, 2, 1);
            float d_h[8];
            d_h[0] = e11 - e10; d_h[1] = e11 - e12; d_h[2] = e11 - e21; d_h[3] = e11 - e01;
            d_h[4] = e11 - e20; d_h[5] = e11 - e22; d_h[6] = e11 - e00; d_h[7] = e11 - e02;
            const bool marked = hg_thermal_marked(P, 0, d_h, in);
            HgF4 z; z.x = 0.0f; z.y = 0.0f; z.z = 0.0f; z.w = 0.0f;
            O4(R::O0R, o_a, 0) = z;
            O4(R::O0L, o_a, 0) = z;
            O1(R::SO0, o_a) = 0.0f;
            HGQ_PUSH(marked, qcnt, qbuf, tid)
        }
    }
    // ------------------------------------------------------------ D(i-6): fresh outflow row i-5
    {
        const int yd = i - 6;
        const HgF4 nl = O4(R::O0R, o_c, -1);      // left neighbour's R, RT, RB
        const HgF4 nr = O4(R::O0L, o_c, 1);       // right neighbour's L, LT, LB
        const float Tf = O4W(R::O0R, o_c), Bf = O4W(R::O0L, o_c), sof = O1(R::SO0, o_c);
        if (FREE || (yd >= gy0 - 3 && yd < gy1 + 3)) {
            const bool in = xin && (FREE || (yd >= 0 && yd < H));
            const HgF2 own = Q8(R::RD, 6, 0);     // (rockE, dirtE) of row i-6
            float delta = hg_thermal_delta(c.so0_d1, c.nR0_d1, c.nL0_d1, Bf, c.T0_d2, nl.z, nr.z, c.nRT0_d2, c.nLT0_d2);
            HgF2 w; w.x = in ? own.x + delta : HG_OOB_HEIGHT; w.y = own.y;
            Q8(R::R1D, 6, 0) = w;
        }
        c.so0_d1 = sof;
        c.T0_d2 = c.T0_d1; c.T0_d1 = Tf;
        c.nR0_d1 = nl.x; c.nL0_d1 = nr.x;
        c.nRT0_d2 = c.nRT0_d1; c.nRT0_d1 = nl.y; c.nLT0_d2 = c.nLT0_d1; c.nLT0_d1 = nr.y;
    }
    // ------------------------------------------------------------ E(i-8): test layer 1
    {
        const int ye = i - 8;
        if (FREE || (ye >= gy0 - 2 && ye < gy1 + 2)) {
            const bool in = xin && (FREE || (ye >= 0 && ye < H));
            const HgF2 w00 = Q8(R::R1D, 9, -1), w01 = Q8(R::R1D, 9, 0), w02 = Q8(R::R1D, 9, 1);
            const HgF2 w10 = Q8(R::R1D, 8, -1), w11 = Q8(R::R1D, 8, 0), w12 = Q8(R::R1D, 8, 1);
            const HgF2 w20 = Q8(R::R1D, 7, -1), w21 = Q8(R::R1D, 7, 0), w22 = Q8(R::R1D, 7, 1);
            float d_h[8];
            d_h[0] = (w11.x - w10.x) + (w11.y - w10.y); d_h[1] = (w11.x - w12.x) + (w11.y - w12.y);
            d_h[2] = (w11.x - w21.x) + (w11.y - w21.y); d_h[3] = (w11.x - w01.x) + (w11.y - w01.y);
            d_h[4] = (w11.x - w20.x) + (w11.y - w20.y); d_h[5] = (w11.x - w22.x) + (w11.y - w22.y);
            d_h[6] = (w11.x - w00.x) + (w11.y - w00.y); d_h[7] = (w11.x - w02.x) + (w11.y - w02.y);
            const bool marked = hg_thermal_marked(P, 1, d_h, in);
            HgF4 z; z.x = 0.0f; z.y = 0.0f; z.z = 0.0f; z.w = 0.0f;
            O4(R::O1R, o_c, 0) = z;               // slot(i-8) = slot(i-5)
            O4(R::O1L, o_c, 0) = z;
            O1(R::SO1, o_c) = 0.0f;
            HGQ_PUSH(marked, qcnt, qbuf, tid | 0x8000)
        }
    }
    // ------------------------------------------------------------ F(i-11): fresh outflow row i-10
    {
        const int yf = i - 11;
        const HgF4 nl = O4(R::O1R, o_b, -1);      // slot(i-10) = slot(i-4)
        const HgF4 nr = O4(R::O1L, o_b, 1);
        const float Tf = O4W(R::O1R, o_b), Bf = O4W(R::O1L, o_b), sof = O1(R::SO1, o_b);
        if (FREE || (yf >= gy0 - 1 && yf < gy1 + 1)) {
            const bool in = xin && (FREE || (yf >= 0 && yf < H));
            const HgF2 own = Q8(R::R1D, 11, 0);   // (rock1, dirtE) of row i-11
            float delta = hg_thermal_delta(c.so1_d1, c.nR1_d1, c.nL1_d1, Bf, c.T1_d2, nl.z, nr.z, c.nRT1_d2, c.nLT1_d2);
            HgF2 w; w.x = own.x; w.y = in ? own.y + delta : HG_OOB_HEIGHT;
            Q2(R::G2, 11, 0) = w;
        }
        c.so1_d1 = sof;
        c.T1_d2 = c.T1_d1; c.T1_d1 = Tf;
        c.nR1_d1 = nl.x; c.nL1_d1 = nr.x;
        c.nRT1_d2 = c.nRT1_d1; c.nRT1_d1 = nl.y; c.nLT1_d2 = c.nLT1_d1; c.nLT1_d1 = nr.y;
    }
    }   // thermal

    // ------------------------------------------------------------ G(i-13): smoothing (thermal group; droplet mode: hydraulic group)
    if (GROUP == (DROPS ? HGF_HYDRO : HGF_THERMAL)) {
        const int yg = i - HGQ_LAG_G;
        if (FREE || (yg >= gy0 && yg < gy1)) {
            const HgF2 l = Q2(R::G2, 13, -1), r = Q2(R::G2, 13, 1);
            const HgF2 dn = Q2(R::G2, 14, 0), own = Q2(R::G2, 13, 0), up = Q2(R::G2, 12, 0);
            float rock = own.x, dirt = own.y;
            float sr_ = rock, sd_ = dirt;
            hg_smooth_cell(P, sr_, sd_, l.x, l.y, r.x, r.y, up.x, up.y, dn.x, dn.y);
            const bool border = (x == 0 || x == W - 1 || (!FREE && (yg == 0 || yg == H - 1)));
            if (owned) {
                const unsigned idx = off - (unsigned)HGQ_LAG_G * pitch;
                if (DROPS) {
                    float water = c.pf_w;
                    float mx = 0.0f, my = 0.0f, mz = 0.0f, mw = 0.0f;
                    if (!border && P.particle_count != 0) {
                        mx = c.pf_m0; my = c.pf_m1; mz = c.pf_m2; mw = c.pf_m3;
                        hg_smooth_momentum(P, mx, my, mz, mw, water);
                    }
                    HgF4 h; h.x = border ? rock : sr_; h.y = border ? dirt : sd_; h.z = water; h.w = h.x + h.y + water;
                    HgF4 m; m.x = mx; m.y = my; m.z = mz; m.w = mw;
                    K.ha_dst[idx] = h;
                    K.ma_dst[idx] = m;
                } else {
                    K.dst[0][idx] = border ? rock : sr_;
                    K.dst[1][idx] = border ? dirt : sd_;
                }
            }
        }
        if (DROPS && owned) {
            const int yn = yg + 1;
            if ((FREE || yn >= gy0) && yn < gy1) {
                const unsigned idn = off - (unsigned)(HGQ_LAG_G - 1) * pitch;
                c.pf_w = HGF_LDG(&K.ha_src[idn].z);
                const HgF4 m = HGF_LDG4(K.ma_src + idn);
                c.pf_m0 = m.x; c.pf_m1 = m.y; c.pf_m2 = m.z; c.pf_m3 = m.w;
            }
        }
    }
#undef Q4
#undef Q1
#undef Q2
#undef Q8
#undef Q8X
#undef O4
#undef O4W
#undef O1
#undef S2
#undef S4
#undef S8
}

// Service: one queued cell.  `i` is the CURRENT iteration: the item was pushed at iteration i-1 by C (layer 0, row i-4)
// or E (layer 1, row i-9).  Re-reads the 3x3 window from RD / R1D, thermal_erosion.glsl:59-115 through the shared
// hg_thermal_outflow, results into the cell's own elements of the outflow rings.
template <int NT>
HG_FN void hg_fusedq_serve(float* sm, const HgFusedK& K, const int i, const int m3, const unsigned item) {
    typedef HgRingsQ<NT> R;
    char* const smc = reinterpret_cast<char*>(sm);
    const int layer = (int)(item >> 15);
    const int e = (int)(item & 0x7fffu) + 1;
    // window rows y-1, y, y+1: layer 0 rows i-5, i-4, i-3 of RD; layer 1 rows i-10, i-9, i-8 of R1D
    const int ring = layer ? R::R1D : R::RD;
    const int k1 = layer ? 9 : 4;
    const char* const r0 = smc + ring + ((((i - k1 - 1) & 7) * R::E) + e) * 8;
    const char* const r1 = smc + ring + ((((i - k1) & 7) * R::E) + e) * 8;
    const char* const r2 = smc + ring + ((((i - k1 + 1) & 7) * R::E) + e) * 8;
    const HgF2 w00 = reinterpret_cast<const HgF2*>(r0)[-1], w01 = reinterpret_cast<const HgF2*>(r0)[0], w02 = reinterpret_cast<const HgF2*>(r0)[1];
    const HgF2 w10 = reinterpret_cast<const HgF2*>(r1)[-1], w11 = reinterpret_cast<const HgF2*>(r1)[0], w12 = reinterpret_cast<const HgF2*>(r1)[1];
    const HgF2 w20 = reinterpret_cast<const HgF2*>(r2)[-1], w21 = reinterpret_cast<const HgF2*>(r2)[0], w22 = reinterpret_cast<const HgF2*>(r2)[1];
    float d_h[8], out[8];
#define HGQ_DH(n) (layer ? (w11.x - (n).x) + (w11.y - (n).y) : w11.x - (n).x)
    d_h[0] = HGQ_DH(w10); d_h[1] = HGQ_DH(w12); d_h[2] = HGQ_DH(w21); d_h[3] = HGQ_DH(w01);
    d_h[4] = HGQ_DH(w20); d_h[5] = HGQ_DH(w22); d_h[6] = HGQ_DH(w00); d_h[7] = HGQ_DH(w02);
#undef HGQ_DH
    const float neg = hg_thermal_outflow(K.P, layer, layer ? w11.y : w11.x, d_h, out, true);
    // slot of the served row: layer 0 slot(i-4); layer 1 slot(i-9) = slot(i-3)
    const int o_a = m3 * R::E, o_b = (m3 == 0 ? 2 : m3 - 1) * R::E;
    const int o = layer ? o_a : o_b;
    HgF4 tr; tr.x = out[1]; tr.y = out[5]; tr.z = out[7]; tr.w = out[2];    // R, RT, RB, T
    HgF4 tl; tl.x = out[0]; tl.y = out[4]; tl.z = out[6]; tl.w = out[3];    // L, LT, LB, B
    *reinterpret_cast<HgF4*>(smc + (layer ? R::O1R : R::O0R) + (o + e) * 16) = tr;
    *reinterpret_cast<HgF4*>(smc + (layer ? R::O1L : R::O0L) + (o + e) * 16) = tl;
    *reinterpret_cast<float*>(smc + (layer ? R::SO1 : R::SO0) + (o + e) * 4) = neg;
}

// Iteration plan with this body's lags (see hg_fused_plan)
HG_FN HgFusedPlan hg_fusedq_plan(int gy0, int gy1, int H) {
    HgFusedPlan p;
    p.i_begin = gy0 - HGF_HX;
    p.i_end = gy1 + HGQ_LAG_G - 1;
    int l = gy0 + HGQ_LAG_G; if (l < HGQ_LAG_G + 1) l = HGQ_LAG_G + 1;
    int h = gy1 + 2; if (h > H - 2) h = H - 2;
    p.free_lo = l; p.free_hi = h;
    if (p.free_hi < p.free_lo) { p.free_lo = p.i_end + 1; p.free_hi = p.i_end; }
    return p;
}
