// hg_fused_body2.cuh — one row iteration of the fused grid erosion step for one thread that advances TWO columns.
//
// Same row-marching software pipeline as hg_fused_body.cuh (stages L, A, B on the hydraulic warp group, C..G on the
// thermal one; lags 0, 1, 3, 3, 5, 7, 9, 11; one CTA barrier per row; every hand-off between threads through a
// shared-memory row ring written at least one iteration earlier).  The difference is the unit of work of a thread:
// a CTA owns a PAIR of adjacent strips of NT columns each (NT-12 owned + 6 recomputed halo columns per side; the
// second strip starts NT-12 columns after the first, so their halos overlap each other's owned columns) and thread t
// holds column t of BOTH strips in the two lanes of a V2 (hg_v2.cuh).  All column arithmetic is then one packed
// FADD2/FMUL2/FFMA2 per two cells, and the ring addressing, the loop control, the barrier and the history moves are
// shared by two cells: about half the issue slots per cell of the one-column body, with bit-identical results (every
// packed lane rounds like the scalar instruction; tests/host_emul runs this body against the oracle on the CPU).
//
// Ring elements hold both lanes of a thread, lane-interleaved so that what a neighbour loads is already a register
// pair: XQ {at, rock, dirt, fR} x V2 (32 B), XL fL (8 B), RD {rockE, dirtE} (16 B), R1D {rock1, dirtE}, G2 {rock1,
// dirt2}, OxR {R, RT, RB, -} / OxL {L, LT, LB, -} (32 B).  Only the sediment ring SS is lane-major, {sr0, sd0, sr1,
// sd1}, because the bilinear gather of a lane fetches (S'.rock, S'.dirt) of one texel from a data-dependent element.
#pragma once
#include "hg_cell2.cuh"
#include "hg_fused_body.cuh"

struct HgV2x4 { V2 a, b, c, d; };    // 32 bytes
struct HgV2x2 { V2 a, b; };          // 16 bytes
#if defined(__CUDACC__)
static_assert(sizeof(HgV2x4) == 32 && sizeof(HgV2x2) == 16 && sizeof(V2) == 8, "packed ring elements");
#endif

// Byte offsets of the rings; every ring row has NT+2 elements (one pad element each side).
template <int NT> struct HgRings2 {
    static constexpr int E = NT + 2;
    static constexpr int XQ = 0;                     // HgV2x4 [2][E]
    static constexpr int O0R = XQ + 2 * E * 32;      // HgV2x4 [2][E]
    static constexpr int O0L = O0R + 2 * E * 32;
    static constexpr int O1R = O0L + 2 * E * 32;
    static constexpr int O1L = O1R + 2 * E * 32;
    static constexpr int SS = O1L + 2 * E * 32;      // HgV2x2 [4][E]  (lane-major)
    static constexpr int RD = SS + 4 * E * 16;       // HgV2x2 [4][E]
    static constexpr int R1D = RD + 4 * E * 16;      // HgV2x2 [4][E]
    static constexpr int G2 = R1D + 4 * E * 16;      // HgV2x2 [4][E]
    static constexpr int XL = G2 + 4 * E * 16;       // V2     [2][E]
    static constexpr int TOTAL_BYTES = XL + 2 * E * 8;
    static constexpr int TOTAL = TOTAL_BYTES / 4;
};
// the raw block of one row: nine planes x (2*NT - 8) columns (both strips of the pair and 2 + 2 alignment columns)
#define HGF2_RAW_LD(NT) (2 * (NT) - 8)
#define HGF2_HALF(NT) ((NT) - 2 * HGF_HX)     // column distance between the two lanes of a thread

// per-thread rolling state (two columns)
struct HgCol2 {
    V2 rk0, rk1, rk2, dt0, dt1, dt2;
    V2 at0, at1, at2;
    V2 w1, w2;
    V2 f1L, f1R, f1T, f1B, f2L, f2R, f2T, f2B, f0T;
    V2 s1r, s1d, s2r, s2d;
    V2 u_d1, v_d1, u_d2, v_d2;
    V2 e_old, de_old;
    V2 so0_d1, so0_d2, T0_d1, T0_d2, T0_d3, B0_d1;
    V2 nR0_d1, nL0_d1, nRT0_d1, nRT0_d2, nLT0_d1, nLT0_d2;
    V2 p_old, q_old;
    V2 so1_d1, so1_d2, T1_d1, T1_d2, T1_d3, B1_d1;
    V2 nR1_d1, nL1_d1, nRT1_d1, nRT1_d2, nLT1_d1, nLT1_d2;
};

HG_FN void hg_col2_init(HgCol2& c) {
    float* f = reinterpret_cast<float*>(&c);
    for (int k = 0; k < (int)(sizeof(HgCol2) / sizeof(float)); k++) f[k] = 0.0f;
    c.at0 = c.at1 = c.at2 = v2s(HG_OOB_HEIGHT);
}

// per-thread constants of the two lanes
struct HgLanes {
    int x0, x1;              // global columns
    B2 xin, owned;           // column inside the map / inside its strip proper
    B2 x_left, x_right;      // hydro_flux.glsl:110-113: x <= 0; x >= W-1 and not x <= 0
    B2 border;               // smoothing.glsl:27: x == 0 or x == W-1
};
HG_FN HgLanes hg_lanes(int x0, int half, int tid, int NT, int W) {
    HgLanes L;
    L.x0 = x0; L.x1 = x0 + half;
    L.xin = b2(L.x0 >= 0 && L.x0 < W, L.x1 >= 0 && L.x1 < W);
    const bool own_t = tid >= HGF_HX && tid < NT - HGF_HX;
    L.owned = b2(own_t && L.x0 < W, own_t && L.x1 < W);
    L.x_left = b2(L.x0 <= 0, L.x1 <= 0);
    L.x_right = b2(!(L.x0 <= 0) && L.x0 >= W - 1, !(L.x1 <= 0) && L.x1 >= W - 1);
    L.border = b2(L.x0 == 0 || L.x0 == W - 1, L.x1 == 0 || L.x1 == W - 1);
    return L;
}

// One lane of stage B: back-trace, fast-path test and the bilinear gather from the sediment ring.  ss_rows: byte
// pointers of the SS ring rows yb-1, yb, yb+1 at this thread's element; lane: 0 / 1 (selects the half of an element).
struct HgGather { float sr, sd; bool fast; };
HG_FN HgGather hg_gather_lane(const HgStepParams& P, int x, int yb, int W, int H, float u, float v,
                              const char* row_m1, const char* row_0, const char* row_p1, int lane) {
    HgBack b = hg_backtrace(P, x, yb, W, H, u, v);
    const int dx = b.px - x, dy = b.py - yb;
    HgGather g;
    g.fast = dx >= -1 && dx <= 0 && dy >= -1 && dy <= 0;
    const int cdx = g.fast ? dx : 0;
    const bool up = g.fast && dy == -1;          // footprint rows (yb-1, yb) instead of (yb, yb+1)
    const char* const r0 = (up ? row_m1 : row_0) + cdx * 16 + lane * 8;
    const char* const r1 = (up ? row_0 : row_p1) + cdx * 16 + lane * 8;
    const V2 t00 = *reinterpret_cast<const V2*>(r0), t10 = *reinterpret_cast<const V2*>(r0 + 16);
    const V2 t01 = *reinterpret_cast<const V2*>(r1), t11 = *reinterpret_cast<const V2*>(r1 + 16);
    // img_bilinear on the (rock, dirt) channel pair at once: mix(a, b, t) = a * (1 - t) + b * t per channel
    const float ix = 1.0f - b.sx, iy = 1.0f - b.sy;
    const V2 v1 = t00 * ix + t10 * b.sx;
    const V2 v2_ = t01 * ix + t11 * b.sx;
    const V2 r = v1 * iy + v2_ * b.sy;
    g.sr = r.x; g.sd = r.y;
    return g;
}

template <int NT, bool FREE, int GROUP>
HG_FN void hg_fused_iter2(HgCol2& c, float* sm, const float* raw, const HgFusedK& K, const int tid, const HgLanes& L,
                          const int gy0, const int gy1, const int i, const unsigned off) {
    typedef HgRings2<NT> R;
    const HgStepParams& P = K.P;
    const int W = K.W, H = K.H;
    const unsigned pitch = (unsigned)K.pitch;
    constexpr int HALF = HGF2_HALF(NT);
    const int e = tid + 1;
    char* const smc = reinterpret_cast<char*>(sm);
    const int u0 = (i & 1) * R::E + e, u1 = (R::E + 2 * e) - u0;
    const int v0 = (i & 3) * R::E + e, v1 = ((i - 1) & 3) * R::E + e, v2i = ((i - 2) & 3) * R::E + e, v3 = ((i - 3) & 3) * R::E + e;
    char* const a32_0 = smc + u0 * 32; char* const a32_1 = smc + u1 * 32;     // 32-byte rings
    char* const a8_0 = smc + u0 * 8; char* const a8_1 = smc + u1 * 8;         // XL
    char* const b16_0 = smc + v0 * 16; char* const b16_1 = smc + v1 * 16; char* const b16_2 = smc + v2i * 16; char* const b16_3 = smc + v3 * 16;
#define A32(k) (((k) & 1) ? a32_1 : a32_0)
#define A8(k) (((k) & 1) ? a8_1 : a8_0)
#define B16(k) (((k) & 3) == 0 ? b16_0 : ((k) & 3) == 1 ? b16_1 : ((k) & 3) == 2 ? b16_2 : b16_3)
#define Q4(ring, k, d) (*reinterpret_cast<HgV2x4*>(A32(k) + (ring) + (d) * 32))
#define Q4AB(ring, k, d) (*reinterpret_cast<HgV2x2*>(A32(k) + (ring) + (d) * 32))          /* first two fields: one 16-byte load */
#define Q4C(ring, k, d) (*reinterpret_cast<V2*>(A32(k) + (ring) + (d) * 32 + 16))          /* third field */
#define Q1(ring, k, d) (*reinterpret_cast<V2*>(A8(k) + (ring) + (d) * 8))
#define Q2(ring, k, d) (*reinterpret_cast<HgV2x2*>(B16(k) + (ring) + (d) * 16))
#define Q2X(ring, k, d) (*reinterpret_cast<V2*>(B16(k) + (ring) + (d) * 16))               /* first field only */

    if (GROUP != HGF_THERMAL) {
    // ------------------------------------------------------------ L(i)
    c.rk0 = c.rk1; c.rk1 = c.rk2; c.dt0 = c.dt1; c.dt1 = c.dt2; c.at0 = c.at1; c.at1 = c.at2; c.w1 = c.w2;
    c.f0T = c.f1T; c.f1L = c.f2L; c.f1R = c.f2R; c.f1T = c.f2T; c.f1B = c.f2B; c.s1r = c.s2r; c.s1d = c.s2d;
    {
        const float* rw = raw + tid + 2;
        constexpr int LD = HGF2_RAW_LD(NT);
        c.rk2 = v2(rw[0 * LD], rw[0 * LD + HALF]); c.dt2 = v2(rw[1 * LD], rw[1 * LD + HALF]); c.w2 = v2(rw[2 * LD], rw[2 * LD + HALF]);
        c.f2L = v2(rw[3 * LD], rw[3 * LD + HALF]); c.f2R = v2(rw[4 * LD], rw[4 * LD + HALF]);
        c.f2T = v2(rw[5 * LD], rw[5 * LD + HALF]); c.f2B = v2(rw[6 * LD], rw[6 * LD + HALF]);
        c.s2r = v2(rw[7 * LD], rw[7 * LD + HALF]); c.s2d = v2(rw[8 * LD], rw[8 * LD + HALF]);
    }
    {
        const bool yin = FREE || (i >= 0 && i < H);
        c.at2 = v2_sel(L.xin && yin, c.rk2 + c.dt2 + c.w2, HG_OOB_HEIGHT);
        HgV2x4 q; q.a = c.at2; q.b = c.rk2; q.c = c.dt2; q.d = c.f2R;
        Q4(R::XQ, 0, 0) = q;
        Q1(R::XL, 0, 0) = c.f2L;
    }

    // ------------------------------------------------------------ A(i-1)
    V2 u_new = v2s(0.0f), v_new = v2s(0.0f);
    {
        const int ya = i - 1;
        if (FREE || (ya >= gy0 - 5 && ya < gy1 + 5)) {
            const bool yin = FREE || (ya >= 0 && ya < H);
            const B2 in = L.xin && yin;
            const HgV2x4 ql = Q4(R::XQ, 1, -1);         // left neighbour: H.a, rock, dirt, fR
            const HgV2x2 qr = Q4AB(R::XQ, 1, 1);        // right neighbour: H.a, rock
            const V2 qr_dirt = Q4C(R::XQ, 1, 1);        //                  dirt
            const V2 inR = Q1(R::XL, 1, 1);             // right neighbour's fL
            HgFluxOut2 o = hg_flux_cell2(P, L.x_left, L.x_right, FREE ? 1 : ya, FREE ? 4 : H, c.at1, ql.a, qr.a, c.at2, c.at0,
                                         c.f1L, c.f1R, c.f1T, c.f1B, ql.d, inR, c.f2B, c.f0T, c.w1);
            HgEroOut e0, e1;
            hg_erosion_cell2(P, c.rk1, c.dt1, c.s1r, c.s1d, o.u, o.v, o.vz, qr.b, qr_dirt, ql.b, ql.c, c.rk0, c.dt0, c.rk2, c.dt2, e0, e1);
            u_new = o.u; v_new = o.v;
            const bool yown = FREE || (ya >= gy0 && ya < gy1);
            const V2 wout = (o.water * P.evap).v();     // sediment_transport.glsl:75
            const unsigned idx = off - pitch;
            if (L.owned.x && in.x && yown) {
                K.dst[3][idx] = o.fL.x; K.dst[4][idx] = o.fR.x; K.dst[5][idx] = o.fT.x; K.dst[6][idx] = o.fB.x;
                K.dst[2][idx] = wout.x;
            }
            if (L.owned.y && in.y && yown) {
                K.dst[3][idx + HALF] = o.fL.y; K.dst[4][idx + HALF] = o.fR.y; K.dst[5][idx + HALF] = o.fT.y; K.dst[6][idx + HALF] = o.fB.y;
                K.dst[2][idx + HALF] = wout.y;
            }
            HgV2x2 rd;
            rd.a = v2(in.x ? e0.rock : HG_OOB_HEIGHT, in.y ? e1.rock : HG_OOB_HEIGHT);
            rd.b = v2(in.x ? e0.dirt : HG_OOB_HEIGHT, in.y ? e1.dirt : HG_OOB_HEIGHT);
            Q2(R::RD, 1, 0) = rd;
            HgV2x2 s;      // lane-major: (sr, sd) of lane 0, then of lane 1
            s.a = v2(in.x ? e0.sr : 0.0f, in.x ? e0.sd : 0.0f);
            s.b = v2(in.y ? e1.sr : 0.0f, in.y ? e1.sd : 0.0f);
            Q2(R::SS, 1, 0) = s;
        }
    }

    // ------------------------------------------------------------ B(i-3)
    {
        const int yb = i - 3;
        if (FREE || (yb >= gy0 && yb < gy1)) {
            const char* const rm1 = B16(4) + R::SS; const char* const r0 = B16(3) + R::SS; const char* const rp1 = B16(2) + R::SS;
            const HgGather g0 = hg_gather_lane(P, L.x0, yb, W, H, c.u_d2.x, c.v_d2.x, rm1, r0, rp1, 0);
            const HgGather g1 = hg_gather_lane(P, L.x1, yb, W, H, c.u_d2.y, c.v_d2.y, rm1, r0, rp1, 1);
            const unsigned idx = off - 3u * pitch;
            if (L.owned.x) {
                if (g0.fast) { K.dst[7][idx] = g0.sr; K.dst[8][idx] = g0.sd; }
                else {
                    unsigned long long slot = HGF_ATOMIC_INC64(K.far_count);
                    K.far_list[slot] = (unsigned)(yb - K.row0) * (unsigned)W + (unsigned)L.x0;
                }
            }
            if (L.owned.y) {
                if (g1.fast) { K.dst[7][idx + HALF] = g1.sr; K.dst[8][idx + HALF] = g1.sd; }
                else {
                    unsigned long long slot = HGF_ATOMIC_INC64(K.far_count);
                    K.far_list[slot] = (unsigned)(yb - K.row0) * (unsigned)W + (unsigned)L.x1;
                }
            }
        }
    }
    c.u_d2 = c.u_d1; c.v_d2 = c.v_d1; c.u_d1 = u_new; c.v_d1 = v_new;
    }   // GROUP != HGF_THERMAL

    if (GROUP != HGF_HYDRO) {
    // ------------------------------------------------------------ C(i-3), D(i-5)
    {
        const HgV2x2 rd01 = Q2(R::RD, 4, 0);
        const V2 e00 = Q2X(R::RD, 4, -1), e01 = rd01.a, e02 = Q2X(R::RD, 4, 1);
        const V2 e10 = Q2X(R::RD, 3, -1), e11 = Q2X(R::RD, 3, 0), e12 = Q2X(R::RD, 3, 1);
        const V2 e20 = Q2X(R::RD, 2, -1), e21 = Q2X(R::RD, 2, 0), e22 = Q2X(R::RD, 2, 1);
        const V2 rockE_d = c.e_old, dirtE_d0 = c.de_old;
        c.e_old = rd01.a; c.de_old = rd01.b;
        const int yc = i - 3;
        V2 so0 = v2s(0.0f), T0 = v2s(0.0f), B0 = v2s(0.0f);
        if (FREE || (yc >= gy0 - 4 && yc < gy1 + 4)) {
            const B2 in = L.xin && (FREE || (yc >= 0 && yc < H));
            V2 out[8], d_h[8];
            d_h[0] = e11 - e10; d_h[1] = e11 - e12; d_h[2] = e11 - e21; d_h[3] = e11 - e01;
            d_h[4] = e11 - e20; d_h[5] = e11 - e22; d_h[6] = e11 - e00; d_h[7] = e11 - e02;
            so0 = hg_thermal_outflow2(P, 0, e11, d_h, out, in);
            T0 = out[2]; B0 = out[3];
            HgV2x4 tr; tr.a = out[1]; tr.b = out[5]; tr.c = out[7]; tr.d = v2s(0.0f);    // R, RT, RB
            HgV2x4 tl; tl.a = out[0]; tl.b = out[4]; tl.c = out[6]; tl.d = v2s(0.0f);    // L, LT, LB
            Q4(R::O0R, 3, 0) = tr;
            Q4(R::O0L, 3, 0) = tl;
        }
        const int yd = i - 5;
        const HgV2x2 nl = Q4AB(R::O0R, 4, -1); const V2 nl_c = Q4C(R::O0R, 4, -1);     // left neighbour's R, RT | RB
        const HgV2x2 nr = Q4AB(R::O0L, 4, 1); const V2 nr_c = Q4C(R::O0L, 4, 1);       // right neighbour's L, LT | LB
        if (FREE || (yd >= gy0 - 3 && yd < gy1 + 3)) {
            const B2 in = L.xin && (FREE || (yd >= 0 && yd < H));
            const V2 delta = hg_thermal_delta2(c.so0_d2, c.nR0_d1, c.nL0_d1, c.B0_d1, c.T0_d3, nl_c, nr_c, c.nRT0_d2, c.nLT0_d2);
            HgV2x2 w; w.a = v2_sel(in, rockE_d + delta, HG_OOB_HEIGHT); w.b = dirtE_d0;
            Q2(R::R1D, 5, 0) = w;
        }
        c.so0_d2 = c.so0_d1; c.so0_d1 = so0;
        c.T0_d3 = c.T0_d2; c.T0_d2 = c.T0_d1; c.T0_d1 = T0;
        c.B0_d1 = B0;
        c.nR0_d1 = nl.a; c.nL0_d1 = nr.a;
        c.nRT0_d2 = c.nRT0_d1; c.nRT0_d1 = nl.b; c.nLT0_d2 = c.nLT0_d1; c.nLT0_d1 = nr.b;
    }

    // ------------------------------------------------------------ E(i-7), F(i-9)
    {
        const HgV2x2 w00 = Q2(R::R1D, 8, -1), w01 = Q2(R::R1D, 8, 0), w02 = Q2(R::R1D, 8, 1);
        const HgV2x2 w10 = Q2(R::R1D, 7, -1), w11 = Q2(R::R1D, 7, 0), w12 = Q2(R::R1D, 7, 1);
        const HgV2x2 w20 = Q2(R::R1D, 6, -1), w21 = Q2(R::R1D, 6, 0), w22 = Q2(R::R1D, 6, 1);
        const V2 rock1_d = c.p_old, dirtE_d = c.q_old;
        c.p_old = w01.a; c.q_old = w01.b;
        const int ye = i - 7;
        V2 so1 = v2s(0.0f), T1 = v2s(0.0f), B1 = v2s(0.0f);
        if (FREE || (ye >= gy0 - 2 && ye < gy1 + 2)) {
            const B2 in = L.xin && (FREE || (ye >= 0 && ye < H));
            V2 out[8], d_h[8];
            d_h[0] = (w11.a - w10.a) + (w11.b - w10.b); d_h[1] = (w11.a - w12.a) + (w11.b - w12.b);
            d_h[2] = (w11.a - w21.a) + (w11.b - w21.b); d_h[3] = (w11.a - w01.a) + (w11.b - w01.b);
            d_h[4] = (w11.a - w20.a) + (w11.b - w20.b); d_h[5] = (w11.a - w22.a) + (w11.b - w22.b);
            d_h[6] = (w11.a - w00.a) + (w11.b - w00.b); d_h[7] = (w11.a - w02.a) + (w11.b - w02.b);
            so1 = hg_thermal_outflow2(P, 1, w11.b, d_h, out, in);
            T1 = out[2]; B1 = out[3];
            HgV2x4 tr; tr.a = out[1]; tr.b = out[5]; tr.c = out[7]; tr.d = v2s(0.0f);
            HgV2x4 tl; tl.a = out[0]; tl.b = out[4]; tl.c = out[6]; tl.d = v2s(0.0f);
            Q4(R::O1R, 7, 0) = tr;
            Q4(R::O1L, 7, 0) = tl;
        }
        const int yf = i - 9;
        const HgV2x2 nl = Q4AB(R::O1R, 8, -1); const V2 nl_c = Q4C(R::O1R, 8, -1);
        const HgV2x2 nr = Q4AB(R::O1L, 8, 1); const V2 nr_c = Q4C(R::O1L, 8, 1);
        if (FREE || (yf >= gy0 - 1 && yf < gy1 + 1)) {
            const B2 in = L.xin && (FREE || (yf >= 0 && yf < H));
            const V2 delta = hg_thermal_delta2(c.so1_d2, c.nR1_d1, c.nL1_d1, c.B1_d1, c.T1_d3, nl_c, nr_c, c.nRT1_d2, c.nLT1_d2);
            HgV2x2 w; w.a = rock1_d; w.b = v2_sel(in, dirtE_d + delta, HG_OOB_HEIGHT);
            Q2(R::G2, 9, 0) = w;
        }
        c.so1_d2 = c.so1_d1; c.so1_d1 = so1;
        c.T1_d3 = c.T1_d2; c.T1_d2 = c.T1_d1; c.T1_d1 = T1;
        c.B1_d1 = B1;
        c.nR1_d1 = nl.a; c.nL1_d1 = nr.a;
        c.nRT1_d2 = c.nRT1_d1; c.nRT1_d1 = nl.b; c.nLT1_d2 = c.nLT1_d1; c.nLT1_d1 = nr.b;
    }

    // ------------------------------------------------------------ G(i-11)
    {
        const int yg = i - HGF_LAG_G;
        if (FREE || (yg >= gy0 && yg < gy1)) {
            const HgV2x2 l = Q2(R::G2, 11, -1), r = Q2(R::G2, 11, 1);
            const HgV2x2 dn = Q2(R::G2, 12, 0), own = Q2(R::G2, 11, 0), up = Q2(R::G2, 10, 0);
            V2 sr_ = own.a, sd_ = own.b;
            hg_smooth_cell2(P, sr_, sd_, l.a, l.b, r.a, r.b, up.a, up.b, dn.a, dn.b);
            const bool yborder = !FREE && (yg == 0 || yg == H - 1);
            const unsigned idx = off - (unsigned)HGF_LAG_G * pitch;
            if (L.owned.x) {
                const bool bd = L.border.x || yborder;
                K.dst[0][idx] = bd ? own.a.x : sr_.x;
                K.dst[1][idx] = bd ? own.b.x : sd_.x;
            }
            if (L.owned.y) {
                const bool bd = L.border.y || yborder;
                K.dst[0][idx + HALF] = bd ? own.a.y : sr_.y;
                K.dst[1][idx + HALF] = bd ? own.b.y : sd_.y;
            }
        }
    }
    }   // GROUP != HGF_HYDRO
#undef A32
#undef A8
#undef B16
#undef Q4
#undef Q4AB
#undef Q4C
#undef Q1
#undef Q2
#undef Q2X
}
