// emul.cpp — host build of the product's per-cell arithmetic (hydro_gen_b200/csrc/hg_cell.cuh,
// hg_noise.cuh compiled as plain C++, no CUDA) driven by simple loops over SoA planes.
// TEST INFRASTRUCTURE for `-m "not gpu"`: it checks, on a machine without a GPU, that the
// device functions reproduce the oracle bit for bit.  It is not a CPU fallback: nothing in
// the product links or loads it.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../hydro_gen_b200/csrc/hg_cell.cuh"
#include "../../hydro_gen_b200/csrc/hg_noise.cuh"
#include "../../hydro_gen_b200/csrc/hg_fused_body.cuh"
#include "../../hydro_gen_b200/csrc/hg_fused_body2.cuh"
#include "../../hydro_gen_b200/csrc/hg_fused_body3.cuh"

namespace {
struct Dom { int W, H; };
inline bool oob(const Dom& d, int x, int y) { return x < 0 || x > d.W - 1 || y < 0 || y > d.H - 1; }
inline float ld(const float* p, const Dom& d, int x, int y, float o) { return oob(d, x, y) ? o : p[(size_t)y * d.W + x]; }
}

extern "C" {

// planes: rock dirt water fL fR fT fB sr sd, each W*H, updated in place by one grid step
// (the sequence of src/erosion.cpp:158-200).  Returns number of cells whose back-trace
// footprint left the +-1 window (what the fused kernel's far-fetch path handles).
long emul_grid_step(const hg_erosion_data* set, int W, int H, float* pl[9]) {
    HgStepParams P = hg_make_step_params(*set);
    Dom d{W, H};
    size_t n = (size_t)W * H;
    std::vector<float> a(n), u(n), v(n), vz(n), nw(n), nf[4], er(n), ed(n), esr(n), esd(n);
    for (auto& f : nf) f.resize(n);
    float *rock = pl[0], *dirt = pl[1], *water = pl[2], *sr = pl[7], *sd = pl[8];
    for (size_t i = 0; i < n; i++) a[i] = rock[i] + dirt[i] + water[i];
    // flux
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        size_t i = (size_t)y * W + x;
        HgFluxOut o = hg_flux_cell(P, x, y, W, H, a[i], ld(a.data(), d, x - 1, y, HG_OOB_HEIGHT), ld(a.data(), d, x + 1, y, HG_OOB_HEIGHT),
            ld(a.data(), d, x, y + 1, HG_OOB_HEIGHT), ld(a.data(), d, x, y - 1, HG_OOB_HEIGHT),
            pl[3][i], pl[4][i], pl[5][i], pl[6][i],
            ld(pl[4], d, x - 1, y, 0), ld(pl[3], d, x + 1, y, 0), ld(pl[6], d, x, y + 1, 0), ld(pl[5], d, x, y - 1, 0), water[i]);
        nf[0][i] = o.fL; nf[1][i] = o.fR; nf[2][i] = o.fT; nf[3][i] = o.fB;
        nw[i] = o.water; u[i] = o.u; v[i] = o.v; vz[i] = o.vz;
    }
    for (int k = 0; k < 4; k++) memcpy(pl[3 + k], nf[k].data(), n * 4);
    // erosion
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        size_t i = (size_t)y * W + x;
        HgEroOut e = hg_erosion_cell(P, rock[i], dirt[i], sr[i], sd[i], u[i], v[i], vz[i],
            ld(rock, d, x + 1, y, 0), ld(dirt, d, x + 1, y, 0), ld(rock, d, x - 1, y, 0), ld(dirt, d, x - 1, y, 0),
            ld(rock, d, x, y - 1, 0), ld(dirt, d, x, y - 1, 0), ld(rock, d, x, y + 1, 0), ld(dirt, d, x, y + 1, 0));
        er[i] = e.rock; ed[i] = e.dirt; esr[i] = e.sr; esd[i] = e.sd;
    }
    // sediment transport + evaporation
    long far = 0;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        size_t i = (size_t)y * W + x;
        HgBack b = hg_backtrace(P, x, y, W, H, u[i], v[i]);
        int dx = b.px - x, dy = b.py - y;
        if (!(dx >= -1 && dx <= 0 && dy >= -1 && dy <= 0)) far++;
        sr[i] = hg_bilerp(ld(esr.data(), d, b.px, b.py, 0), ld(esr.data(), d, b.px + 1, b.py, 0),
                          ld(esr.data(), d, b.px, b.py + 1, 0), ld(esr.data(), d, b.px + 1, b.py + 1, 0), b.sx, b.sy);
        sd[i] = hg_bilerp(ld(esd.data(), d, b.px, b.py, 0), ld(esd.data(), d, b.px + 1, b.py, 0),
                          ld(esd.data(), d, b.px, b.py + 1, 0), ld(esd.data(), d, b.px + 1, b.py + 1, 0), b.sx, b.sy);
        water[i] = nw[i] * P.evap;
    }
    // thermal, layers 0 then 1
    const int ox[8] = {-1, 1, 0, 0, -1, 1, -1, 1}, oy[8] = {0, 0, 1, -1, 1, 1, -1, -1};
    std::vector<float> out[8], neg(n);
    for (auto& o : out) o.resize(n);
    for (int layer = 0; layer < 2; layer++) {
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            size_t i = (size_t)y * W + x;
            float d_h[8], o8[8];
            for (int k = 0; k < 8; k++) {
                float dh = 0.0f;
                dh += er[i] - ld(er.data(), d, x + ox[k], y + oy[k], HG_OOB_HEIGHT);
                if (layer == 1) dh += ed[i] - ld(ed.data(), d, x + ox[k], y + oy[k], HG_OOB_HEIGHT);
                d_h[k] = dh;
            }
            neg[i] = hg_thermal_outflow(P, layer, layer == 0 ? er[i] : ed[i], d_h, o8);
            for (int k = 0; k < 8; k++) out[k][i] = o8[k];
        }
        std::vector<float>& tgt = layer == 0 ? er : ed;
        std::vector<float> nt(n);
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            size_t i = (size_t)y * W + x;
            float delta = hg_thermal_delta(neg[i], ld(out[1].data(), d, x - 1, y, 0), ld(out[0].data(), d, x + 1, y, 0),
                ld(out[3].data(), d, x, y + 1, 0), ld(out[2].data(), d, x, y - 1, 0),
                ld(out[7].data(), d, x - 1, y + 1, 0), ld(out[6].data(), d, x + 1, y + 1, 0),
                ld(out[5].data(), d, x - 1, y - 1, 0), ld(out[4].data(), d, x + 1, y - 1, 0));
            nt[i] = tgt[i] + delta;
        }
        tgt.swap(nt);
    }
    // smoothing
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        size_t i = (size_t)y * W + x;
        float r = er[i], g = ed[i];
        if (!(x == 0 || y == 0 || x == W - 1 || y == H - 1))
            hg_smooth_cell(P, r, g, er[i - 1], ed[i - 1], er[i + 1], ed[i + 1], er[i + W], ed[i + W], er[i - W], ed[i - W]);
        rock[i] = r; dirt[i] = g;
    }
    return far;
}

void emul_heightmap(const hg_map_settings_data* cfg, int W, int H, float* rock, float* dirt) {
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) hg_heightmap_cell(*cfg, x, y, W, H, rock[(size_t)y * W + x], dirt[(size_t)y * W + x]);
}

// table form of gln_simplex (hg_simplex_tab, what k_rain runs) against the float form on n points: number of mismatches
long emul_simplex_tab_mismatches(const float* xs, const float* ys, long n) {
    static int ti[HG_PERM_N];
    static float tf[HG_PERM_N];
    static HgGrad tg[HG_PERM_N];
    for (int k = 0; k < HG_PERM_N; k++) { tf[k] = hg_permute((float)k); ti[k] = (int)tf[k]; tg[k] = hg_simplex_grad(tf[k]); }
    HgPermTab T{ti, tf, tg};
    long bad = 0;
    for (long i = 0; i < n; i++) {
        float a = hg_simplex(xs[i], ys[i]), b = hg_simplex_tab(xs[i], ys[i], T);
        if (memcmp(&a, &b, 4) != 0) bad++;
    }
    return bad;
}

void emul_rain(const hg_rain_data* set, const hg_map_settings_data* map_set, float time, int W, int H,
               const float* rock, const float* dirt, float* water) {
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        size_t i = (size_t)y * W + x;
        water[i] += hg_rain_cell(*set, *map_set, time, x, y, rock[i] + dirt[i] + water[i]);
    }
}
}

// The fused kernel's body (hg_fused_body.cuh) run thread by thread: CTAs one after the other,
// inside a CTA all threads execute iteration i before any executes i+1 (the kernel's one
// barrier per row), with the same generic / FREE iteration plan as k_fused_step.
// src/dst: 9 planes of W*H floats (no ghost rows).  far_out receives the local cell indices
// whose sediment the kernel leaves to the far-fetch fix-up; returns their number.
// ws = 0: one thread per column runs all stages (k_fused_step).  ws = 1 / 2: the warp-specialised
// form (k_fused_ws): a hydraulic thread and a thermal thread per column, each with its own column
// history; within an iteration all hydraulic threads run before all thermal threads (ws = 1) or
// after them (ws = 2) -- on the GPU the two groups run concurrently between two barriers, so both
// orders must give the same bits (any same-iteration hand-off through the rings would break one).
template <int NT>
static long fused_step_emul(const hg_erosion_data* set, int W, int H, int seg, int ws, const float* const src[9], float* const dst[9], unsigned* far_out) {
    const int HALO = 8;
    size_t pe = (size_t)(H + 2 * HALO) * W;
    std::vector<std::vector<float>> ps(9, std::vector<float>(pe, 0.0f)), pd(9, std::vector<float>(pe, 0.0f));
    for (int p = 0; p < 9; p++) memcpy(ps[p].data() + (size_t)HALO * W, src[p], (size_t)W * H * 4);
    unsigned long long far_count = 0;
    HgFusedK K;
    memset(&K, 0, sizeof(K));
    for (int p = 0; p < 9; p++) { K.src[p] = ps[p].data(); K.dst[p] = pd[p].data(); }
    K.W = W; K.H = H; K.row0 = 0; K.rows = H; K.pitch = W; K.seg = seg;
    K.nstrips = (W + (NT - 12) - 1) / (NT - 12);
    K.far_list = far_out; K.far_count = &far_count;
    K.P = hg_make_step_params(*set);
    int nseg = (H + seg - 1) / seg;
    std::vector<float> sm(HgRings<NT>::TOTAL + 4);
    std::vector<HgCol> cols(NT), colsT(NT);
    for (int blk = 0; blk < K.nstrips * nseg; blk++) {
        std::fill(sm.begin(), sm.end(), 0.0f);
        float* smp = sm.data();
        while ((uintptr_t)smp % 16) smp++;
        int strip = blk % K.nstrips, segi = blk / K.nstrips;
        int gy0 = segi * seg, gy1 = gy0 + seg < H ? gy0 + seg : H;
        HgFusedPlan pl = hg_fused_plan(gy0, gy1, H);
        auto xof = [&](int tid) { return strip * (NT - 12) - 6 + tid; };
        auto offof = [&](int tid, int i) { return (unsigned)(i + HALO) * (unsigned)W + (unsigned)xof(tid); };
        for (int tid = 0; tid < NT; tid++) { hg_fused_begin(cols[tid]); hg_fused_begin(colsT[tid]); }
        std::vector<float> raw(9 * HGF_RAW_LD(NT));
        for (int i = pl.i_begin; i <= pl.i_end; i++) {
            bool fr = i >= pl.free_lo && i <= pl.free_hi;
            // what the kernel's TMA box load delivers: plane rows with zero fill outside the allocation
            for (int p = 0; p < 9; p++) for (int t = 0; t < HGF_RAW_LD(NT); t++) {
                int x = xof(0) - 2 + t, lr = i + HALO;
                raw[p * HGF_RAW_LD(NT) + t] = (x >= 0 && x < W && lr >= 0 && lr < H + 2 * HALO) ? ps[p][(size_t)lr * W + x] : 0.0f;
            }
            auto run_group = [&](int group) {
                for (int tid = 0; tid < NT; tid++) {
                    int x = xof(tid);
                    bool xin = x >= 0 && x < W, owned = tid >= 6 && tid < NT - 6 && x < W;
                    unsigned off = offof(tid, i);
                    if (group == HGF_ALL) {
                        if (fr) hg_fused_iter<NT, true>(cols[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, off);
                        else hg_fused_iter<NT, false>(cols[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, off);
                    } else if (group == HGF_HYDRO) {
                        if (fr) hg_fused_iter<NT, true, HGF_HYDRO>(cols[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, off);
                        else hg_fused_iter<NT, false, HGF_HYDRO>(cols[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, off);
                    } else {
                        if (fr) hg_fused_iter<NT, true, HGF_THERMAL>(colsT[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, off);
                        else hg_fused_iter<NT, false, HGF_THERMAL>(colsT[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, off);
                    }
                }
            };
            if (ws == 0) run_group(HGF_ALL);
            else if (ws == 1) { run_group(HGF_HYDRO); run_group(HGF_THERMAL); }
            else { run_group(HGF_THERMAL); run_group(HGF_HYDRO); }
        }
    }
    for (int p = 0; p < 9; p++) memcpy(dst[p], pd[p].data() + (size_t)HALO * W, (size_t)W * H * 4);
    return (long)far_count;
}

extern "C" long emul_fused_step(const hg_erosion_data* set, int W, int H, int nt, int seg, int ws, const float* const src[9], float* const dst[9], unsigned* far_out) {
    if (nt == 32) return fused_step_emul<32>(set, W, H, seg, ws, src, dst, far_out);
    if (nt == 128) return fused_step_emul<128>(set, W, H, seg, ws, src, dst, far_out);
    if (nt == 224) return fused_step_emul<224>(set, W, H, seg, ws, src, dst, far_out);
    return -1;
}


// The two-columns-per-thread body (hg_fused_body2.cuh, k_fused_ws2) run the same way: a CTA owns a pair of strips,
// thread t carries column t of both in the two lanes of a V2 (scalar pairs on the host).  ws as above.
template <int NT>
static long fused2_step_emul(const hg_erosion_data* set, int W, int H, int seg, int ws, const float* const src[9], float* const dst[9], unsigned* far_out) {
    const int HALO = 8;
    constexpr int HALF = HGF2_HALF(NT), LD = HGF2_RAW_LD(NT);
    size_t pe = (size_t)(H + 2 * HALO) * W;
    std::vector<std::vector<float>> ps(9, std::vector<float>(pe, 0.0f)), pd(9, std::vector<float>(pe, 0.0f));
    for (int p = 0; p < 9; p++) memcpy(ps[p].data() + (size_t)HALO * W, src[p], (size_t)W * H * 4);
    unsigned long long far_count = 0;
    HgFusedK K;
    memset(&K, 0, sizeof(K));
    for (int p = 0; p < 9; p++) { K.src[p] = ps[p].data(); K.dst[p] = pd[p].data(); }
    K.W = W; K.H = H; K.row0 = 0; K.rows = H; K.pitch = W; K.seg = seg;
    K.nstrips = (W + 2 * HALF - 1) / (2 * HALF);
    K.far_list = far_out; K.far_count = &far_count;
    K.P = hg_make_step_params(*set);
    int nseg = (H + seg - 1) / seg;
    std::vector<float> sm(HgRings2<NT>::TOTAL + 8);
    std::vector<HgCol2> cols(NT), colsT(NT);
    for (int blk = 0; blk < K.nstrips * nseg; blk++) {
        std::fill(sm.begin(), sm.end(), 0.0f);
        float* smp = sm.data();
        while ((uintptr_t)smp % 16) smp++;
        int strip = blk % K.nstrips, segi = blk / K.nstrips;
        int gy0 = segi * seg, gy1 = gy0 + seg < H ? gy0 + seg : H;
        HgFusedPlan pl = hg_fused_plan(gy0, gy1, H);
        const int x0 = strip * (2 * HALF) - HGF_HX;
        for (int tid = 0; tid < NT; tid++) { hg_col2_init(cols[tid]); hg_col2_init(colsT[tid]); }
        std::vector<float> raw(9 * LD);
        for (int i = pl.i_begin; i <= pl.i_end; i++) {
            bool fr = i >= pl.free_lo && i <= pl.free_hi;
            for (int p = 0; p < 9; p++) for (int t = 0; t < LD; t++) {
                int x = x0 - 2 + t, lr = i + HALO;
                raw[p * LD + t] = (x >= 0 && x < W && lr >= 0 && lr < H + 2 * HALO) ? ps[p][(size_t)lr * W + x] : 0.0f;
            }
            auto run_group = [&](int group) {
                for (int tid = 0; tid < NT; tid++) {
                    const HgLanes L = hg_lanes(x0 + tid, HALF, tid, NT, W);
                    unsigned off = (unsigned)(i + HALO) * (unsigned)W + (unsigned)(x0 + tid);
                    if (group == HGF_ALL) {
                        if (fr) hg_fused_iter2<NT, true, HGF_ALL>(cols[tid], smp, raw.data(), K, tid, L, gy0, gy1, i, off);
                        else hg_fused_iter2<NT, false, HGF_ALL>(cols[tid], smp, raw.data(), K, tid, L, gy0, gy1, i, off);
                    } else if (group == HGF_HYDRO) {
                        if (fr) hg_fused_iter2<NT, true, HGF_HYDRO>(cols[tid], smp, raw.data(), K, tid, L, gy0, gy1, i, off);
                        else hg_fused_iter2<NT, false, HGF_HYDRO>(cols[tid], smp, raw.data(), K, tid, L, gy0, gy1, i, off);
                    } else {
                        if (fr) hg_fused_iter2<NT, true, HGF_THERMAL>(colsT[tid], smp, raw.data(), K, tid, L, gy0, gy1, i, off);
                        else hg_fused_iter2<NT, false, HGF_THERMAL>(colsT[tid], smp, raw.data(), K, tid, L, gy0, gy1, i, off);
                    }
                }
            };
            if (ws == 0) run_group(HGF_ALL);
            else if (ws == 1) { run_group(HGF_HYDRO); run_group(HGF_THERMAL); }
            else { run_group(HGF_THERMAL); run_group(HGF_HYDRO); }
        }
    }
    for (int p = 0; p < 9; p++) memcpy(dst[p], pd[p].data() + (size_t)HALO * W, (size_t)W * H * 4);
    return (long)far_count;
}

extern "C" long emul_fused2_step(const hg_erosion_data* set, int W, int H, int nt, int seg, int ws, const float* const src[9], float* const dst[9], unsigned* far_out) {
    if (nt == 32) return fused2_step_emul<32>(set, W, H, seg, ws, src, dst, far_out);
    if (nt == 128) return fused2_step_emul<128>(set, W, H, seg, ws, src, dst, far_out);
    return -1;
}

// The queued-outflow body (hg_fused_body3.cuh, k_fused_q) run the same way: hydraulic threads, thermal threads and the
// service role (the cells queued during the previous iteration) within an iteration in the order `ws` names --
// 1: H T S, 2: S T H, 3: T S H, 4: S H T -- all of which the GPU may produce between two barriers.  The queue starts each
// CTA empty, smem starts as garbage-free zeros here (the kernel never reads an element it did not write for a cell
// whose result is used).
template <int NT>
static long fusedq_step_emul(const hg_erosion_data* set, int W, int H, int seg, int ws, const float* const src[9], float* const dst[9], unsigned* far_out) {
    const int HALO = 8;
    typedef HgRingsQ<NT> R;
    size_t pe = (size_t)(H + 2 * HALO) * W;
    std::vector<std::vector<float>> ps(9, std::vector<float>(pe, 0.0f)), pd(9, std::vector<float>(pe, 0.0f));
    for (int p = 0; p < 9; p++) memcpy(ps[p].data() + (size_t)HALO * W, src[p], (size_t)W * H * 4);
    unsigned long long far_count = 0;
    HgFusedK K;
    memset(&K, 0, sizeof(K));
    for (int p = 0; p < 9; p++) { K.src[p] = ps[p].data(); K.dst[p] = pd[p].data(); }
    K.W = W; K.H = H; K.row0 = 0; K.rows = H; K.pitch = W; K.seg = seg;
    K.nstrips = (W + (NT - 12) - 1) / (NT - 12);
    K.far_list = far_out; K.far_count = &far_count;
    K.P = hg_make_step_params(*set);
    int nseg = (H + seg - 1) / seg;
    std::vector<float> sm(R::TOTAL + 4);
    std::vector<HgColQ> cols(NT), colsT(NT);
    for (int blk = 0; blk < K.nstrips * nseg; blk++) {
        // poison instead of zeros: a value read before it was written for a cell that matters shows up as a mismatch
        for (auto& f : sm) f = 12345.678f;
        float* smp = sm.data();
        while ((uintptr_t)smp % 16) smp++;
        unsigned* qcnt = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(smp) + R::QCNT);
        const unsigned short* qbuf = reinterpret_cast<const unsigned short*>(reinterpret_cast<char*>(smp) + R::QUEUE);
        for (int k = 0; k < 4; k++) qcnt[k] = 0u;
        int strip = blk % K.nstrips, segi = blk / K.nstrips;
        int gy0 = segi * seg, gy1 = gy0 + seg < H ? gy0 + seg : H;
        HgFusedPlan pl = hg_fusedq_plan(gy0, gy1, H);
        auto xof = [&](int tid) { return strip * (NT - 12) - 6 + tid; };
        for (int tid = 0; tid < NT; tid++) { hg_colq_init(cols[tid]); hg_colq_init(colsT[tid]); }
        std::vector<float> raw(9 * HGF_RAW_LD(NT));
        for (int i = pl.i_begin; i <= pl.i_end; i++) {
            bool fr = i >= pl.free_lo && i <= pl.free_hi;
            const int m3 = ((i - 3) % 3 + 3) % 3;
            for (int p = 0; p < 9; p++) for (int t = 0; t < HGF_RAW_LD(NT); t++) {
                int x = xof(0) - 2 + t, lr = i + HALO;
                raw[p * HGF_RAW_LD(NT) + t] = (x >= 0 && x < W && lr >= 0 && lr < H + 2 * HALO) ? ps[p][(size_t)lr * W + x] : 0.0f;
            }
            auto run_group = [&](int group) {
                if (group == HGQ_SERVICE) {
                    const unsigned n = qcnt[(i - 1) & 3];
                    for (unsigned b = 0; b < n; b++) hg_fusedq_serve<NT>(smp, K, i, m3, qbuf[((i - 1) & 1) * (2 * NT) + b]);
                    qcnt[(i + 1) & 3] = 0u;
                    return;
                }
                for (int tid = 0; tid < NT; tid++) {
                    int x = xof(tid);
                    bool xin = x >= 0 && x < W, owned = tid >= 6 && tid < NT - 6 && x < W;
                    unsigned off = (unsigned)(i + HALO) * (unsigned)W + (unsigned)x;
                    if (group == HGF_HYDRO) {
                        if (fr) hg_fusedq_iter<NT, true, HGF_HYDRO>(cols[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, m3, off);
                        else hg_fusedq_iter<NT, false, HGF_HYDRO>(cols[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, m3, off);
                    } else {
                        if (fr) hg_fusedq_iter<NT, true, HGF_THERMAL>(colsT[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, m3, off);
                        else hg_fusedq_iter<NT, false, HGF_THERMAL>(colsT[tid], smp, raw.data(), K, tid, x, xin, owned, gy0, gy1, i, m3, off);
                    }
                }
            };
            static const int order[5][3] = {{HGF_HYDRO, HGF_THERMAL, HGQ_SERVICE}, {HGF_HYDRO, HGF_THERMAL, HGQ_SERVICE}, {HGQ_SERVICE, HGF_THERMAL, HGF_HYDRO},
                                            {HGF_THERMAL, HGQ_SERVICE, HGF_HYDRO}, {HGQ_SERVICE, HGF_HYDRO, HGF_THERMAL}};
            for (int g = 0; g < 3; g++) run_group(order[ws >= 0 && ws < 5 ? ws : 0][g]);
        }
    }
    for (int p = 0; p < 9; p++) memcpy(dst[p], pd[p].data() + (size_t)HALO * W, (size_t)W * H * 4);
    return (long)far_count;
}

extern "C" long emul_fusedq_step(const hg_erosion_data* set, int W, int H, int nt, int seg, int ws, const float* const src[9], float* const dst[9], unsigned* far_out) {
    if (nt == 32) return fusedq_step_emul<32>(set, W, H, seg, ws, src, dst, far_out);
    if (nt == 128) return fusedq_step_emul<128>(set, W, H, seg, ws, src, dst, far_out);
    return -1;
}

// ---- the balanced partition's arithmetic (hg_plan.cuh), as k_plan_segments applies it: plan in, plan out ----
#include "../../hydro_gen_b200/csrc/hg_plan.cuh"
extern "C" int emul_plan(int n_cta, int nstrips, int row0, int rows, int min_rows, const int* old_plan /* n_cta x (strip, gy0, gy1) */,
                         const unsigned* cta_ns, int* new_plan /* n_cta x (strip, gy0, gy1) */) {
    std::vector<HgPlanItem> old(n_cta), out(n_cta);
    for (int b = 0; b < n_cta; b++) { old[b].strip = old_plan[3 * b]; old[b].gy0 = old_plan[3 * b + 1]; old[b].gy1 = old_plan[3 * b + 2]; old[b].pad = 0; }
    std::vector<int> first(nstrips + 1, 0), new_n(nstrips), new_first(nstrips + 1);
    std::vector<float> cost(nstrips, 0.0f), frac(nstrips);
    for (int b = 0; b < n_cta; b++) if (b == 0 || old[b].strip != old[b - 1].strip) first[old[b].strip] = b;
    first[nstrips] = n_cta;
    for (int s = 0; s < nstrips; s++) for (int k = first[s]; k < first[s + 1]; k++) cost[s] += hg_plan_cost(cta_ns[k]);
    hg_plan_apportion(n_cta, nstrips, cost.data(), rows / min_rows > 1 ? rows / min_rows : 1, new_n.data(), frac.data());
    int b = 0;
    for (int s = 0; s < nstrips; s++) { new_first[s] = b; b += new_n[s]; }
    if (b != n_cta) return -1;
    for (int s = 0; s < nstrips; s++)
        hg_plan_cut_strip(s, new_n[s], old.data() + first[s], cta_ns + first[s], first[s + 1] - first[s], row0, rows, min_rows, out.data() + new_first[s]);
    for (int k = 0; k < n_cta; k++) { new_plan[3 * k] = out[k].strip; new_plan[3 * k + 1] = out[k].gy0; new_plan[3 * k + 2] = out[k].gy1; }
    return 0;
}
