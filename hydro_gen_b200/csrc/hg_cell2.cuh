// hg_cell2.cuh — the per-cell arithmetic of hg_cell.cuh for TWO cells at once (V2, hg_v2.cuh).
//
// Same operations in the same order per lane, so each lane reproduces hg_cell.cuh bit for bit (tests/host_emul runs
// both against the oracle).  Additions, subtractions and multiplications become packed instructions; min/max,
// compares and selects stay per lane (there is no packed form); divisions, square roots and the arctangent are
// evaluated per lane by the scalar code; short data-dependent branches become selects where both sides are a few
// packed operations, and stay branches (taken if EITHER lane needs them) where they are long and rare.
#pragma once
#include "hg_v2.cuh"

// ---------------------------------------------------------------- hydro_flux.glsl:77-166
struct HgFluxOut2 { V2 fL, fR, fT, fB, water, u, v, vz; };

// x_left: x <= 0, x_right: x >= W-1 and not x_left (hydro_flux.glsl:110-113), per lane; y, H as in hg_flux_cell (both lanes share the row).
HG_FN HgFluxOut2 hg_flux_cell2(const HgStepParams& P, B2 x_left, B2 x_right, int y, int H,
                               V2 a, V2 aL, V2 aR, V2 aT, V2 aB, V2 fL, V2 fR, V2 fT, V2 fB,
                               V2 inL, V2 inR, V2 inT, V2 inB, V2 water) {
    HgFluxOut2 o;
    const V2 d1 = water;
    const V2 dhx = a - aL, dhy = a - aR, dhz = a - aT, dhw = a - aB;
    V2 ox = v2_max_c(0.0f, P.ENERGY_KEPT * fL + P.d_t * (P.G * dhx));
    V2 oy = v2_max_c(0.0f, P.ENERGY_KEPT * fR + P.d_t * (P.G * dhy));
    V2 oz = v2_max_c(0.0f, P.ENERGY_KEPT * fT + P.d_t * (P.G * dhz));
    V2 ow = v2_max_c(0.0f, P.ENERGY_KEPT * fB + P.d_t * (P.G * dhw));
    ox = v2_sel(x_left, 0.0f, ox);
    oy = v2_sel(x_right, 0.0f, oy);
    if (y <= 0) ow = v2s(0.0f);
    else if (y >= H - 1) oz = v2s(0.0f);
    const V2 sum_in = inL + inR + inT + inB;
    V2 sum_out = ox + oy + oz + ow;
    const V2 den = (sum_out * P.d_t).v();
    V2 K;
    K.x = hg_min_c(1.0f, hg_div_zero_num(water.x, den.x));
    K.y = hg_min_c(1.0f, hg_div_zero_num(water.y, den.y));
    ox = (ox * K).v(); oy = (oy * K).v(); oz = (oz * K).v(); ow = (ow * K).v();
    const V2P sum_out_k = sum_out * K;
    const V2P d_volume = P.d_t * (sum_in - sum_out_k);
    const V2 d2 = v2_max_c(0.0f, d1 + d_volume);
    o.fL = ox; o.fR = oy; o.fT = oz; o.fB = ow;
    o.water = d2;
    o.vz = d1 + d2;
    const V2 nu = inL - fL + fR - inR;
    const V2 nv = inB - fB + fT - inT;
    o.u = v2s(0.0f); o.v = v2s(0.0f);
    if (o.vz.x > 0.0f) { o.u.x = nu.x / o.vz.x; o.v.x = nv.x / o.vz.x; }
    if (o.vz.y > 0.0f) { o.u.y = nu.y / o.vz.y; o.v.y = nv.y / o.vz.y; }
    return o;
}

// ------------------------------------------------------------- hydro_erosion.glsl:37-92
// One lane of the part of hg_erosion_cell that is data-dependent control flow: the erosion velocity
// (hydro_erosion.glsl:45-52) ...
HG_FN float hg_ero_vel(float len, float vz) {
    float dd = vz;
    if (dd < 1e-3f) {
        dd = hg_max_c(5e-4f, dd);
        return hg_mix(len, 0.0f, hg_smoothstep(1e-3f, 5e-4f, dd));
    }
    return len;
}
// ... and the layer loop with the Kconv conversion (hydro_erosion.glsl:58-88); kv = Kc * max(0.02, sin_a) * ero_vel.
HG_FN HgEroOut hg_erosion_layers(const HgStepParams& P, float kv, float rock, float dirt, float sr, float sd) {
    float terrain[2] = {rock, dirt};
    float sediment[2] = {sr, sd};
    float cap = 0.0f;
#pragma unroll
    for (int i = HG_SED_LAYERS - 1; i >= 0; i--) {
        float c = hg_max_c(0.0f, kv - cap);
        if (c > sediment[i]) {
            float old_terr = terrain[i];
            float delta = P.Kls[i] * (c - sediment[i]);
            terrain[i] -= delta;
            sediment[i] += delta;
            if (terrain[i] < 0.0f) {
                sediment[i] += terrain[i];
                terrain[i] = 0.0f;
                cap += old_terr;
            } else {
                break;
            }
        } else {
            float delta = P.Kld[i] * (sediment[i] - c);
            terrain[i] += delta;
            sediment[i] -= delta;
        }
    }
    float conv = sediment[0] * P.Kconv * P.d_t;
    sediment[1] += conv;
    sediment[0] -= conv;
    HgEroOut o;
    o.rock = terrain[0]; o.dirt = terrain[1]; o.sr = sediment[0]; o.sd = sediment[1];
    return o;
}

// Both lanes of hg_erosion_cell; results per lane (each lane's (sr, sd) stays a pair for the sediment ring).
HG_FN void hg_erosion_cell2(const HgStepParams& P, V2 rock, V2 dirt, V2 sr, V2 sd, V2 u, V2 v, V2 vz,
                            V2 rR, V2 gR, V2 rL, V2 gL, V2 rB, V2 gB, V2 rT, V2 gT, HgEroOut& o0, HgEroOut& o1) {
    const V2 l2 = u * u + v * v;
    const V2 len = v2(hg_sqrt_pos(l2.x), hg_sqrt_pos(l2.y));
    const V2 ero_vel = v2(hg_ero_vel(len.x, vz.x), hg_ero_vel(len.y, vz.y));
    // get_terr_normal, hydro_erosion.glsl:23-35
    const V2 dx = rR + gR - rL - gL;
    const V2 dz = rT + gT - rB - gB;
    const V2 nx = dx * 2.0f - 0.0f * dz;
    const float ny0 = 0.0f * 0.0f - 2.0f * 2.0f;
    const V2 nz = 2.0f * dz - dx * 0.0f;
    const V2 n2 = nx * nx + ny0 * ny0 + nz * nz;
    const V2 inv = v2(1.0f / sqrtf(n2.x), 1.0f / sqrtf(n2.y));
    const V2 ny = (ny0 * inv).v();
    const V2 s2 = 1.0f - ny * ny;
    const V2 sin_a = v2(fabsf(fabsf(hg_sqrt_pos(s2.x))), fabsf(fabsf(hg_sqrt_pos(s2.y))));
    const V2 kv = (P.Kc * v2_max_c(0.02f, sin_a) * ero_vel).v();
    o0 = hg_erosion_layers(P, kv.x, rock.x, dirt.x, sr.x, sd.x);
    o1 = hg_erosion_layers(P, kv.y, rock.y, dirt.y, sr.y, sd.y);
}

// ---------------------------------------------------------- thermal_erosion.glsl:28-115
// Both lanes of hg_thermal_outflow.  live: the lane's cell is inside the map.  Returns the negated own-outflow sums.
HG_FN V2 hg_thermal_outflow2(const HgStepParams& P, int layer, V2 own, const V2 d_h[8], V2 out[8], B2 live) {
    const float thc = P.th_mark[layer][0], thd = P.th_mark[layer][1];
    const V2 mc = v2_fmax(v2_fmax(d_h[0], d_h[1]), v2_fmax(d_h[2], d_h[3]));
    const V2 md = v2_fmax(v2_fmax(d_h[4], d_h[5]), v2_fmax(d_h[6], d_h[7]));
    const B2 hot = live && (v2_ge(mc, thc) || v2_ge(md, thd));
    if (!b2_any(hot)) {      // nothing is marked in either cell
#pragma unroll
        for (int k = 0; k < 8; k++) out[k] = v2s(0.0f);
        return v2s(0.0f);
    }
    const V2 mm = v2_fmax(mc, md);
    V2 Hm = v2_fmax(v2s(0.0f), mm);
    Hm = v2(hg_min(own.x, Hm.x), hg_min(own.y, Hm.y));
    // bk in the shader's order.  A lane that is not hot is carried along with S = 0: every outflow it computes is
    // selected away or is an exact +0 (marks of an out-of-map lane may be set; its S * d_h is +0 then).
    V2 bk = v2s(0.0f);
    B2 mark[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        mark[k] = v2_ge(d_h[k], k >= 4 ? thd : thc);
        bk = bk + v2_sel(mark[k], d_h[k], 0.0f);        // marked d_h are > 0, so bk is never -0 and bk + 0 == bk
    }
    V2 S;
    {
        float s2[2];
        const float mcs[2] = {mc.x, mc.y}, mds[2] = {md.x, md.y}, hms[2] = {Hm.x, Hm.y};
        const bool hs[2] = {hot.x, hot.y};
#pragma unroll
        for (int l = 0; l < 2; l++) {
            s2[l] = 0.0f;
            if (hs[l]) {
                const float ratio = fmaxf(mcs[l], mds[l] / 1.41421356237309504880f);
                const float alph = hg_atanf(ratio);
                float sharpness = 1.0f;
                const float newsh = 1.0f + alph - P.Kalpha[layer];
                if (newsh > sharpness) sharpness = newsh;
                sharpness *= sharpness * sharpness;
                s2[l] = P.d_t * P.Kspeed[layer] * sharpness * 1.0f * hms[l] / 2.0f;
            }
        }
        S = v2(s2[0], s2[1]);
    }
    V2 neg = v2s(0.0f);
#if HG_DEVICE_FAST
    // shared-reciprocal division, see hg_thermal_outflow: exact inside the guarded range
    const bool ok0 = !hot.x || ((S.x == 0.0f || S.x >= P.th_slo[layer]) && S.x <= 1048576.0f && mm.x <= 1.0995116278e12f);
    const bool ok1 = !hot.y || ((S.y == 0.0f || S.y >= P.th_slo[layer]) && S.y <= 1048576.0f && mm.y <= 1.0995116278e12f);
    if (ok0 && ok1) {
        V2 r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(bk.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(bk.y));
        const V2 nbk = v2_neg(bk);
        r = v2_fma(r, v2_fma(nbk, r, v2s(1.0f)), r);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const V2 a = (S * d_h[k]).v();
            const V2 q0 = (a * r).v();
            const V2 q = v2_fma(r, v2_fma(nbk, q0, a), q0);
            out[k] = v2_sel(mark[k], q, 0.0f);
            neg = neg - out[k];
        }
        // a lane that is not hot must return exact zeros whatever its garbage reciprocal produced
        if (!hot.x) { for (int k = 0; k < 8; k++) out[k].x = 0.0f; neg.x = 0.0f; }
        if (!hot.y) { for (int k = 0; k < 8; k++) out[k].y = 0.0f; neg.y = 0.0f; }
        return neg;
    }
    {   // cold: the generic IEEE divisions, lane by lane (out of line)
        float n2[2];
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const bool h = l ? hot.y : hot.x;
            unsigned mask = 0;
            float d[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { d[k] = l ? d_h[k].y : d_h[k].x; mask |= (unsigned)((l ? mark[k].y : mark[k].x) && h) << k; }
            const HgOut8 g = hg_thermal_outflow_generic(l ? S.y : S.x, l ? bk.y : bk.x, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], mask);
            float ng = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; k++) { if (l) out[k].y = g.v[k]; else out[k].x = g.v[k]; ng -= g.v[k]; }
            n2[l] = ng;
        }
        return v2(n2[0], n2[1]);
    }
#else
#pragma unroll
    for (int k = 0; k < 8; k++) {
        out[k].x = (hot.x && mark[k].x) ? S.x * d_h[k].x / bk.x : 0.0f;
        out[k].y = (hot.y && mark[k].y) ? S.y * d_h[k].y / bk.y : 0.0f;
        neg = neg - out[k];
    }
    return neg;
#endif
}

// thermal_transport.glsl:31-65, both lanes
HG_FN V2 hg_thermal_delta2(V2 neg_out, V2 fromL, V2 fromR, V2 fromT, V2 fromB, V2 fromLT, V2 fromRT, V2 fromLB, V2 fromRB) {
    V2 in_flux = v2s(0.0f);
    in_flux = in_flux + fromL; in_flux = in_flux + fromR; in_flux = in_flux + fromT; in_flux = in_flux + fromB;
    in_flux = in_flux + fromLT; in_flux = in_flux + fromRT; in_flux = in_flux + fromLB; in_flux = in_flux + fromRB;
    return neg_out + in_flux;
}

// ---------------------------------------------------------------- smoothing.glsl:22-75, both lanes (interior cells)
HG_FN V2 hg_div5_2(V2 x) {
#if HG_DEVICE_FAST
    const V2 q = (x * 0.2f).v();
    return v2_fma(v2_fma(v2s(-5.0f), q, x), v2s(0.2f), q);
#else
    return v2(x.x / 5.0f, x.y / 5.0f);
#endif
}
HG_FN B2 hg_smooth_extremum(V2 dl, V2 dr, V2 dt, V2 db, V2 hdiff) {
    const V2 xc = (dl * dr).v(), yc = (dt * db).v();
    B2 m;
    m.x = (((-dl.x) > hdiff.x || (-dr.x) > hdiff.x) && xc.x > 0.0f) || (((-dt.x) > hdiff.x || (-db.x) > hdiff.x) && yc.x > 0.0f);
    m.y = (((-dl.y) > hdiff.y || (-dr.y) > hdiff.y) && xc.y > 0.0f) || (((-dt.y) > hdiff.y || (-db.y) > hdiff.y) && yc.y > 0.0f);
    return m;
}
HG_FN void hg_smooth_cell2(const HgStepParams& P, V2& rock, V2& dirt, V2 lr, V2 lg, V2 rr, V2 rg, V2 tr, V2 tg, V2 br, V2 bg) {
    V2 terr_r = rock, terr_g = dirt;
    const V2 dlr = terr_r - lr; V2 dlg = terr_g - lg; dlg = dlg + dlr;
    const V2 drr = terr_r - rr; V2 drg = terr_g - rg; drg = drg + drr;
    const V2 dtr = terr_r - tr; V2 dtg = terr_g - tg; dtg = dtg + dtr;
    const V2 dbr = terr_r - br; V2 dbg = terr_g - bg; dbg = dbg + dbr;
    V2 g_hdiff = ((dlg + drg + dtg + dbg) * 0.25f).v();      // x / 4 and x * 0.25 round identically (exact scaling by a power of two)
    V2 r_hdiff = ((dlr + drr + dtr + dbr) * 0.25f).v();
    g_hdiff = v2(fabsf(g_hdiff.x), fabsf(g_hdiff.y));
    r_hdiff = v2(fabsf(r_hdiff.x), fabsf(r_hdiff.y));
    const B2 mr = hg_smooth_extremum(dlr, drr, dtr, dbr, r_hdiff);
    const B2 mg = hg_smooth_extremum(dlg, drg, dtg, dbg, g_hdiff);
    if (b2_any(mr)) terr_r = v2_sel(mr, hg_div5_2(terr_r + lr + rr + tr + br), terr_r);
    if (b2_any(mg)) terr_g = v2_sel(mg, hg_div5_2(terr_g + lg + rg + tg + bg), terr_g);
    const float m = P.smooth_mul;
    rock = m * terr_r + (1.0f - m) * rock;
    dirt = m * terr_g + (1.0f - m) * dirt;
}
