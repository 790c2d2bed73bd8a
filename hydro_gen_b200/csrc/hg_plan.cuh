// hg_plan.cuh — the arithmetic of the balanced partition of the fused step, free of any
// CUDA construct so that k_plan_segments (hg_fused.cu) and the CPU tests (tests/host_emul) run the same code.
//
// A plan is n_cta work items grouped by strip, in row order inside a strip; every item reported a duration
// ("cost") in the step that just ran.  The next plan gives a strip a number of segments proportional to its
// total cost and cuts it where the cumulative cost -- taken as uniform inside an old segment -- reaches equal
// shares.  Whatever the costs, the segments of a strip tile its rows exactly, each at least min_rows tall.
#pragma once
#include "hg_fused_body.cuh"

HG_FN float hg_plan_cost(unsigned ns) { return ns < 1u ? 1.0f : (float)ns; }

// Segments per strip: proportional to cost, largest remainders first, every strip at least 1 and at most `cap`;
// the counts sum to n_cta (given nstrips <= n_cta <= nstrips * cap).  Three steps; the first two are independent per
// strip (one device thread each), the last one is a serial repair that the clamps rarely make necessary.
HG_FN void hg_plan_share(int n_cta, float cost, float total, int cap, int* n_out, float* frac_out) {
    const float share = (float)n_cta * cost / total;
    int n = (int)share;                             // share >= 0: truncation is floor
    *frac_out = share - (float)n;
    *n_out = n < 1 ? 1 : (n > cap ? cap : n);
}
// strip k takes one of the `left` left-over segments if fewer than `left` strips have a larger remainder
HG_FN int hg_plan_bonus(int k, int nstrips, const float* frac, int left, int cap, int n_k) {
    if (left <= 0 || n_k >= cap) return 0;
    int rank = 0;
    for (int j = 0; j < nstrips; j++) rank += (frac[j] > frac[k] || (frac[j] == frac[k] && j < k)) ? 1 : 0;
    return rank < left ? 1 : 0;
}
HG_FN void hg_plan_repair(int n_cta, int nstrips, const float* strip_cost, int cap, int* new_n) {
    int given = 0;
    for (int k = 0; k < nstrips; k++) given += new_n[k];
    while (given != n_cta) {        // one at a time, where the cost per segment is largest / smallest
        int best = -1; float best_v = 0.0f;
        for (int k = 0; k < nstrips; k++) {
            if (given < n_cta ? new_n[k] >= cap : new_n[k] <= 1) continue;
            const float v = strip_cost[k] / (float)(given < n_cta ? new_n[k] : new_n[k] - 1);
            if (best < 0 || (given < n_cta ? v > best_v : v < best_v)) { best = k; best_v = v; }
        }
        if (best < 0) break;
        new_n[best] += given < n_cta ? 1 : -1;
        given += given < n_cta ? 1 : -1;
    }
}
// the three steps in sequence (host; the kernel runs the first two with one thread per strip).  frac: scratch of nstrips floats
HG_FN void hg_plan_apportion(int n_cta, int nstrips, const float* strip_cost, int cap, int* new_n, float* frac) {
    float total = 0.0f;
    for (int k = 0; k < nstrips; k++) total += strip_cost[k];
    int given = 0;
    for (int k = 0; k < nstrips; k++) { hg_plan_share(n_cta, strip_cost[k], total, cap, &new_n[k], &frac[k]); given += new_n[k]; }
    const int left = n_cta - given;
    int bonus[256];
    for (int k = 0; k < nstrips; k++) bonus[k] = hg_plan_bonus(k, nstrips, frac, left, cap, new_n[k]);
    for (int k = 0; k < nstrips; k++) new_n[k] += bonus[k];
    hg_plan_repair(n_cta, nstrips, strip_cost, cap, new_n);
}

// Strip s: n_old old segments old[0..n_old) with durations ns[0..n_old) -> n new segments out[0..n) covering rows
// [row0, row0 + rows) with equal forecast cost.
HG_FN void hg_plan_cut_strip(int s, int n, const HgPlanItem* old, const unsigned* ns, int n_old, int row0, int rows, int min_rows, HgPlanItem* out) {
    float total = 0.0f;
    for (int k = 0; k < n_old; k++) total += hg_plan_cost(ns[k]);
    const float target = total / (float)n;
    int k = 0;                                   // old segment being consumed
    float before = 0.0f;                         // cost of the old segments before k
    int y_prev = row0;
    for (int m = 1; m <= n; m++) {
        int y;
        if (m == n) {
            y = row0 + rows;
        } else {
            const float want = target * (float)m;
            while (k < n_old - 1 && before + hg_plan_cost(ns[k]) < want) { before += hg_plan_cost(ns[k]); k++; }
            float f = (want - before) / hg_plan_cost(ns[k]);
            f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
            y = old[k].gy0 + (int)(f * (float)(old[k].gy1 - old[k].gy0) + 0.5f);
            if (y < y_prev + min_rows) y = y_prev + min_rows;                                 // every segment at least min_rows tall,
            if (y > row0 + rows - (n - m) * min_rows) y = row0 + rows - (n - m) * min_rows;   // also the ones still to come
        }
        HgPlanItem it; it.strip = s; it.gy0 = y_prev; it.gy1 = y; it.pad = 0;
        out[m - 1] = it;
        y_prev = y;
    }
}
