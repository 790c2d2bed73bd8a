/* Exhaustive proof (all 2^31 non-negative float bit patterns below +inf; negatives follow by symmetry)
 * that division by the constants the erosion step uses can be done with one multiply and two FMAs:
 *     q = x * c;  r = fma(-d, q, x);  q' = fma(r, c, q)   ==   x / d   (round-to-nearest, bit for bit)
 * with c = RN(1/d).  hg_cell.cuh relies on it for d = 5 (smoothing.glsl:63,69) and hg_noise.cuh for
 * d = 289 (simplex_noise.glsl:320, 385-386).  It does NOT hold for d = sqrt(2)f (4.4 M mismatches), which
 * therefore keeps the generic division (thermal_erosion.glsl:73).   gcc -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp check_div_const.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static int check(float d) {
    const float c = 1.0f / d;
    unsigned long long bad = 0;
#pragma omp parallel for reduction(+ : bad)
    for (unsigned long long u = 0; u < 0x7f800000ull; u++) {
        uint32_t b = (uint32_t)u;
        float x; memcpy(&x, &b, 4);
        float q = x * c, r = fmaf(-d, q, x), q2 = fmaf(r, c, q), t = x / d;
        if (memcmp(&q2, &t, 4)) bad++;
    }
    printf("d = %.9g (c = %.9g): %llu mismatches over all finite non-negative floats\n", d, c, bad);
    return bad != 0;
}
int main(void) { return check(5.0f) | check(289.0f); }
