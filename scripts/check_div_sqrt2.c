/* Exhaustive check (all finite non-negative floats; negatives follow by symmetry) of the division by sqrt(2)f that
 * hg_thermal_outflow (hg_cell.cuh) does with one multiply and two fmas:
 *     q = x c;  q' = fma(fma(-d, q, x), c, q),  d = 1.41421356237309504880f, c = RN(1/d) = 0.707106769084930419921875f.
 * Result: q' == x / d bit for bit for every x > 2.18499e-32 (0x0ae2e6eb); the 4.4 M mismatches all lie at or below that value,
 * where both forms return less than 2e-32 -- which the caller's maximum with mc >= 2^-40 makes irrelevant.
 * gcc -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp check_div_sqrt2.c -lm ; optional argument: stride (1 = exhaustive) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv) {
    const unsigned long long step = argc > 1 ? strtoull(argv[1], 0, 10) : 1ull;
    const float d = 1.41421356237309504880f, c = 0.707106769084930419921875f;
    if (c != 1.0f / d) { printf("c is not RN(1/d)\n"); return 1; }
    unsigned long long bad = 0, bad_big = 0; uint32_t maxbad = 0;
#pragma omp parallel for reduction(+ : bad, bad_big) reduction(max : maxbad)
    for (unsigned long long u = 0; u < 0x7f800000ull; u += step) {
        uint32_t b = (uint32_t)u;
        float x; memcpy(&x, &b, 4);
        float q = x * c, q2 = fmaf(fmaf(-d, q, x), c, q), t = x / d;
        if (memcmp(&q2, &t, 4)) { bad++; if (b > maxbad) maxbad = b; if (fabsf(q2) > 2e-32f || fabsf(t) > 2e-32f) bad_big++; }
    }
    float hi; memcpy(&hi, &maxbad, 4);
    printf("%llu mismatches, the largest input %g (0x%08x); mismatches with a result above 2e-32: %llu\n", bad, hi, maxbad, bad_big);
    return bad_big != 0 || hi > 2.2e-32f;
}
