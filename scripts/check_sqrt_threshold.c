/* Exhaustive check (all non-negative floats) that sqrtf(s) < 1e-12f  <=>  s < T, T = 0x179abe14: the momentum cut-off of
 * smoothing.glsl:93-95 without the square root (hg_cell.cuh: hg_smooth_momentum).   gcc -O2 -ffp-contract=off check_sqrt_threshold.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
int main(int argc, char** argv) {
    const uint32_t step = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1u;      /* 1 = exhaustive */
    const float c = 1e-12f;
    uint32_t lo = 0, hi = 0x7f800000u;   // smallest bits with sqrtf(x) >= c
    while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; float x; memcpy(&x, &mid, 4); if (sqrtf(x) >= c) hi = mid; else lo = mid + 1; }
    float T; memcpy(&T, &lo, 4);
    printf("T bits 0x%08x = %.9g ; sqrtf(T) = %.9g, sqrtf(prev) = %.9g, c = %.9g\n", lo, T, sqrtf(T), sqrtf(nextafterf(T, 0)), c);
    // exhaustive check over all non-negative floats
    unsigned long long bad = 0;
    for (uint64_t b64 = 0; b64 < 0x7f800000ull; b64 += step) { uint32_t b = (uint32_t)b64; float x; memcpy(&x, &b, 4); if ((sqrtf(x) < c) != (x < T)) bad++; }
    printf("mismatches: %llu\n", bad);
    return 0;
}
