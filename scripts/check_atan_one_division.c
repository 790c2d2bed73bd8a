/* Exhaustive check (all 2^32 bit patterns) that hg_atanf with its range reduction as ONE division (include/hg_defined_math.h) returns
 * the bits of the three-branch form it replaced.   gcc -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp check_atan_one_division.c -lm (from scripts/) */
#include "../include/hg_defined_math.h"
#include <stdio.h>
static inline float old_atanf(float xx) {
    float x = fabsf(xx);
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    else { y = 0.0f; }
    float z = x * x;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x;
    y = y + p;
    return (xx < 0.0f) ? -y : y;
}
#include <stdlib.h>
int main(int argc, char** argv) {
    const unsigned long long step = argc > 1 ? strtoull(argv[1], 0, 10) : 1ull;      /* 1 = exhaustive; a prime > 1 = a sample (the CPU test suite) */
    unsigned long long bad = 0;
#pragma omp parallel for reduction(+ : bad)
    for (unsigned long long u = 0; u <= 0xffffffffull; u += step) {
        uint32_t b = (uint32_t)u; float x; memcpy(&x, &b, 4);
        float a = old_atanf(x), c = hg_atanf(x);
        if (memcmp(&a, &c, 4)) { if (!(a != a && c != c)) bad++; }
    }
    printf("old vs one-division hg_atanf over every %llu-th of the 2^32 bit patterns: %llu mismatches (NaN results compared as NaN)\n", step, bad);
    return bad != 0;
}
