/* hg_defined_math.h — the GLSL built-ins whose results OpenGL leaves to the
 * driver, pinned to one definition.
 *
 * GLSL 4.60 gives `atan`, `exp`, `sin` a precision of "implementation
 * defined" (atan: 4096 ULP allowed) and leaves min/max of NaN and of signed
 * zeros open.  The reference's results therefore differ between GL drivers in
 * exactly these places (SURVEY.md §8a hazards 1, 2, 8).  This header is part of
 * the boundary specification: it says what those built-ins mean for this
 * implementation, using only IEEE-754 +,-,*,/ and floor with no contraction,
 * so the same inputs give the same bits on an x86 host and on sm_100a.
 * tests/test_defined_math.py checks each against libm (<= 4 ulp).
 *
 * Used by the CUDA kernels (compiled with -fmad=false) and by the CPU oracle
 * (compiled with -ffp-contract=off).  Everything else in the oracle is an
 * independent restatement; only these primitives are shared.
 */
#ifndef HG_DEFINED_MATH_H
#define HG_DEFINED_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HG_FN __host__ __device__ __forceinline__
#else
#define HG_FN static inline
#endif

/* GLSL 4.60 §8.3: max(x,y) = y if x < y else x; min(x,y) = y if y < x else x.
 * Taken literally this fixes NaN and signed-zero behaviour:
 * min(1, NaN) = 1 (hydro_flux.glsl:125 with an empty cell), max(0, NaN) = 0. */
HG_FN float hg_max(float x, float y) { return (x < y) ? y : x; }
HG_FN float hg_min(float x, float y) { return (y < x) ? y : x; }
HG_FN float hg_clamp(float x, float lo, float hi) { return hg_min(hg_max(x, lo), hi); }
HG_FN float hg_fract(float x) { return x - floorf(x); }
HG_FN float hg_mod(float x, float y) { return x - y * floorf(x / y); }
HG_FN float hg_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
/* smoothstep with the edges as given, also when edge0 > edge1
 * (hydro_erosion.glsl:49 passes 1e-3, 5e-4). */
HG_FN float hg_smoothstep(float e0, float e1, float x) {
    float t = hg_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

/* atan(x): three-range reduction and a degree-9 odd polynomial (the classic
 * single-precision Cephes scheme), about 2 ulp. */
HG_FN float hg_atanf(float xx) {
    float x = fabsf(xx);
    /* The three ranges as ONE division: -(1/x) = (-1)/x and x = x/1 exactly (IEEE division is sign-symmetric and
     * division by one is exact), so selecting numerator and denominator gives the bits of the three-branch form --
     * one division instead of two in the warp-divergent code of the thermal outflow path. */
    const int hi = x > 2.414213562373095f, mid = x > 0.4142135623730950f;
    const float y = hi ? 1.5707963267948966f : (mid ? 0.7853981633974483f : 0.0f);
    const float num = hi ? -1.0f : (mid ? x - 1.0f : x);
    const float den = hi ? x : (mid ? x + 1.0f : 1.0f);
    x = num / den;
    float z = x * x;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
               - 3.33329491539e-1f) * z * x + x;
    const float r = y + p;
    return (xx < 0.0f) ? -r : r;
}

/* exp(x) for |x| < 80: x = n ln2 + r, degree-5 polynomial on r, scale by 2^n. */
HG_FN float hg_expf(float x) {
    float n = floorf(x * 1.44269504088896341f + 0.5f);
    float r = x - n * 0.693359375f;
    r = r - n * -2.12194440e-4f;
    float z = r * r;
    float p = (((((1.9875691500e-4f * r + 1.3981999507e-3f) * r + 8.3334519073e-3f) * r
                 + 4.1665795894e-2f) * r + 1.6666665459e-1f) * r + 5.0000001201e-1f) * z
              + r + 1.0f;
    int32_t e = (int32_t)n;
    if (e < -126) return 0.0f;
    if (e > 127) return INFINITY;
    uint32_t bits = (uint32_t)(e + 127) << 23;
    float s;
    memcpy(&s, &bits, sizeof(s));
    return p * s;
}

/* sin(x) for the droplet spawn hash (particle.glsl:41-44), whose arguments
 * reach ~1e6: the reduction to [-pi/4, pi/4] is done in double, the
 * polynomial in float. */
HG_FN float hg_sinf(float xf) {
    double x = (double)xf;
    double k = floor(x * 0.63661977236758134308 + 0.5);
    double rd = (x - k * 1.57079632673412561417e+00) - k * 6.07710050650619224932e-11;
    float r = (float)rd;
    float z = r * r;
    /* quadrant: k mod 4, for negative k as well */
    double q4 = k - 4.0 * floor(k * 0.25);
    int q = (int)q4;
    float s = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float c = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z
               + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    float v = (q & 1) ? c : s;
    return (q & 2) ? -v : v;
}

#endif /* HG_DEFINED_MATH_H */
