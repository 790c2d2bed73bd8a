// hg_v2.cuh — two fp32 lanes per thread.
//
// The fused step is bound by instruction issue, not by HBM.  sm_100 has packed fp32
// arithmetic: add/mul/fma.rn.f32x2 (SASS FADD2 / FMUL2 / FFMA2) work on an aligned register pair and take one
// issue slot for two IEEE-754 operations, with per-operand negation, a scalar-broadcast operand form and
// constant/uniform-register operands.  Each lane is rounded exactly like the scalar instruction (round to
// nearest even, denormals kept: no .ftz), so a kernel whose thread advances TWO columns in the two lanes gives the
// same bits as one that advances them in two threads.  V2 is that pair; B2 a per-lane predicate.
//
// The host build (tests/host_emul) runs the same code with two scalar operations per V2 operation.
#pragma once
#include "hg_cell.cuh"

struct alignas(8) V2 { float x, y; };
struct B2 { bool x, y; };

#if defined(__CUDA_ARCH__)
#define HG_V2_DEVICE 1
#else
#define HG_V2_DEVICE 0
#endif

HG_FN V2 v2(float a, float b) { V2 r; r.x = a; r.y = b; return r; }
HG_FN V2 v2s(float a) { V2 r; r.x = a; r.y = a; return r; }

// A PRODUCT of a packed multiplication has its own type, V2P.  ptxas 12.9 contracts mul.rn.f32x2 followed by
// add.rn.f32x2 into one FFMA2 -- although both carry an explicit rounding modifier, which by the PTX rules forbids
// contraction, and although the scalar mul.rn.f32 / add.rn.f32 pair is left alone (-fmad=false changes nothing: the
// fusion happens inside ptxas, also for inline asm).  A fused multiply-add rounds once where the shaders round twice,
// so every addition or subtraction that takes a product as an operand is issued as two scalar FADDs here, which
// ptxas does not merge with a packed multiply.  The type makes the compiler find those places: a V2P only converts
// to a V2 explicitly (v()), for uses that are not additions (another multiplication, min/max, a store, an fma
// operand).
#if defined(__CUDACC__)
#define HG_MFN __host__ __device__ __forceinline__
#else
#define HG_MFN inline
#endif
struct V2P {
    float x, y;
    HG_MFN V2 v() const { V2 r; r.x = x; r.y = y; return r; }
};
#if HG_V2_DEVICE
HG_FN float2 hg_f2(V2 a) { return make_float2(a.x, a.y); }
HG_FN float2 hg_f2(V2P a) { return make_float2(a.x, a.y); }
HG_FN V2 hg_v2(float2 a) { V2 r; r.x = a.x; r.y = a.y; return r; }
HG_FN V2P hg_v2p(float2 a) { V2P r; r.x = a.x; r.y = a.y; return r; }
HG_FN V2 operator+(V2 a, V2 b) { return hg_v2(__fadd2_rn(hg_f2(a), hg_f2(b))); }
HG_FN V2 operator-(V2 a, V2 b) { return hg_v2(__fadd2_rn(hg_f2(a), make_float2(-b.x, -b.y))); }      // FADD2 with a negated operand
HG_FN V2 operator+(V2 a, float s) { return hg_v2(__fadd2_rn(hg_f2(a), make_float2(s, s))); }
HG_FN V2 operator-(V2 a, float s) { return hg_v2(__fadd2_rn(hg_f2(a), make_float2(-s, -s))); }
HG_FN V2 operator-(float s, V2 b) { return hg_v2(__fadd2_rn(make_float2(s, s), make_float2(-b.x, -b.y))); }
HG_FN V2P operator*(V2 a, V2 b) { return hg_v2p(__fmul2_rn(hg_f2(a), hg_f2(b))); }
HG_FN V2P operator*(float s, V2 b) { return hg_v2p(__fmul2_rn(make_float2(s, s), hg_f2(b))); }          // scalar-broadcast operand
HG_FN V2P operator*(V2 a, float s) { return hg_v2p(__fmul2_rn(hg_f2(a), make_float2(s, s))); }
// a * b + c with ONE rounding: only where the scalar code uses an explicit fma (exact division sequences)
HG_FN V2 v2_fma(V2 a, V2 b, V2 c) { return hg_v2(__ffma2_rn(hg_f2(a), hg_f2(b), hg_f2(c))); }
#define HG_SADD(a, b) __fadd_rn((a), (b))
#else
HG_FN V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
HG_FN V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
HG_FN V2 operator+(V2 a, float s) { return v2(a.x + s, a.y + s); }
HG_FN V2 operator-(V2 a, float s) { return v2(a.x - s, a.y - s); }
HG_FN V2 operator-(float s, V2 b) { return v2(s - b.x, s - b.y); }
HG_FN V2P operator*(V2 a, V2 b) { V2P r; r.x = a.x * b.x; r.y = a.y * b.y; return r; }
HG_FN V2P operator*(float s, V2 b) { V2P r; r.x = s * b.x; r.y = s * b.y; return r; }
HG_FN V2P operator*(V2 a, float s) { V2P r; r.x = a.x * s; r.y = a.y * s; return r; }
HG_FN V2 v2_fma(V2 a, V2 b, V2 c) { return v2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#define HG_SADD(a, b) ((a) + (b))
#endif
// products multiply on as values
HG_FN V2P operator*(V2P a, V2 b) { return a.v() * b; }
HG_FN V2P operator*(V2 a, V2P b) { return a * b.v(); }
HG_FN V2P operator*(V2P a, V2P b) { return a.v() * b.v(); }
HG_FN V2P operator*(float s, V2P b) { return s * b.v(); }
HG_FN V2P operator*(V2P a, float s) { return a.v() * s; }
// sums with a product operand: two scalar additions (see above)
HG_FN V2 operator+(V2P a, V2P b) { return v2(HG_SADD(a.x, b.x), HG_SADD(a.y, b.y)); }
HG_FN V2 operator+(V2P a, V2 b) { return v2(HG_SADD(a.x, b.x), HG_SADD(a.y, b.y)); }
HG_FN V2 operator+(V2 a, V2P b) { return v2(HG_SADD(a.x, b.x), HG_SADD(a.y, b.y)); }
HG_FN V2 operator+(V2P a, float s) { return v2(HG_SADD(a.x, s), HG_SADD(a.y, s)); }
HG_FN V2 operator-(V2P a, V2P b) { return v2(HG_SADD(a.x, -b.x), HG_SADD(a.y, -b.y)); }
HG_FN V2 operator-(V2P a, V2 b) { return v2(HG_SADD(a.x, -b.x), HG_SADD(a.y, -b.y)); }
HG_FN V2 operator-(V2 a, V2P b) { return v2(HG_SADD(a.x, -b.x), HG_SADD(a.y, -b.y)); }
HG_FN V2 operator-(V2P a, float s) { return v2(HG_SADD(a.x, -s), HG_SADD(a.y, -s)); }
HG_FN V2 operator-(float s, V2P b) { return v2(HG_SADD(s, -b.x), HG_SADD(s, -b.y)); }
HG_FN V2 v2_neg(V2 a) { return v2(-a.x, -a.y); }

// per-lane operations (no packed form exists: FMNMX, FSEL, FSETP)
HG_FN V2 v2_max_c(float c, V2 v) { return v2(hg_max_c(c, v.x), hg_max_c(c, v.y)); }
HG_FN V2 v2_min_c(float c, V2 v) { return v2(hg_min_c(c, v.x), hg_min_c(c, v.y)); }
HG_FN V2 v2_fmax(V2 a, V2 b) { return v2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
HG_FN V2 v2_sel(B2 m, V2 a, V2 b) { return v2(m.x ? a.x : b.x, m.y ? a.y : b.y); }
HG_FN V2 v2_sel(B2 m, V2 a, float b) { return v2(m.x ? a.x : b, m.y ? a.y : b); }
HG_FN V2 v2_sel(B2 m, float a, V2 b) { return v2(m.x ? a : b.x, m.y ? a : b.y); }
HG_FN B2 b2(bool a, bool b) { B2 r; r.x = a; r.y = b; return r; }
HG_FN B2 operator&&(B2 a, B2 b) { return b2(a.x && b.x, a.y && b.y); }
HG_FN B2 operator&&(B2 a, bool b) { return b2(a.x && b, a.y && b); }
HG_FN B2 operator||(B2 a, B2 b) { return b2(a.x || b.x, a.y || b.y); }
HG_FN B2 operator!(B2 a) { return b2(!a.x, !a.y); }
HG_FN bool b2_any(B2 a) { return a.x || a.y; }
HG_FN B2 v2_gt(V2 a, V2 b) { return b2(a.x > b.x, a.y > b.y); }
HG_FN B2 v2_gt(V2 a, float b) { return b2(a.x > b, a.y > b); }
HG_FN B2 v2_ge(V2 a, float b) { return b2(a.x >= b, a.y >= b); }
HG_FN B2 v2_lt(V2 a, float b) { return b2(a.x < b, a.y < b); }
