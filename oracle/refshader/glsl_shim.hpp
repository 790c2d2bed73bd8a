// glsl_shim.hpp — just enough of the GLSL 4.60 vocabulary, in C++, to compile the REFERENCE'S OWN
// compute shaders (/root/reference/glsl/*.glsl, read where they lie by oracle/refshader/build_ref.py)
// for the CPU.  TEST INFRASTRUCTURE: the result (oracle/_ref/libhg_refshaders.so) is what the oracle
// restatement (oracle/hg_oracle.c) is validated against; nothing in the product loads it.
//
// What is taken from the reference: every expression, its association order, every branch, constant,
// texture fetch and store -- the shader text itself, compiled by g++ -ffp-contract=off.
// What this header has to define (GLSL leaves it to the implementation): the built-ins.
// +,-,*,/ and sqrt are IEEE single operations; min/max/clamp/mix/fract/mod/smoothstep/normalize/length/
// dot/cross follow the formulas of GLSL 4.60 §8 literally, in float; atan/exp/sin come from
// include/hg_defined_math.h (the same definitions the product documents as part of its boundary);
// pow(x, 1.0) = x; texelFetch/imageLoad outside the image return 0 and imageStore outside is dropped
// (the robust-buffer-access behaviour the shaders rely on).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../../include/hg_defined_math.h"

namespace glsl {

typedef unsigned int uint;
struct vec2; struct vec3; struct vec4; struct ivec2; struct uvec2; struct uvec3;
struct IXY { int d[2]; operator ivec2() const; };      // ivec2 .xy

// ---- swizzles: views into the parent's storage (N = parent size) ----
template <int N, int A, int B> struct Sw2 {
    float d[N];
    operator vec2() const;
    Sw2& operator=(const vec2& v);
    Sw2& operator=(const Sw2& v) { float a = v.d[A], b = v.d[B]; d[A] = a; d[B] = b; return *this; }
    Sw2& operator+=(const vec2& v); Sw2& operator-=(const vec2& v); Sw2& operator*=(const vec2& v);
    Sw2& operator*=(float s) { d[A] *= s; d[B] *= s; return *this; }
    Sw2& operator/=(float s) { d[A] /= s; d[B] /= s; return *this; }
};
template <int N, int A, int B, int C> struct Sw3 {
    float d[N];
    operator vec3() const;
    Sw3& operator=(const vec3& v);
};
template <int N, int A, int B, int C, int D> struct Sw4 {
    float d[N];
    operator vec4() const;
};

struct alignas(8) vec2 {
    union {
        struct { float x, y; }; struct { float r, g; }; float d[2];
        Sw2<2, 0, 1> xy, rg; Sw2<2, 1, 0> yx; Sw2<2, 0, 0> xx; Sw2<2, 1, 1> yy;
        Sw4<2, 0, 1, 0, 1> xyxy;
    };
    vec2() : x(0), y(0) {}
    vec2(float a) : x(a), y(a) {}
    vec2(float a, float b) : x(a), y(b) {}
    vec2(const vec2& o) : x(o.x), y(o.y) {}
    vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
    vec2(const ivec2& v); vec2(const uvec2& v); vec2(const struct IXY& v);
    float& operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
    vec2& operator+=(const vec2& o) { x += o.x; y += o.y; return *this; }
    vec2& operator-=(const vec2& o) { x -= o.x; y -= o.y; return *this; }
    vec2& operator*=(const vec2& o) { x *= o.x; y *= o.y; return *this; }
    vec2& operator*=(float s) { x *= s; y *= s; return *this; }
    vec2& operator/=(float s) { x /= s; y /= s; return *this; }
};
struct vec3 {
    union {
        struct { float x, y, z; }; struct { float r, g, b; }; float d[3];
        Sw2<3, 0, 1> xy, rg; Sw2<3, 1, 2> yz; Sw2<3, 0, 2> xz;
        Sw3<3, 0, 1, 2> xyz, rgb; Sw3<3, 2, 0, 1> zxy; Sw3<3, 1, 2, 0> yzx;
    };
    vec3() : x(0), y(0), z(0) {}
    vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
    vec3(float a, const vec2& b) : x(a), y(b.x), z(b.y) {}
    vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
    float& operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};
struct vec4 {
    union {
        struct { float x, y, z, w; }; struct { float r, g, b, a; }; float d[4];
        Sw2<4, 0, 1> xy, rg; Sw2<4, 2, 3> zw, ba; Sw2<4, 1, 2> yz; Sw2<4, 0, 2> xz; Sw2<4, 1, 3> yw;
        Sw2<4, 0, 0> xx; Sw2<4, 1, 1> yy;
        Sw3<4, 0, 1, 2> xyz, rgb; Sw3<4, 3, 3, 3> www;
        Sw4<4, 0, 0, 2, 2> xxzz; Sw4<4, 0, 1, 0, 1> xyxy;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a) : x(a), y(a), z(a), w(a) {}
    vec4(float a, float b, float c, float e) : x(a), y(b), z(c), w(e) {}
    vec4(const vec2& a, const vec2& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    vec4(const vec3& a, float e) : x(a.x), y(a.y), z(a.z), w(e) {}
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    float& operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
    vec4& operator+=(const vec4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    vec4& operator-=(const vec4& o) { x -= o.x; y -= o.y; z -= o.z; w -= o.w; return *this; }
    vec4& operator*=(const vec4& o) { x *= o.x; y *= o.y; z *= o.z; w *= o.w; return *this; }
    vec4& operator*=(float s) { x *= s; y *= s; z *= s; w *= s; return *this; }
};
struct alignas(8) ivec2 {
    union { struct { int x, y; }; IXY xy; };
    ivec2() : x(0), y(0) {}
    ivec2(int a) : x(a), y(a) {}
    ivec2(int a, int b) : x(a), y(b) {}
    ivec2(const ivec2& o) : x(o.x), y(o.y) {}
    ivec2& operator=(const ivec2& o) { x = o.x; y = o.y; return *this; }
    explicit ivec2(const vec2& v) : x((int)v.x), y((int)v.y) {}     // float -> int conversion truncates (GLSL 4.60 §5.4.1)
    explicit ivec2(const uvec2& v);
    template <int N, int A, int B> explicit ivec2(const Sw2<N, A, B>& s) : x((int)s.d[A]), y((int)s.d[B]) {}
};
inline IXY::operator ivec2() const { return ivec2(d[0], d[1]); }
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct uvec3 {
    union { struct { uint x, y, z; }; struct { uint d[3]; }; };
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    struct XY { uint d[3]; operator uvec2() const { return uvec2(d[0], d[1]); } };
};
inline ivec2::ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}
inline vec2::vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
inline vec2::vec2(const IXY& v) : x((float)v.d[0]), y((float)v.d[1]) {}
inline vec2::vec2(const uvec2& v) : x((float)v.x), y((float)v.y) {}

template <int N, int A, int B> Sw2<N, A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int N, int A, int B> Sw2<N, A, B>& Sw2<N, A, B>::operator=(const vec2& v) { d[A] = v.x; d[B] = v.y; return *this; }
template <int N, int A, int B> Sw2<N, A, B>& Sw2<N, A, B>::operator+=(const vec2& v) { d[A] += v.x; d[B] += v.y; return *this; }
template <int N, int A, int B> Sw2<N, A, B>& Sw2<N, A, B>::operator-=(const vec2& v) { d[A] -= v.x; d[B] -= v.y; return *this; }
template <int N, int A, int B> Sw2<N, A, B>& Sw2<N, A, B>::operator*=(const vec2& v) { d[A] *= v.x; d[B] *= v.y; return *this; }
template <int N, int A, int B, int C> Sw3<N, A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int N, int A, int B, int C> Sw3<N, A, B, C>& Sw3<N, A, B, C>::operator=(const vec3& v) { d[A] = v.x; d[B] = v.y; d[C] = v.z; return *this; }
template <int N, int A, int B, int C, int D> Sw4<N, A, B, C, D>::operator vec4() const { return vec4(d[A], d[B], d[C], d[D]); }

// GLSL int arithmetic wraps modulo 2^32 (GLSL 4.60 §4.1.3): done in uint32 here, signed overflow is undefined in C++
#define GLSL_IOP(OP)                                                                                                   \
    inline ivec2 operator OP(const ivec2& a, const ivec2& b) { return ivec2((int)((uint)a.x OP (uint)b.x), (int)((uint)a.y OP (uint)b.y)); } \
    inline ivec2 operator OP(const ivec2& a, int b) { return ivec2((int)((uint)a.x OP (uint)b), (int)((uint)a.y OP (uint)b)); }               \
    inline ivec2 operator OP(int a, const ivec2& b) { return ivec2((int)((uint)a OP (uint)b.x), (int)((uint)a OP (uint)b.y)); }
GLSL_IOP(+) GLSL_IOP(-) GLSL_IOP(*) GLSL_IOP(^) GLSL_IOP(&)
inline ivec2 operator<<(const ivec2& a, int s) { return ivec2((int)((uint)a.x << s), (int)((uint)a.y << s)); }
inline bool operator==(const ivec2& a, const ivec2& b) { return a.x == b.x && a.y == b.y; }

// component-wise arithmetic; scalars broadcast.  S is float, or int/uint/double literals converted to float.
#define GLSL_VEC_OPS(V, ...)                                                                                 \
    inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] + b.d[i]; return r; } \
    inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] - b.d[i]; return r; } \
    inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] * b.d[i]; return r; } \
    inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] / b.d[i]; return r; } \
    inline V operator+(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] + b; return r; }      \
    inline V operator-(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] - b; return r; }      \
    inline V operator*(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] * b; return r; }      \
    inline V operator/(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a.d[i] / b; return r; }      \
    inline V operator+(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a + b.d[i]; return r; }      \
    inline V operator-(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a - b.d[i]; return r; }      \
    inline V operator*(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a * b.d[i]; return r; }      \
    inline V operator/(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = a / b.d[i]; return r; }      \
    inline V operator-(const V& a) { V r; for (int i = 0; i < __VA_ARGS__; i++) r.d[i] = -a.d[i]; return r; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)

// ---- built-ins (GLSL 4.60 §8.1-8.5), scalar ----
inline float max(float x, float y) { return hg_max(x, y); }      // y if x < y else x
inline float min(float x, float y) { return hg_min(x, y); }      // y if y < x else x
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float fract(float x) { return x - floorf(x); }
inline float floor(float x) { return floorf(x); }
inline float abs(float x) { return fabsf(x); }
inline float sqrt(float x) { return sqrtf(x); }
inline float mod(float x, float y) { return x - y * floorf(x / y); }
inline float length(float x) { return fabsf(x); }                // sqrt(x*x) for a scalar
inline float atan(float x) { return hg_atanf(x); }
inline float exp(float x) { return hg_expf(x); }
inline float sin(float x) { return hg_sinf(x); }
// pow is implementation-defined in GLSL; the exponents the shaders use are literal 1, 2 and 3: repeated products
inline float pow(float x, float y) { return y == 1.0f ? x : y == 2.0f ? x * x : y == 3.0f ? x * x * x : powf(x, y); }
template <class B> inline float pow(float x, B y) { return pow(x, (float)y); }
inline float smoothstep(float e0, float e1, float x) { return hg_smoothstep(e0, e1, x); }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
// vectors
#define GLSL_MAP1(FN, V, N) inline V FN(const V& a) { V r; for (int i = 0; i < N; i++) r.d[i] = FN(a.d[i]); return r; }
#define GLSL_MAP1_ALL(FN) GLSL_MAP1(FN, vec2, 2) GLSL_MAP1(FN, vec3, 3) GLSL_MAP1(FN, vec4, 4)
GLSL_MAP1_ALL(fract) GLSL_MAP1_ALL(floor) GLSL_MAP1_ALL(abs) GLSL_MAP1_ALL(sqrt)
#define GLSL_MAP2S(FN, V, N) inline V FN(const V& a, float b) { V r; for (int i = 0; i < N; i++) r.d[i] = FN(a.d[i], b); return r; } \
                             inline V FN(const V& a, const V& b) { V r; for (int i = 0; i < N; i++) r.d[i] = FN(a.d[i], b.d[i]); return r; }
#define GLSL_MAP2S_ALL(FN) GLSL_MAP2S(FN, vec2, 2) GLSL_MAP2S(FN, vec3, 3) GLSL_MAP2S(FN, vec4, 4)
GLSL_MAP2S_ALL(max) GLSL_MAP2S_ALL(min) GLSL_MAP2S_ALL(mod)
#define GLSL_MIX(V, N) inline V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < N; i++) r.d[i] = mix(a.d[i], b.d[i], t); return r; }
GLSL_MIX(vec2, 2) GLSL_MIX(vec3, 3) GLSL_MIX(vec4, 4)
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec2& a) { return sqrtf(a.x * a.x + a.y * a.y); }
inline float length(const vec3& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);      // GLSL 4.60 §8.5
}
inline vec3 normalize(const vec3& a) { float inv = 1.0f / length(a); return vec3(a.x * inv, a.y * inv, a.z * inv); }
// int / uint / double arguments where GLSL converts implicitly to float
template <class A, class B> inline float max(A a, B b) { return max((float)a, (float)b); }
template <class A, class B> inline float min(A a, B b) { return min((float)a, (float)b); }
template <class A, class B, class C> inline float clamp(A x, B lo, C hi) { return clamp((float)x, (float)lo, (float)hi); }
template <class B> inline float mix(float x, B y, float a) { return mix(x, (float)y, a); }

// ---- images and textures: RGBA32F, [y][x][4]; the same storage class serves sampler2D and image2D ----
struct Image {
    float* p; int w, h;
    Image() : p(nullptr), w(0), h(0) {}
    bool inside(const ivec2& q) const { return p && q.x >= 0 && q.y >= 0 && q.x < w && q.y < h; }
};
typedef Image sampler2D;
typedef Image image2D;
inline vec4 fetch_(const Image& im, const ivec2& q) {
    if (!im.inside(q)) return vec4(0.0f);
    const float* t = im.p + ((size_t)q.y * im.w + q.x) * 4;
    return vec4(t[0], t[1], t[2], t[3]);
}
inline vec4 texelFetch(const Image& im, const ivec2& q, int) { return fetch_(im, q); }
inline vec4 imageLoad(const Image& im, const ivec2& q) { return fetch_(im, q); }
inline void imageStore(const Image& im, const ivec2& q, const vec4& v) {
    if (!im.inside(q)) return;
    float* t = im.p + ((size_t)q.y * im.w + q.x) * 4;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
}
// r32ui image (the droplet lock map) and its atomics; invocations run one after the other here, so a lock is
// always free when it is tried and the spin loop of particle_erosion.glsl:89-99 takes it at once
struct uimage2D {
    uint* p; int w, h;
    uimage2D() : p(nullptr), w(0), h(0) {}
    uint* at(const ivec2& q) const { return (p && q.x >= 0 && q.y >= 0 && q.x < w && q.y < h) ? p + (size_t)q.y * w + q.x : nullptr; }
};
inline uint imageAtomicCompSwap(const uimage2D& im, const ivec2& q, uint compare, uint data) {
    uint* t = im.at(q);
    if (!t) return 0u;                 // out of bounds: the load returns 0 and nothing is stored
    uint old = *t;
    if (old == compare) *t = data;
    return old;
}
inline uint imageAtomicExchange(const uimage2D& im, const ivec2& q, uint data) {
    uint* t = im.at(q);
    if (!t) return 0u;
    uint old = *t; *t = data; return old;
}
inline void memoryBarrierImage() {}
inline void barrier() {}
inline vec2 normalize(const vec2& a) { float inv = 1.0f / length(a); return vec2(a.x * inv, a.y * inv); }
// GLSL bool occupies 4 bytes in a buffer block (std430) and as a uniform
struct gbool {
    uint v;
    gbool() : v(0) {}
    gbool(bool b) : v(b ? 1u : 0u) {}
    operator bool() const { return v != 0; }
};
inline ivec2 imageSize(const Image& im) { return ivec2(im.w, im.h); }
inline ivec2 textureSize(const Image& im, int) { return ivec2(im.w, im.h); }

// ---- compute built-in variables (set by the dispatcher; one invocation at a time) ----
struct U3 { uint x, y, z; uvec2 xy; };
extern thread_local U3 gl_GlobalInvocationID, gl_NumWorkGroups, gl_WorkGroupSize;

}  // namespace glsl
