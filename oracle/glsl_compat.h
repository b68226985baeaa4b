// glsl_compat.h — enough of GLSL 4.50 as C++ to compile the REFERENCE'S shader sources verbatim (test infrastructure only).
//
// oracle/glsl_to_cpp.py turns a shader file under /root/reference/src/shaders into the body of a C++ struct (interface
// qualifiers stripped, float literals suffixed, nothing else touched); this header supplies the types and built-ins that
// body needs: vec2/3/4, ivec2/3, uvec3, mat3/4 with the swizzles the shaders use (as aliasing proxy members, the GLM
// technique), the arithmetic / geometric / exponential built-ins in IEEE float32, and sampler types whose look-ups are
// std::function hooks — the harness (oracle/glslref_harness.cpp) binds them to the oracle's texture units
// (oracle/orc_texture.cpp), because texture filtering is GL fixed function, not shader text.
// Every float operation is a single float32 operation (compile with -ffp-contract=off), as in the oracle.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>

namespace glsl {

typedef unsigned int uint;
struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3;

// ---- swizzle proxies: members of an anonymous union over the parent's P components ----
template <int P, int A, int B> struct Sw2 {
    float v[P];
    inline operator vec2() const;
    inline Sw2& operator=(const vec2& o);
    inline Sw2& operator=(const Sw2& o);
};
template <int P, int A, int B, int C> struct Sw3 {
    float v[P];
    inline operator vec3() const;
    inline Sw3& operator=(const vec3& o);
    inline Sw3& operator=(const Sw3& o);
    inline Sw3& operator+=(const vec3& o);
    inline Sw3& operator-=(const vec3& o);
    inline Sw3& operator*=(const vec3& o);
    inline Sw3& operator*=(float s);
    inline Sw3& operator/=(float s);
};
template <int P, int A, int B, int C, int D> struct Sw4 {
    float v[P];
    inline operator vec4() const;
    inline Sw4& operator=(const vec4& o);
    inline Sw4& operator=(const Sw4& o);
};
template <int P, int A, int B> struct ISw2 {
    int v[P];
    inline operator ivec2() const;
    inline operator vec2() const;
};

struct vec2 {
    union { struct { float x, y; }; struct { float r, g; }; struct { float s, t; }; float d[2]; Sw2<2, 0, 1> xy; Sw2<2, 1, 0> yx; };
    vec2() : x(0), y(0) {}
    explicit vec2(float a) : x(a), y(a) {}
    vec2(float a, float b) : x(a), y(b) {}
    vec2(const vec2& o) : x(o.x), y(o.y) {}
    vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
    float& operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
};
struct vec3 {
    union {
        struct { float x, y, z; }; struct { float r, g, b; }; float d[3];
        Sw2<3, 0, 1> xy; Sw2<3, 1, 2> yz; Sw3<3, 0, 1, 2> xyz; Sw3<3, 0, 1, 2> rgb;
    };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
    inline explicit vec3(const vec4& o);
    vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
    float& operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
};
struct vec4 {
    union {
        struct { float x, y, z, w; }; struct { float r, g, b, a; }; float d[4];
        Sw2<4, 0, 1> xy; Sw2<4, 1, 2> yz; Sw2<4, 2, 3> zw; Sw2<4, 0, 2> xz;
        Sw3<4, 0, 1, 2> xyz; Sw3<4, 0, 1, 2> rgb;
        Sw4<4, 0, 1, 3, 3> xyww; Sw4<4, 0, 1, 3, 2> xywz; Sw4<4, 0, 1, 2, 3> xyzw;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float p, float q, float s, float t) : x(p), y(q), z(s), w(t) {}
    vec4(const vec3& p, float t) : x(p.x), y(p.y), z(p.z), w(t) {}
    vec4(const vec2& p, float s, float t) : x(p.x), y(p.y), z(s), w(t) {}
    vec4(const vec2& p, const vec2& q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    float& operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
};
inline vec3::vec3(const vec4& o) : x(o.x), y(o.y), z(o.z) {}

struct ivec2 {
    union { struct { int x, y; }; int d[2]; };
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const vec2& f) : x((int)f.x), y((int)f.y) {}   // float -> int conversion truncates toward zero
    operator vec2() const { return vec2((float)x, (float)y); }
};
struct ivec3 {
    union { struct { int x, y, z; }; int d[3]; ISw2<3, 0, 1> xy; };
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    ivec3(const ivec3& o) : x(o.x), y(o.y), z(o.z) {}
    ivec3& operator=(const ivec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};
struct uvec3 {
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }

// proxy bodies
template <int P, int A, int B> Sw2<P, A, B>::operator vec2() const { return vec2(v[A], v[B]); }
template <int P, int A, int B> Sw2<P, A, B>& Sw2<P, A, B>::operator=(const vec2& o) { v[A] = o.x; v[B] = o.y; return *this; }
template <int P, int A, int B> Sw2<P, A, B>& Sw2<P, A, B>::operator=(const Sw2& o) { vec2 t = o; return *this = t; }
template <int P, int A, int B, int C> Sw3<P, A, B, C>::operator vec3() const { return vec3(v[A], v[B], v[C]); }
template <int P, int A, int B, int C> Sw3<P, A, B, C>& Sw3<P, A, B, C>::operator=(const vec3& o) { v[A] = o.x; v[B] = o.y; v[C] = o.z; return *this; }
template <int P, int A, int B, int C> Sw3<P, A, B, C>& Sw3<P, A, B, C>::operator=(const Sw3& o) { vec3 t = o; return *this = t; }
template <int P, int A, int B, int C> Sw3<P, A, B, C>& Sw3<P, A, B, C>::operator+=(const vec3& o) { v[A] += o.x; v[B] += o.y; v[C] += o.z; return *this; }
template <int P, int A, int B, int C> Sw3<P, A, B, C>& Sw3<P, A, B, C>::operator-=(const vec3& o) { v[A] -= o.x; v[B] -= o.y; v[C] -= o.z; return *this; }
template <int P, int A, int B, int C> Sw3<P, A, B, C>& Sw3<P, A, B, C>::operator*=(const vec3& o) { v[A] *= o.x; v[B] *= o.y; v[C] *= o.z; return *this; }
template <int P, int A, int B, int C> Sw3<P, A, B, C>& Sw3<P, A, B, C>::operator*=(float s) { v[A] *= s; v[B] *= s; v[C] *= s; return *this; }
template <int P, int A, int B, int C> Sw3<P, A, B, C>& Sw3<P, A, B, C>::operator/=(float s) { v[A] /= s; v[B] /= s; v[C] /= s; return *this; }
template <int P, int A, int B, int C, int D> Sw4<P, A, B, C, D>::operator vec4() const { return vec4(v[A], v[B], v[C], v[D]); }
template <int P, int A, int B, int C, int D> Sw4<P, A, B, C, D>& Sw4<P, A, B, C, D>::operator=(const vec4& o) { v[A] = o.x; v[B] = o.y; v[C] = o.z; v[D] = o.w; return *this; }
template <int P, int A, int B, int C, int D> Sw4<P, A, B, C, D>& Sw4<P, A, B, C, D>::operator=(const Sw4& o) { vec4 t = o; return *this = t; }
template <int P, int A, int B> ISw2<P, A, B>::operator ivec2() const { return ivec2(v[A], v[B]); }
template <int P, int A, int B> ISw2<P, A, B>::operator vec2() const { return vec2((float)v[A], (float)v[B]); }

// ---- component-wise arithmetic (non-template on purpose: swizzle proxies and ivec2 reach them by implicit conversion) ----
#define GLSL_VEC_OPS(V, N)                                                                                              \
    inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; } \
    inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; } \
    inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.d[i]; return r; } \
    inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b.d[i]; return r; } \
    inline V operator+(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + s; return r; }         \
    inline V operator-(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - s; return r; }         \
    inline V operator*(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }         \
    inline V operator/(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / s; return r; }         \
    inline V operator+(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s + a.d[i]; return r; }         \
    inline V operator-(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s - a.d[i]; return r; }         \
    inline V operator*(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }         \
    inline V operator/(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s / a.d[i]; return r; }         \
    inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }                     \
    inline V& operator+=(V& a, const V& b) { for (int i = 0; i < N; ++i) a.d[i] += b.d[i]; return a; }                  \
    inline V& operator-=(V& a, const V& b) { for (int i = 0; i < N; ++i) a.d[i] -= b.d[i]; return a; }                  \
    inline V& operator*=(V& a, const V& b) { for (int i = 0; i < N; ++i) a.d[i] *= b.d[i]; return a; }                  \
    inline V& operator/=(V& a, const V& b) { for (int i = 0; i < N; ++i) a.d[i] /= b.d[i]; return a; }                  \
    inline V& operator+=(V& a, float s) { for (int i = 0; i < N; ++i) a.d[i] += s; return a; }                          \
    inline V& operator-=(V& a, float s) { for (int i = 0; i < N; ++i) a.d[i] -= s; return a; }                          \
    inline V& operator*=(V& a, float s) { for (int i = 0; i < N; ++i) a.d[i] *= s; return a; }                          \
    inline V& operator/=(V& a, float s) { for (int i = 0; i < N; ++i) a.d[i] /= s; return a; }                          \
    inline bool operator==(const V& a, const V& b) { for (int i = 0; i < N; ++i) if (!(a.d[i] == b.d[i])) return false; return true; } \
    inline bool operator!=(const V& a, const V& b) { return !(a == b); }                                                \
    inline float dot(const V& a, const V& b) { float s = a.d[0] * b.d[0]; for (int i = 1; i < N; ++i) s = s + a.d[i] * b.d[i]; return s; } \
    inline float length(const V& a) { return std::sqrt(dot(a, a)); }                                                    \
    inline V normalize(const V& a) { float l = std::sqrt(dot(a, a)); V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / l; return r; } \
    inline V pow(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = std::pow(a.d[i], b.d[i]); return r; } \
    inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] < b.d[i] ? b.d[i] : a.d[i]; return r; } \
    inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = b.d[i] < a.d[i] ? b.d[i] : a.d[i]; return r; } \
    inline V abs(const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = std::fabs(a.d[i]); return r; }                 \
    inline V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * (1.0f - t) + b.d[i] * t; return r; } \
    inline V clamp(const V& a, float lo, float hi) { V r; for (int i = 0; i < N; ++i) { float m = a.d[i] < lo ? lo : a.d[i]; r.d[i] = hi < m ? hi : m; } return r; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
#undef GLSL_VEC_OPS

// scalars (GLSL 4.50 §8: max/min/clamp are defined with comparisons, NaN behaviour undefined — clamp(NaN) yields lo here
// only if the comparison chain says so; the harness never relies on it)
inline float max(float a, float b) { return a < b ? b : a; }
inline float min(float a, float b) { return b < a ? b : a; }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float abs(float a) { return std::fabs(a); }
inline float sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float inversesqrt(float a) { return 1.0f / std::sqrt(a); }
inline float sin(float a) { return std::sin(a); }
inline float cos(float a) { return std::cos(a); }
inline float asin(float a) { return std::asin(a); }
inline float acos(float a) { return std::acos(a); }
inline float atan(float y, float x) { return std::atan2(y, x); }
inline float exp2(float a) { return std::exp2(a); }
inline float log2(float a) { return std::log2(a); }
inline float floor(float a) { return std::floor(a); }
inline float smoothstep(float e0, float e1, float x) { float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline vec3 reflect(const vec3& I, const vec3& N) { return I - 2.0f * dot(N, I) * N; }

// ---- matrices (column-major, m[col][row]) ----
struct mat4;
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(float s) { c[0] = vec3(s, 0, 0); c[1] = vec3(0, s, 0); c[2] = vec3(0, 0, s); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    inline explicit mat3(const mat4& m);
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
    explicit mat4(const mat3& m) { c[0] = vec4(m.c[0], 0.0f); c[1] = vec4(m.c[1], 0.0f); c[2] = vec4(m.c[2], 0.0f); c[3] = vec4(0, 0, 0, 1); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m) { for (int i = 0; i < 3; ++i) c[i] = vec3(m.c[i]); }
inline vec3 operator*(const mat3& m, const vec3& v) {
    vec3 r;
    for (int i = 0; i < 3; ++i) r.d[i] = m.c[0].d[i] * v.x + m.c[1].d[i] * v.y + m.c[2].d[i] * v.z;
    return r;
}
inline vec4 operator*(const mat4& m, const vec4& v) {
    vec4 r;
    for (int i = 0; i < 4; ++i) r.d[i] = m.c[0].d[i] * v.x + m.c[1].d[i] * v.y + m.c[2].d[i] * v.z + m.c[3].d[i] * v.w;
    return r;
}
inline mat4 operator*(const mat4& a, const mat4& b) { mat4 r; for (int j = 0; j < 4; ++j) r.c[j] = a * b.c[j]; return r; }

// ---- samplers: fixed-function look-ups supplied by the harness ----
struct floatS { float r; operator float() const { return r; } };   // texture() on a shadow sampler: float, the shader writes ".r"
struct sampler2D {
    std::function<vec4(vec2)> sample;                 // texture(): implicit-LOD filtered look-up
    std::function<vec4(ivec2, int)> fetch;            // texelFetch()
    int levels = 1;
};
struct sampler2DRect {
    std::function<vec4(vec2)> sample;                 // texture(): unnormalised coordinates
    std::function<vec4(ivec2)> fetch;
    ivec2 size;
};
struct samplerCube { std::function<vec4(vec3, float)> sample_lod; float implicit_lod = 0.0f; };
struct sampler2DArrayShadow { std::function<float(vec4)> compare; ivec3 size; };
inline vec4 texture(const sampler2D& s, const vec2& uv) { return s.sample(uv); }
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int level) { return s.fetch(p, level); }
inline int textureQueryLevels(const sampler2D& s) { return s.levels; }
inline vec4 texture(const sampler2DRect& s, const vec2& uv) { return s.sample(uv); }
inline vec4 texelFetch(const sampler2DRect& s, const ivec2& p) { return s.fetch(p); }
inline ivec2 textureSize(const sampler2DRect& s) { return s.size; }
inline vec4 texture(const samplerCube& s, const vec3& dir) { return s.sample_lod(dir, s.implicit_lod); }
inline vec4 textureLod(const samplerCube& s, const vec3& dir, float lod) { return s.sample_lod(dir, lod); }
inline floatS texture(const sampler2DArrayShadow& s, const vec4& p) { floatS f; f.r = s.compare(p); return f; }
inline ivec3 textureSize(const sampler2DArrayShadow& s, int) { return s.size; }

}  // namespace glsl
