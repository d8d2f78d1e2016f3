// vecmath.hpp — 3-vector / 3x3 helpers shared by host (g++) and device (nvcc).
//
// Every function spells out its float operations in the order the reference's
// math library evaluates them (LIB/math/vec3.inl, mat3.inl, math.inl;
// LIB = path-tracer-core/path_tracer_lib/path_tracer), because closest-hit ids
// are only bit-exact when each intermediate rounds identically.  Host code is
// built with -ffp-contract=off and device code with --fmad=false, so none of
// the a*b+c below is ever fused.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PTB_HD __host__ __device__ __forceinline__
#else
#define PTB_HD inline
#endif

namespace ptb {

constexpr float kEpsilon = 0.0001f; // math::epsilon, LIB/math/math.hpp:16

struct V3 {
    float x, y, z;
};

// column-major like math::mat3: x, y, z are the columns (mat3.inl:13-29)
struct M3 {
    V3 x, y, z;
};

struct Xform { // scene::transform: origin + basis (transform.hpp:14-15)
    V3 origin;
    M3 basis;
};

PTB_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
PTB_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PTB_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PTB_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
PTB_HD V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
PTB_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
PTB_HD V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
PTB_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
PTB_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }

// math::min / math::max are "b < a ? b : a" / "b > a ? b : a" (math.inl:169-187):
// a NaN first argument is returned, a NaN second argument is dropped.
PTB_HD float rmin(float a, float b) { return b < a ? b : a; }
PTB_HD float rmax(float a, float b) { return b > a ? b : a; }
PTB_HD float rclamp(float x, float lo, float hi) { return rmin(rmax(x, lo), hi); } // math.inl:154-157
PTB_HD float rlerp(float a, float b, float w) { return a + (b - a) * w; }           // math.inl:164-167
PTB_HD V3 rlerp(V3 a, V3 b, float w) { return V3{rlerp(a.x, b.x, w), rlerp(a.y, b.y, w), rlerp(a.z, b.z, w)}; }
PTB_HD V3 rlerp(V3 a, V3 b, V3 w) { return V3{rlerp(a.x, b.x, w.x), rlerp(a.y, b.y, w.y), rlerp(a.z, b.z, w.z)}; }

PTB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // vec3.inl:235-238
PTB_HD V3 cross(V3 l, V3 r) {                                               // vec3.inl:221-228
    return V3{(l.y * r.z) - (l.z * r.y), (l.z * r.x) - (l.x * r.z), (l.x * r.y) - (l.y * r.x)};
}
PTB_HD float length(V3 v) { return sqrtf(dot(v, v)); }              // vec3.inl:245-248
PTB_HD V3 normalize(V3 v) { return v * (1 / length(v)); }           // vec3.inl:250-253
PTB_HD V3 reflect(V3 i, V3 n) { return i - 2 * dot(n, i) * n; }      // core/utils.hpp:38-40

// mat3 * vec3 goes through transpose + row dots (mat3.inl:219-224)
PTB_HD V3 mul(const M3& m, V3 v) {
    return V3{dot(V3{m.x.x, m.y.x, m.z.x}, v), dot(V3{m.x.y, m.y.y, m.z.y}, v), dot(V3{m.x.z, m.y.z, m.z.z}, v)};
}

// mat3 * mat3 builds each result column as x*r.x + y*r.y + z*r.z (mat3.inl:144-152)
PTB_HD M3 mul(const M3& a, const M3& b) {
    return M3{a.x * b.x.x + a.y * b.x.y + a.z * b.x.z, a.x * b.y.x + a.y * b.y.y + a.z * b.y.z,
              a.x * b.z.x + a.y * b.z.y + a.z * b.z.z};
}

PTB_HD M3 transpose(const M3& m) { // mat3.inl:323-330
    return M3{V3{m.x.x, m.y.x, m.z.x}, V3{m.x.y, m.y.y, m.z.y}, V3{m.x.z, m.y.z, m.z.z}};
}

// adjugate times (1/det), mat3.inl:245-263
PTB_HD M3 inverse(const M3& m) {
    float det1 = +(m.y.y * m.z.z - m.z.y * m.y.z);
    float det2 = -(m.x.y * m.z.z - m.z.y * m.x.z);
    float det3 = +(m.x.y * m.y.z - m.y.y * m.x.z);
    float det = m.x.x * det1 + m.y.x * det2 + m.z.x * det3;
    float r = 1 / det;
    M3 a{V3{det1, det2, det3},
         V3{-(m.y.x * m.z.z - m.z.x * m.y.z), +(m.x.x * m.z.z - m.z.x * m.x.z), -(m.x.x * m.y.z - m.y.x * m.x.z)},
         V3{+(m.y.x * m.z.y - m.z.x * m.y.y), -(m.x.x * m.z.y - m.z.x * m.x.y), +(m.x.x * m.y.y - m.y.x * m.x.y)}};
    return M3{a.x * r, a.y * r, a.z * r};
}

PTB_HD V3 apply(const Xform& t, V3 v) { return mul(t.basis, v) + t.origin; } // transform.cpp:117-119
PTB_HD Xform inverse(const Xform& t) {                                        // transform.cpp:33-36
    M3 b = inverse(t.basis);
    return Xform{mul(b, -t.origin), b};
}
PTB_HD Xform compose(const Xform& a, const Xform& b) { // transform.cpp:110-115
    return Xform{mul(a.basis, b.origin) + a.origin, mul(a.basis, b.basis)};
}

} // namespace ptb
