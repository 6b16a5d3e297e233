// fb_math.h — minimal fp32 vector algebra for host and device code.
#pragma once
#include "fb_types.h"
#include <math.h>
#include <string.h>

namespace fb {

struct V3
{
	float x, y, z;
	FB_HD V3() {}
	FB_HD V3(float a) : x(a), y(a), z(a) {}
	FB_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
	FB_HD V3(const float3 f) : x(f.x), y(f.y), z(f.z) {}
	FB_HD V3(const float4 f) : x(f.x), y(f.y), z(f.z) {}
	FB_HD float  operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct V2 { float x, y; FB_HD V2() {} FB_HD V2(float a, float b) : x(a), y(b) {} };

FB_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
FB_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
FB_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
FB_HD V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
FB_HD V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
FB_HD V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
FB_HD V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
FB_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
FB_HD V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
FB_HD V3& operator*=(V3& a, V3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
FB_HD V3& operator*=(V3& a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
FB_HD V3& operator/=(V3& a, float s) { a.x /= s; a.y /= s; a.z /= s; return a; }
FB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FB_HD V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
FB_HD float square_length(V3 a) { return dot(a, a); }
FB_HD float length(V3 a) { return sqrtf(dot(a, a)); }
// same rounding as cugar::normalize: component-wise division by the length, zero vectors pass through
// (reference contrib/cugar/linalg/vector_inl.h:345-349)
FB_HD V3 normalize(V3 a) { const float l = sqrtf(dot(a, a)); return l > 0.0f ? V3(a.x / l, a.y / l, a.z / l) : a; }
FB_HD float max_comp(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
FB_HD float min_comp(V3 a) { return fminf(a.x, fminf(a.y, a.z)); }
// (comparisons rather than fminf / fmaxf on the host: without -ffast-math gcc calls libm for those, and the BVH builders spend their
// time here; boxes never hold NaNs)
#ifdef __CUDA_ARCH__
FB_HD V3 vmin(V3 a, V3 b) { return V3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
FB_HD V3 vmax(V3 a, V3 b) { return V3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
#else
FB_HD V3 vmin(V3 a, V3 b) { return V3(b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z); }
FB_HD V3 vmax(V3 a, V3 b) { return V3(b.x > a.x ? b.x : a.x, b.y > a.y ? b.y : a.y, b.z > a.z ? b.z : a.z); }
#endif
FB_HD bool  is_finite(float a) { return isfinite(a); }
FB_HD bool  is_finite(V3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
FB_HD float average(V3 a) { return (a.x + a.y + a.z) / 3.0f; }

FB_HD uint32 float_as_uint(float f) { uint32 u; memcpy(&u, &f, 4); return u; }
FB_HD float  uint_as_float(uint32 u) { float f; memcpy(&f, &u, 4); return f; }

struct Bbox3
{
	V3 lo, hi;
	FB_HD Bbox3() : lo(1.0e30f), hi(-1.0e30f) {}
	FB_HD void insert(V3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
	FB_HD void insert(const Bbox3& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
	FB_HD float half_area() const { V3 d = hi - lo; return d.x * d.y + d.y * d.z + d.z * d.x; }
};

} // namespace fb
