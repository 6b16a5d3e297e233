// oracle_math.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md): scalar fp32 helpers of the CPU oracle.
// Nothing under fermat_b200/ may include this file.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace oracle {

struct vec3
{
	float x, y, z;
	vec3() {}
	explicit vec3(float a) : x(a), y(a), z(a) {}
	vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	float operator[](int i) const { return (&x)[i]; }
};
struct vec2 { float x, y; vec2() {} vec2(float a, float b) : x(a), y(b) {} };

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator-(float s, vec3 a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3  cross(vec3 a, vec3 b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline float square_length(vec3 a) { return dot(a, a); }
// cugar::normalize (contrib/cugar/linalg/vector_inl.h:345-349)
inline vec3  normalize(vec3 a) { const float l = length(a); return l > 0.0f ? a / l : a; }
inline float max_comp(vec3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
inline float average(vec3 a) { return (a.x + a.y + a.z) / 3.0f; }
inline bool  finite3(vec3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }
inline vec3  lerp(vec3 a, vec3 b, float u) { return a * (1.0f - u) + b * u; }   // vector_inl.h:723
inline float saturate(float x) { return fmaxf(fminf(x, 1.0f), 0.0f); }
inline float sqr(float x) { return x * x; }
inline float mod1(float x, float m) { return x > 0.0f ? fmodf(x, m) : m - fmodf(-x, m); }  // cugar::mod, numbers.h:606

// sin/cos. The reference calls the platform's sinf/cosf (CUDA's on the device), whose last-place rounding is not
// part of its specification. Mode 0 = this host's libm (used when the oracle is pinned against the reference's own
// code compiled on this host); mode 1 (default) = a fixed sequence of IEEE fp32 operations (Cody-Waite reduction by
// pi/2 in three parts + the classic degree-7/8 minimax polynomials on [-pi/4, pi/4], ~1 ulp) that the CUDA kernels
// execute identically, so that sampled directions — and with them whole paths — agree bit for bit between the two.
inline int& trig_mode() { static int m = 1; return m; }
inline void det_sincosf(float x, float* s, float* c)
{
	const float kf = rintf(x * 0.636619772f);
	const int k = (int)kf;
	float r = x - kf * 1.5703125f;
	r = r - kf * 4.837512969970703125e-4f;
	r = r - kf * 7.54978995489188216e-8f;
	const float z = r * r;
	const float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
	const float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
	switch (k & 3)
	{
	case 0: *s = ps; *c = pc; break;
	case 1: *s = pc; *c = -ps; break;
	case 2: *s = -ps; *c = -pc; break;
	default: *s = -pc; *c = ps; break;
	}
}
inline void o_sincosf(float x, float* s, float* c)
{
	if (trig_mode() == 0) { *s = sinf(x); *c = cosf(x); }
	else det_sincosf(x, s, c);
}

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// IEEE half <-> float (round to nearest even), independent restatement
inline uint16_t f2h(float f)
{
	const uint32_t x = f2u(f), sign = (x >> 16) & 0x8000u;
	uint32_t a = x & 0x7FFFFFFFu;
	if (a > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);
	if (a >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);         // >= 65536 -> inf (65520..65536 handled by rounding below)
	if (a < 0x38800000u)                                               // subnormal half or zero
	{
		if (a < 0x33000000u) return (uint16_t)sign;
		const int shift = 126 - (int)(a >> 23);                       // 14..24
		uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
		const uint32_t lsb = 1u << shift, half_ = lsb >> 1, rem = m & (lsb - 1);
		m >>= shift;
		if (rem > half_ || (rem == half_ && (m & 1u))) m++;
		return (uint16_t)(sign | m);
	}
	uint32_t m = a & 0x7FFFFFu, e = (a >> 23) - 112;
	uint32_t h = (e << 10) | (m >> 13);
	const uint32_t rem = m & 0x1FFFu;
	if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;          // may carry into the exponent / inf: correct
	return (uint16_t)(sign | h);
}
inline float h2f(uint16_t h)
{
	const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3FFu;
	if (e == 0) { const float v = ldexpf((float)m, -24); return sign ? -v : v; }
	if (e == 31) return u2f(sign | 0x7F800000u | (m << 13));
	return u2f(sign | ((e + 112u) << 23) | (m << 13));
}

} // namespace oracle
