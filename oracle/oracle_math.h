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
