// bsdf.cuh — device implementation of Fermat's layered BSDF for the `-pt` path:
// clearcoat o { Lambert R, Lambert T, GGX-Smith glossy R, GGX-Smith glossy T }.
//
// Behaviour (every branch, clamp and constant) follows the reference:
//   src/bsdf.h:219-243 (construction), :366-412 (f_and_p), :530-627 (sampling weights),
//   :632-792 (Fresnel / component weights), :921-1199 (sample, USE_EFFICIENT_SAMPLER_WITH_APPROXIMATE_PDFS),
//   :1202-1268 (clearcoat, radiance compression, albedo table)
//   contrib/cugar/bsdf/{ggx_smith.h, ggx_common.h, lambert.h, lambert_trans.h, refraction.h}
//   contrib/cugar/spherical/mappings_inline.h:56-127
// What is ours: the organisation. The reference evaluates four lobe objects and then combines them; here
// the shared sub-expressions (N.V, N.L, the microfacet H, D, G, G1, the clearcoat Fresnel) are computed once
// and the lobes are assembled from them, which removes most of the redundant dot products and square roots
// and keeps the live register set small enough for 8 CTAs of 128 threads per SM.
// Compiled with -fmad=false: +,-,*,/ and sqrtf are IEEE and unfused, so the results track the CPU oracle
// bit for bit except where sinf/cosf differ in the last place.
#pragma once
#include "../host/fb_math.h"

namespace fb {

#define FB_PI 3.14159265358979323846f

struct Frame            // cugar::DifferentialGeometry
{
	V3 normal_s, tangent, binormal;
	FB_D V3 to_local(V3 v) const { return V3(dot(v, tangent), dot(v, binormal), dot(v, normal_s)); }
	FB_D V3 from_local(V3 v) const { return v.x * tangent + v.y * binormal + v.z * normal_s; }
};

struct BsdfParams       // the Bsdf members that are not derivable on the fly
{
	V3 diffuse;         // Kd / pi
	V3 diffuse_trans;   // Td / pi
	V3 fresnel;         // Ks / pi
	V3 reflectivity;    // clearcoat normal-incidence reflectivity
	float roughness, inv_roughness;
	float ior, opacity, clearcoat_ior;
};

enum { B_DR = 0, B_DT = 1, B_GR = 2, B_GT = 3 };

FB_D float bsdf_saturate(float x) { return fmaxf(fminf(x, 1.0f), 0.0f); }
FB_D float clamp_inf(float p) { return (!isfinite(p) || isnan(p)) ? 1.0e8f : fmaxf(p, 0.0f); }

FB_D void bsdf_init(BsdfParams& b, V3 kd, V3 td, V3 ks, V3 kr, float roughness, float ior, float opacity)
{
	b.diffuse = kd / FB_PI;
	b.diffuse_trans = td / FB_PI;
	b.fresnel = ks / FB_PI;
	b.reflectivity = kr;
	b.roughness = fmaxf(roughness * 1.0f + 0.0f, 0.0f);
	b.inv_roughness = 1.0f / b.roughness;
	b.ior = ior; b.opacity = opacity;
	const float R0 = fminf(max_comp(kr), 0.95f);
	b.clearcoat_ior = (1 + sqrtf(R0)) / (1 - sqrtf(R0));
}

// 32^4 albedo table lookup (src/bsdf.h:1253-1268)
FB_D float glossy_reflectance(const BsdfParams& b, const float* __restrict__ table, float cos_theta)
{
	const float eta = cos_theta > 0.0f ? 1.0f / b.ior : b.ior;
	const uint32 ci = min(31u, (uint32)(fabsf(cos_theta) * 31u));
	const uint32 bi = min(31u, (uint32)(max_comp(b.fresnel) * 31u));
	const uint32 ei = min(31u, (uint32)((eta / 2.0f) * 31u));
	const uint32 ri = min(31u, (uint32)(b.roughness * 31u));
	return __ldg(table + (ei * 32768u + bi * 1024u + ri * 32u + ci));
}

FB_D V3 fresnel_schlick(float ci, float eta, V3 base)
{
	ci = bsdf_saturate(fabsf(ci));
	const float ct2 = bsdf_saturate(1.f - eta * eta * (1.f - ci * ci));
	if (ct2 < 0.0f) return V3(1.0f);
	const float c = eta > 1.0f ? sqrtf(ct2) : ci;
	const float x = 1 - c, x2 = x * x;
	const float Fc = x2 * x2 * x;
	return V3(Fc) + (1 - Fc) * base;
}
FB_D void fresnel_weights(const BsdfParams& b, float VoH, float eta, V3& r, V3& t)
{
	if (eta == 0.0f) { r = V3(0.0f); t = V3(1.0f); }
	else { r = fresnel_schlick(VoH, eta, b.fresnel); t = V3(1.0f - max_comp(r)); }
}

// clearcoat Fresnel (src/bsdf.h:1202-1232); returns false on total internal reflection
FB_D bool clearcoat_transmission(const BsdfParams& b, float cos_theta_i, V3& Fc_1, V3& Tc_1)
{
	const float R0 = fminf(max_comp(b.reflectivity), 0.95f);
	const float eta = 1.0f / b.clearcoat_ior;
	float F;
	if (eta == 1.0f) F = 0.0f;
	else
	{
		const float ct2 = 1.f - eta * eta * (1.f - cos_theta_i * cos_theta_i);
		if (ct2 < 0.0f) { Fc_1 = V3(1.0f); Tc_1 = V3(0.0f); return false; }
		const float ct = fabsf((cos_theta_i >= 0.0f ? -1.0f : 1.0f) * sqrtf(ct2));
		const float ci = fabsf(cos_theta_i);
		// fresnel_dielectric (refraction.h:50-68); eta != 1 here
		const float Rs = (ci - eta * ct) / (ci + eta * ct);
		const float Rp = (eta * ci - ct) / (eta * ci + ct);
		F = 0.5f * (Rs * Rs + Rp * Rp);
	}
	const float u = fmaxf(F - R0, 0.0f) / (1 - R0);
	Fc_1 = b.reflectivity * (1.0f - u) + V3(1.0f) * u;
	Tc_1 = V3(1.0f) - Fc_1;
	return true;
}

FB_D float compression_factor(const BsdfParams& b, float NoV, float NoL)
{
	if (b.ior && NoV * NoL < 0.0f) { const float e = NoV > 0.0f ? b.ior : 1.0f / b.ior; return e * e; }
	return 1.0f;
}

FB_D void sampling_weights(const BsdfParams& b, const float* __restrict__ table, float NoV_signed, float w[4])
{
	float r, t;
	if (b.ior == 0) { r = 0.0f; t = 1.0f; }
	else { r = glossy_reflectance(b, table, NoV_signed); t = 1.0f - r; }
	w[B_GR] = r;
	w[B_GT] = (1 - b.opacity) * t;
	w[B_DR] = b.opacity * max_comp(V3(t) * b.diffuse) * FB_PI;
	w[B_DT] = b.opacity * max_comp(V3(t) * b.diffuse_trans) * FB_PI;
}

// cugar::microfacet (ggx_common.h:50-64): H oriented along N
FB_D V3 microfacet_n(V3 V, V3 L, V3 N, float NoV, float NoL, float inv_eta)
{
	V3 H = (NoV * NoL >= 0.0f) ? V + L : V + L * inv_eta;
	if (dot(H, H) == 0.0f) return N;
	if (dot(N, H) < 0.0f) H = -H;
	return normalize(H);
}
// cugar::vndf_microfacet (ggx_common.h:66-81): H oriented along V
FB_D V3 microfacet_v(V3 V, V3 L, V3 N, float NoV, float NoL, float inv_eta)
{
	V3 H = (NoV * NoL >= 0.0f) ? V + L : V + L * inv_eta;
	if (dot(H, H) < 1.0e-12f) return N;
	if (dot(V, H) < 0.0f) H = -H;
	return normalize(H);
}

// inner_component_weights (src/bsdf.h:722-743)
FB_D void inner_component_weights(const BsdfParams& b, const float* __restrict__ table, V3 N, V3 V, V3 L, float NoV, float NoL, V3 w[4])
{
	float eta = 0.0f, VoH = 0.0f;
	if (b.ior)
	{
		eta = NoV > 0.0f ? 1.0f / b.ior : b.ior;
		const float inv_eta = NoV > 0.0f ? b.ior : 1.0f / b.ior;
		VoH = dot(V, microfacet_n(V, L, N, NoV, NoL, inv_eta));
	}
	V3 r, t;
	fresnel_weights(b, VoH, eta, r, t);
	const float dw = (1.0f - glossy_reflectance(b, table, NoV)) * (1.0f - glossy_reflectance(b, table, NoL));
	w[B_GR] = r;
	w[B_GT] = t * (1 - b.opacity);
	w[B_DR] = t * b.opacity * dw;
	w[B_DT] = w[B_DR];
}

FB_D float hvd_ggx_eval(float inv_alpha, float nh, float ht, float hb)
{
	const float x = ht * inv_alpha, y = hb * inv_alpha;
	const float aniso = x * x + y * y;
	const float f = aniso + nh * nh;
	return (1.0f / FB_PI) * inv_alpha * inv_alpha / (f * f);
}
FB_D float smith_joint_approx(float a, float NoV, float NoL)
{
	const float vis_v = NoL * (NoV * (1 - a) + a), vis_l = NoV * (NoL * (1 - a) + a);
	return 0.5f * 1.0f / (vis_v + vis_l);
}
FB_D float smith_g1v(float a, float NoV, float NoL)
{
	const float a2 = a * a;
	const float G_V = NoV + sqrtf((NoV - NoV * a2) * NoV + a2);
	return 0.5f / (G_V * NoL);
}
FB_D float transmission_factor(float VoH, float LoH, float eta, float inv_eta)
{
	const float ci = fabsf(VoH);
	const float ct2 = 1.f - eta * eta * (1.f - ci * ci);
	if (ct2 < 0.0f) return 0.0f;
	const float sd = VoH + inv_eta * LoH;
	return 4 * inv_eta * inv_eta * fabsf(VoH * LoH) / (sd * sd);
}

// GGXSmithBsdf::f_and_p in projected-solid-angle measure (ggx_smith.h:428-479). `transmissive` selects
// the reflection lobe (int_ior = -1) or the transmission lobe (int_ior = ior, ext_ior = 1).
FB_D void ggx_f_and_p(const BsdfParams& b, const Frame& g, bool transmissive, V3 V, V3 L, float NoV, float NoL, float& f, float& p)
{
	const float int_ior = transmissive ? b.ior : -1.0f, ext_ior = transmissive ? 1.0f : -1.0f;
	const float eta = NoV >= 0.0f ? ext_ior / int_ior : int_ior / ext_ior;
	const float inv_eta = NoV >= 0.0f ? int_ior / ext_ior : ext_ior / int_ior;
	const bool is_trans = int_ior > 0.0f;
	const V3 H = microfacet_v(V, L, g.normal_s, NoV, NoL, inv_eta);
	const float NoH = dot(g.normal_s, H);
	const float sgn = is_trans ? -1.0f : 1.0f;
	if (sgn * NoL * NoV <= 0.0f || NoH == 0.0f) { p = 0.0f; f = 0.0f; return; }
	const float D = hvd_ggx_eval(b.inv_roughness, fabsf(NoH), dot(g.tangent, H), dot(g.binormal, H));
	const float G = smith_joint_approx(b.roughness, fabsf(NoV), fabsf(NoL));
	const float G1 = smith_g1v(b.roughness, fabsf(NoV), fabsf(NoL));
	float tf = 1.0f;
	if (is_trans) tf = transmission_factor(dot(V, H), dot(L, H), eta, inv_eta);
	f = clamp_inf(G * D * tf);
	p = clamp_inf(G1 * D * tf);
}

// Bsdf::f_and_p (src/bsdf.h:366-412): per-component f (rgb) and projected-solid-angle pdf, RR = true.
// Outputs the two sums the path tracer needs: diffuse (R+T) and glossy (R+T).
FB_D void bsdf_f_and_p(const BsdfParams& b, const float* __restrict__ table, const Frame& g, V3 V, V3 L,
					   V3& f_diffuse, V3& f_glossy, float& p_diffuse, float& p_glossy)
{
	const V3 N = g.normal_s;
	const float NoV = dot(N, V), NoL = dot(N, L);
	V3 Fc_1, Tc_1;
	V3 w[4];
	// component_weights: cos_theta_i = dot(w_i, H_c) with H_c = N
	if (!clearcoat_transmission(b, dot(V, N), Fc_1, Tc_1)) { w[0] = w[1] = w[2] = w[3] = V3(0.0f); }
	else
	{
		inner_component_weights(b, table, N, V, L, NoV, NoL, w);
		const V3 tc = Tc_1 * V3(1.0f);
		w[0] *= tc; w[1] *= tc; w[2] *= tc; w[3] *= tc;
	}
	const float coat_r = average(Fc_1), coat_t = 1.0f - coat_r;

	// Lambert lobes (lambert.h:99-111, lambert_trans.h)
	const float s = NoL * NoV;
	const bool refl = s > 0.0f, tran = s < 0.0f;
	const V3 f_d = refl ? b.diffuse : V3(0.0f), f_dt = tran ? b.diffuse_trans : V3(0.0f);
	const float p_d = refl ? 1.0f / FB_PI : 0.0f, p_dt = tran ? 1.0f / FB_PI : 0.0f;
	float f_g, p_g, f_gt, p_gt;
	ggx_f_and_p(b, g, false, V, L, NoV, NoL, f_g, p_g);
	ggx_f_and_p(b, g, true, V, L, NoV, NoL, f_gt, p_gt);

	float w_p[4];
	sampling_weights(b, table, NoV, w_p);
	w_p[0] *= coat_t; w_p[1] *= coat_t; w_p[2] *= coat_t; w_p[3] *= coat_t;

	const float factor = compression_factor(b, NoV, NoL);
	const V3 fdr = f_d * w[B_DR] * factor, fdt = f_dt * w[B_DT] * factor;
	const V3 fgr = V3(f_g) * w[B_GR] * factor, fgt = V3(f_gt) * w[B_GT] * factor;
	f_diffuse = fdr + fdt;
	f_glossy = fgr + fgt;
	p_diffuse = p_d * w_p[B_DR] + p_dt * w_p[B_DT];
	p_glossy = p_g * w_p[B_GR] + p_gt * w_p[B_GT];
}

// per-component variant used by the parity harness
FB_D void bsdf_f_and_p_components(const BsdfParams& b, const float* __restrict__ table, const Frame& g, V3 V, V3 L, V3 f[4], float p[4])
{
	const V3 N = g.normal_s;
	const float NoV = dot(N, V), NoL = dot(N, L);
	V3 Fc_1, Tc_1; V3 w[4];
	if (!clearcoat_transmission(b, dot(V, N), Fc_1, Tc_1)) { w[0] = w[1] = w[2] = w[3] = V3(0.0f); }
	else { inner_component_weights(b, table, N, V, L, NoV, NoL, w); const V3 tc = Tc_1 * V3(1.0f); w[0] *= tc; w[1] *= tc; w[2] *= tc; w[3] *= tc; }
	const float coat_t = 1.0f - average(Fc_1);
	const float s = NoL * NoV;
	const bool refl = s > 0.0f, tran = s < 0.0f;
	float f_g, p_g, f_gt, p_gt;
	ggx_f_and_p(b, g, false, V, L, NoV, NoL, f_g, p_g);
	ggx_f_and_p(b, g, true, V, L, NoV, NoL, f_gt, p_gt);
	float w_p[4];
	sampling_weights(b, table, NoV, w_p);
	const float factor = compression_factor(b, NoV, NoL);
	f[B_DR] = (refl ? b.diffuse : V3(0.0f)) * w[B_DR] * factor;
	f[B_DT] = (tran ? b.diffuse_trans : V3(0.0f)) * w[B_DT] * factor;
	f[B_GR] = V3(f_g) * w[B_GR] * factor;
	f[B_GT] = V3(f_gt) * w[B_GT] * factor;
	p[B_DR] = (refl ? 1.0f / FB_PI : 0.0f) * (w_p[B_DR] * coat_t);
	p[B_DT] = (tran ? 1.0f / FB_PI : 0.0f) * (w_p[B_DT] * coat_t);
	p[B_GR] = p_g * (w_p[B_GR] * coat_t);
	p[B_GT] = p_gt * (w_p[B_GT] * coat_t);
}

// sin and cos as a fixed sequence of IEEE fp32 operations (Cody-Waite reduction by pi/2 in three parts, minimax
// polynomials on [-pi/4, pi/4], ~1 ulp): with -fmad=false the host executes exactly the same sequence, which
// makes sampled directions reproducible to the bit between the device and the CPU checker. (The reference calls
// sinf/cosf, whose last-place rounding differs between CUDA and any libm and is not part of its specification.)
FB_D void fb_sincosf(float x, float* s, float* c)
{
	const float kf = rintf(x * 0.636619772f);
	const int k = (int)kf;
	float r = x - kf * 1.5703125f;
	r = r - kf * 4.837512969970703125e-4f;
	r = r - kf * 7.54978995489188216e-8f;
	const float z = r * r;
	const float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
	const float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
	const bool swap = k & 1;
	const float ss = swap ? pc : ps, cc = swap ? ps : pc;
	*s = (k & 2) ? -ss : ss;
	*c = ((k + 1) & 2) ? -cc : cc;
}

FB_D V2 square_to_unit_disk(float sx, float sy)
{
	float phi, r;
	const float a = 2 * sx - 1, bb = 2 * sy - 1;
	if (a > -bb)
	{
		if (a > bb) { r = a; phi = (FB_PI / 4) * (bb / a); }
		else        { r = bb; phi = (FB_PI / 4) * (2 - (a / bb)); }
	}
	else
	{
		if (a < bb) { r = -a; phi = (FB_PI / 4) * (4 + (bb / a)); }
		else        { r = -bb; phi = bb != 0 ? (FB_PI / 4) * (6 - (a / bb)) : 0; }
	}
	float sp, cp;
	fb_sincosf(phi, &sp, &cp);
	return V2(r * cp, r * sp);
}

// vndf_ggx_smith_sample (ggx_common.h:264-290) wrapped by GGXSmithMicrofacetDistribution::sample (ggx_smith.h:114-134)
FB_D V3 ggx_sample_h_local(float alpha, float u0, float u1, V3 Vl)
{
	const float sgn = Vl.z >= 0.0f ? 1.0f : -1.0f;
	const V3 V = normalize(V3(alpha * Vl.x, alpha * Vl.y, Vl.z * sgn));
	const V3 T1 = (V.z < 0.9999f) ? normalize(cross(V, V3(0, 0, 1))) : V3(1, 0, 0);
	const V3 T2 = cross(T1, V);
	const float a = 1.0f / (1.0f + V.z);
	const float r = sqrtf(u0);
	const float phi = (u1 < a) ? u1 / a * FB_PI : FB_PI + (u1 - a) / (1.0f - a) * FB_PI;
	float sp, cp;
	fb_sincosf(phi, &sp, &cp);
	const float P1 = r * cp;
	const float P2 = r * sp * ((u1 < a) ? 1.0f : V.z);
	V3 N = P1 * T1 + P2 * T2 + sqrtf(fmaxf(0.0f, 1.0f - P1 * P1 - P2 * P2)) * V;
	N = normalize(V3(alpha * N.x, alpha * N.y, fmaxf(0.0f, N.z)));
	N.z *= sgn;
	return N;
}

// GGXSmithBsdf::sample given H (ggx_smith.h:529-620)
FB_D void ggx_sample_given_h(const BsdfParams& b, const Frame& geo, bool transmissive, V3 H, V3 V, V3& L, float& g, float& p, float& p_proj)
{
	const float int_ior = transmissive ? b.ior : -1.0f, ext_ior = transmissive ? 1.0f : -1.0f;
	const V3 N = geo.normal_s;
	const float NoV = dot(N, V);
	const float eta = NoV >= 0.0f ? ext_ior / int_ior : int_ior / ext_ior;
	const float inv_eta = NoV >= 0.0f ? int_ior / ext_ior : ext_ior / int_ior;
	const bool is_trans = int_ior > 0.0f;
	if (NoV == 0.0f) { p = 0.0f; p_proj = 0.0f; g = 0.0f; return; }
	if (!is_trans) L = 2 * dot(V, H) * H - V;
	else
	{
		const float VoH = dot(V, H);
		const float ct2 = 1.f - eta * eta * (1.f - VoH * VoH);
		if (ct2 < 0.0f) { L = 2 * dot(V, H) * H - V; p = 0.0f; p_proj = 0.0f; g = 0.0f; return; }
		const float ct = (VoH >= 0.0f ? 1.0f : -1.0f) * sqrtf(ct2);
		L = (eta * VoH - ct) * H - eta * V;
	}
	const float NoL = dot(N, L), NoH = dot(N, H);
	const float sgn = is_trans ? -1.0f : 1.0f;
	if (sgn * NoL * NoV <= 0.0f || NoH == 0.0f) { p = 0.0f; p_proj = 0.0f; g = 0.0f; return; }
	const float D = hvd_ggx_eval(b.inv_roughness, fabsf(NoH), dot(geo.tangent, H), dot(geo.binormal, H));
	const float G = smith_joint_approx(b.roughness, fabsf(NoV), fabsf(NoL));
	const float G1 = smith_g1v(b.roughness, fabsf(NoV), fabsf(NoL));
	float tf = 1.0f;
	if (is_trans) tf = transmission_factor(dot(V, H), dot(L, H), eta, inv_eta);
	p_proj = clamp_inf(G1 * D * tf);
	p = p_proj * fabsf(NoL);
	g = clamp_inf(G / G1);
}

// Bsdf::sample, RR = true, evaluate_full_bsdf = false, all components (src/bsdf.h:921-1199)
FB_D bool bsdf_sample(const BsdfParams& b, const float* __restrict__ table, const Frame& g, float z0, float z1, float z2, V3 in,
					  uint32& out_comp, V3& out, float& out_p, float& out_p_proj, V3& out_g)
{
	const V3 N = g.normal_s;
	const float cos_theta_i = dot(in, N);
	V3 Fc_1, Tc_1;
	out_comp = kAbsorption;
	if (!clearcoat_transmission(b, cos_theta_i, Fc_1, Tc_1))
	{
		out = V3(0.0f); out_p = 0.0f; out_p_proj = 0.0f; out_g = V3(0.0f);
		return false;
	}
	const float coat_r = average(Fc_1), coat_t = 1.0f - coat_r;
	float w_p[4];
	sampling_weights(b, table, cos_theta_i, w_p);

	const V3 V_local = g.to_local(in);
	const V3 H_local = ggx_sample_h_local(b.roughness, z0, z1, V_local);
	const V3 H = g.from_local(H_local);
	V3 r, t;
	const float eta = V_local.z > 0.0f ? 1.0f / b.ior : b.ior;
	fresnel_weights(b, dot(V_local, H_local), eta, r, t);
	w_p[B_GR] = (w_p[B_GR] + max_comp(r)) * 0.5f;
	w_p[B_GT] = (w_p[B_GT] + (1 - b.opacity) * max_comp(t)) * 0.5f;
	w_p[B_DR] = (w_p[B_DR] + b.opacity * max_comp(t * b.diffuse) * FB_PI) * 0.5f;
	w_p[B_DT] = (w_p[B_DT] + b.opacity * max_comp(t * b.diffuse_trans) * FB_PI) * 0.5f;
	w_p[0] *= coat_t; w_p[1] *= coat_t; w_p[2] *= coat_t; w_p[3] *= coat_t;

	V3 gg(0.0f); float p = 0.0f, p_proj = 0.0f, p_comp = 0.0f;
	V3 w_o(0.0f);
	const float c0 = w_p[B_DR], c1 = c0 + w_p[B_GR], c2 = c1 + w_p[B_DT], c3 = c2 + w_p[B_GT], c4 = c3 + coat_r;
	if (z2 < c0 || (z2 >= c1 && z2 < c2))
	{
		// Lambert reflection / transmission (lambert.h:131-153)
		const bool trans = !(z2 < c0);
		p_comp = trans ? w_p[B_DT] : w_p[B_DR];
		const V2 d = square_to_unit_disk(z0, z1);
		const float r2 = d.x * d.x + d.y * d.y;
		float lz = sqrtf(fmaxf(1.0f - r2, 0.0f));
		if (trans ? (cos_theta_i > 0.0f) : (cos_theta_i < 0.0f)) lz = -lz;
		w_o = d.x * g.tangent + d.y * g.binormal + lz * N;
		gg = (trans ? b.diffuse_trans : b.diffuse) * FB_PI;
		p = fabsf(lz) / FB_PI;
		p_proj = 1.0f / FB_PI;
		out_comp = trans ? kDiffuseTransmission : kDiffuseReflection;
	}
	else if (z2 < c1 || (z2 >= c2 && z2 < c3))
	{
		const bool trans = !(z2 < c1);
		p_comp = trans ? w_p[B_GT] : w_p[B_GR];
		float gs;
		ggx_sample_given_h(b, g, trans, H, in, w_o, gs, p, p_proj);
		gg = V3(gs);
		out_comp = trans ? kGlossyTransmission : kGlossyReflection;
	}
	else if (z2 < c4)
	{
		p_comp = coat_r;
		out = 2 * cos_theta_i * N - in;
		gg = Fc_1 / p_comp;
		p_proj = __int_as_float(0x7f800000); p = __int_as_float(0x7f800000);
		out_comp = kClearcoatReflection;
	}

	if (out_comp == kAbsorption) { out = V3(0.0f); out_p = 0.0f; out_p_proj = 0.0f; out_g = V3(0.0f); return false; }

	if (out_comp != kClearcoatReflection)
	{
		gg *= Tc_1 * V3(1.0f);
		out = w_o;
		V3 w[4];
		const float NoL = dot(N, out);
		inner_component_weights(b, table, N, in, out, cos_theta_i, NoL, w);
		gg *= (out_comp & kGlossyReflection) ? w[B_GR] : (out_comp & kGlossyTransmission) ? w[B_GT] : (out_comp & kDiffuseReflection) ? w[B_DR] : w[B_DT];
		gg /= p_comp;
		p *= p_comp;
		p_proj *= p_comp;
	}
	const float factor = compression_factor(b, cos_theta_i, dot(out, N));
	out_p = p; out_p_proj = p_proj; out_g = gg * factor;
	return true;
}

} // namespace fb
