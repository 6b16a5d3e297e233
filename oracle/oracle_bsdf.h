// oracle_bsdf.h — TEST INFRASTRUCTURE ONLY. Scalar CPU restatement of Fermat's layered BSDF as the
// `-pt` renderer uses it. Nothing under fermat_b200/ may include this file.
//
// Follows (paths relative to the Fermat repository):
//   src/bsdf.h:219-243        Bsdf::Bsdf(transport, renderer, material)
//   src/bsdf.h:366-412        Bsdf::f_and_p (array form)
//   src/bsdf.h:530-627        sampling_weights / normalize_sampling_weights
//   src/bsdf.h:632-792        fresnel_weights / inner_component_weights / component_weights
//   src/bsdf.h:921-1199       Bsdf::sample (USE_EFFICIENT_SAMPLER_WITH_APPROXIMATE_PDFS = 1, :53)
//   src/bsdf.h:1202-1268      clearcoat_transmission / compression_factor / glossy_reflectance
//   contrib/cugar/bsdf/ggx_smith.h:54-196, 204-620    GGXSmithMicrofacetDistribution, GGXSmithBsdf
//   contrib/cugar/bsdf/ggx_common.h:50-110, 264-290   microfacet, vndf_microfacet, hvd_ggx_eval, vndf_ggx_smith_sample
//   contrib/cugar/bsdf/lambert.h:64-153, lambert_trans.h   Lambert lobes
//   contrib/cugar/bsdf/refraction.h:50-183            fresnel_dielectric, fresnel_schlick, refract
//   contrib/cugar/spherical/mappings_inline.h:56-127  square_to_unit_disk, square_to_cosine_hemisphere
// Pinned against the reference's own headers compiled verbatim (oracle/_ref, see oracle/Makefile)
// through tests/golden/bsdf_golden.bin.
#pragma once
#include "oracle_math.h"

namespace oracle {

static const float PI_F = 3.14159265358979323846f;

struct Geom   // cugar::DifferentialGeometry + position/texcoords (src/vertex.h:92-98)
{
	vec3 normal_s, normal_g, tangent, binormal, position;
	float st[2];
	vec3 to_local(vec3 v) const { return vec3(dot(v, tangent), dot(v, binormal), dot(v, normal_s)); }
	vec3 from_local(vec3 v) const { return v.x * tangent + v.y * binormal + v.z * normal_s; }
};

enum { kDR = 0, kDT = 1, kGR = 2, kGT = 3 };   // component indices, src/bsdf.h:93-101
enum { cAbsorption = 0, cDiffuseReflection = 1, cDiffuseTransmission = 2, cGlossyReflection = 4, cGlossyTransmission = 8, cClearcoatReflection = 16,
	   cDiffuseMask = 3, cGlossyMask = 12 };

// ---- lobes ---------------------------------------------------------------------------------

inline vec2 square_to_unit_disk(vec2 seed)
{
	float phi, r;
	const float a = 2 * seed.x - 1, b = 2 * seed.y - 1;
	if (a > -b)
	{
		if (a > b) { r = a; phi = (PI_F / 4) * (b / a); }
		else       { r = b; phi = (PI_F / 4) * (2 - (a / b)); }
	}
	else
	{
		if (a < b) { r = -a; phi = (PI_F / 4) * (4 + (b / a)); }
		else       { r = -b; phi = b != 0 ? (PI_F / 4) * (6 - (a / b)) : 0; }
	}
	float sp, cp;
	o_sincosf(phi, &sp, &cp);
	return vec2(r * cp, r * sp);
}
inline vec3 square_to_cosine_hemisphere(vec2 uv)
{
	const vec2 d = square_to_unit_disk(uv);
	const float r2 = d.x * d.x + d.y * d.y;
	return vec3(d.x, d.y, sqrtf(fmaxf(1.0f - r2, 0.0f)));
}

// Lambert reflection (transmission = false) or transmission (true)
inline void lambert_f_and_p(bool trans, vec3 color, const Geom& g, vec3 V, vec3 L, vec3& f, float& p)
{
	const float s = dot(g.normal_s, L) * dot(g.normal_s, V);
	const bool on = trans ? (s < 0.0f) : (s > 0.0f);
	f = on ? color : vec3(0.0f);
	p = on ? 1.0f / PI_F : 0.0f;
}
inline void lambert_sample(bool trans, vec3 color, vec2 u, const Geom& geo, vec3 V, vec3& L, vec3& g, float& p, float& p_proj)
{
	vec3 l = square_to_cosine_hemisphere(u);
	const float NoV = dot(V, geo.normal_s);
	if (trans ? (NoV > 0.0f) : (NoV < 0.0f)) l.z = -l.z;
	L = l.x * geo.tangent + l.y * geo.binormal + l.z * geo.normal_s;
	g = color * PI_F;
	p = fabsf(l.z) / PI_F;
	p_proj = 1.0f / PI_F;
}

inline vec3 vndf_microfacet(vec3 V, vec3 L, vec3 N, float inv_eta)
{
	vec3 H = (dot(V, N) * dot(L, N) >= 0.0f) ? V + L : V + L * inv_eta;
	if (dot(H, H) < 1.0e-12f) return N;
	if (dot(V, H) < 0.0f) H = -H;
	return normalize(H);
}
inline vec3 microfacet(vec3 V, vec3 L, vec3 N, float inv_eta)
{
	vec3 H = (dot(V, N) * dot(L, N) >= 0.0f) ? V + L : V + L * inv_eta;
	if (dot(H, H) == 0.0f) return N;
	if (dot(N, H) < 0.0f) H = -H;
	return normalize(H);
}
inline float hvd_ggx_eval(float inv_alpha, float nh, float ht, float hb)
{
	const float x = ht * inv_alpha, y = hb * inv_alpha;
	const float aniso = x * x + y * y;
	const float f = aniso + nh * nh;
	return (1.0f / PI_F) * inv_alpha * inv_alpha / (f * f);
}
inline vec3 vndf_ggx_smith_sample(vec2 s, float alpha, vec3 _V)
{
	const vec3 V = normalize(vec3(alpha * _V.x, alpha * _V.y, _V.z));
	const vec3 T1 = (V.z < 0.9999f) ? normalize(cross(V, vec3(0, 0, 1))) : vec3(1, 0, 0);
	const vec3 T2 = cross(T1, V);
	const float a = 1.0f / (1.0f + V.z);
	const float r = sqrtf(s.x);
	const float phi = (s.y < a) ? s.y / a * PI_F : PI_F + (s.y - a) / (1.0f - a) * PI_F;
	float sp, cp;
	o_sincosf(phi, &sp, &cp);
	const float P1 = r * cp;
	const float P2 = r * sp * ((s.y < a) ? 1.0f : V.z);
	vec3 N = P1 * T1 + P2 * T2 + sqrtf(fmaxf(0.0f, 1.0f - P1 * P1 - P2 * P2)) * V;
	N = normalize(vec3(alpha * N.x, alpha * N.y, fmaxf(0.0f, N.z)));
	return N;
}

struct GGXSmith   // cugar::GGXSmithBsdf
{
	float roughness, inv_roughness, int_ior, ext_ior;
	GGXSmith() {}
	GGXSmith(float r, bool transmission = false, float ii = 1.0f, float ei = 1.0f)
		: roughness(r), inv_roughness(1.0f / r), int_ior(transmission ? ii : -1.0f), ext_ior(transmission ? ei : -1.0f) {}
	bool  is_transmissive() const { return int_ior > 0.0f; }
	float get_eta(float NoV) const { return NoV >= 0.0f ? ext_ior / int_ior : int_ior / ext_ior; }
	float get_inv_eta(float NoV) const { return NoV >= 0.0f ? int_ior / ext_ior : ext_ior / int_ior; }
	static float clamp_inf(float p) { return (!std::isfinite(p) || std::isnan(p)) ? 1.0e8f : fmaxf(p, 0.0f); }
	float smith_joint_approx(float NoV, float NoL) const        // PredividedSmithJointApprox
	{
		const float a = roughness;
		const float vis_v = NoL * (NoV * (1 - a) + a), vis_l = NoV * (NoL * (1 - a) + a);
		return 0.5f * 1.0f / (vis_v + vis_l);
	}
	float smith_g1v(float NoV, float NoL) const                 // PredividedSmithG1V
	{
		const float a2 = roughness * roughness;
		const float G_V = NoV + sqrtf((NoV - NoV * a2) * NoV + a2);
		return 0.5f / (G_V * NoL);
	}
	float transmission_factor(float VoH, float LoH, float eta, float inv_eta) const   // dwo_dh_transmission_factor
	{
		const float ci = fabsf(VoH);
		const float ct2 = 1.f - eta * eta * (1.f - ci * ci);
		if (ct2 < 0.0f) return 0.0f;
		const float sd = VoH + inv_eta * LoH;
		return 4 * inv_eta * inv_eta * fabsf(VoH * LoH) / (sd * sd);
	}
	void f_and_p(const Geom& g, vec3 V, vec3 L, vec3& f, float& p) const   // projected solid angle measure
	{
		const vec3 N = g.normal_s;
		const float NoL = dot(N, L), NoV = dot(N, V);
		const float eta = get_eta(NoV), inv_eta = get_inv_eta(NoV);
		const vec3 H = vndf_microfacet(V, L, N, inv_eta);
		const float NoH = dot(N, H);
		const float sgn = is_transmissive() ? -1.0f : 1.0f;
		if (sgn * NoL * NoV <= 0.0f || NoH == 0.0f) { p = 0.0f; f = vec3(0.0f); return; }
		const float D = hvd_ggx_eval(inv_roughness, fabsf(NoH), dot(g.tangent, H), dot(g.binormal, H));
		const float G = smith_joint_approx(fabsf(NoV), fabsf(NoL));
		const float G1 = smith_g1v(fabsf(NoV), fabsf(NoL));
		float tf = 1.0f;
		if (is_transmissive()) tf = transmission_factor(dot(V, H), dot(L, H), eta, inv_eta);
		f = vec3(clamp_inf(G * D * tf));
		p = clamp_inf(G1 * D * tf);
	}
	// sample L given the microfacet H (ggx_smith.h:529-620)
	void sample_given_h(const Geom& geo, vec3 H, vec3 V, vec3& L, vec3& g, float& p, float& p_proj) const
	{
		const vec3 N = geo.normal_s;
		const float NoV = dot(N, V);
		const float eta = get_eta(NoV), inv_eta = get_inv_eta(NoV);
		if (NoV == 0.0f) { p = 0.0f; p_proj = 0.0f; g = vec3(0.0f); return; }
		if (!is_transmissive()) L = 2 * dot(V, H) * H - V;
		else
		{
			const float VoH = dot(V, H);
			const float ct2 = 1.f - eta * eta * (1.f - VoH * VoH);
			if (ct2 < 0.0f) { L = 2 * dot(V, H) * H - V; p = 0.0f; p_proj = 0.0f; g = vec3(0.0f); return; }
			const float ct = (VoH >= 0.0f ? 1.0f : -1.0f) * sqrtf(ct2);
			L = (eta * VoH - ct) * H - eta * V;
		}
		const float NoL = dot(N, L), NoH = dot(N, H);
		const float sgn = is_transmissive() ? -1.0f : 1.0f;
		if (sgn * NoL * NoV <= 0.0f || NoH == 0.0f) { p = 0.0f; p_proj = 0.0f; g = vec3(0.0f); return; }
		const float D = hvd_ggx_eval(inv_roughness, fabsf(NoH), dot(geo.tangent, H), dot(geo.binormal, H));
		const float G = smith_joint_approx(fabsf(NoV), fabsf(NoL));
		const float G1 = smith_g1v(fabsf(NoV), fabsf(NoL));
		float tf = 1.0f;
		if (is_transmissive()) tf = transmission_factor(dot(V, H), dot(L, H), eta, inv_eta);
		p_proj = clamp_inf(G1 * D * tf);
		p = p_proj * fabsf(NoL);
		g = vec3(clamp_inf(G / G1));
	}
};

// GGXSmithMicrofacetDistribution::sample(u, V_local) (ggx_smith.h:114-134)
inline vec3 ggx_distribution_sample(float roughness, vec2 u, vec3 V)
{
	const float sgn = V.z >= 0.0f ? 1.0f : -1.0f;
	vec3 H = vndf_ggx_smith_sample(u, roughness, vec3(V.x, V.y, V.z * sgn));
	H.z *= sgn;
	return H;
}

// ---- Fresnel -------------------------------------------------------------------------------
inline float fresnel_dielectric(float ci, float ct, float eta)
{
	if (eta == 1.0f) return 0.0f;
	const float Rs = (ci - eta * ct) / (ci + eta * ct);
	const float Rp = (eta * ci - ct) / (eta * ci + ct);
	return 0.5f * (Rs * Rs + Rp * Rp);
}
inline vec3 fresnel_schlick(float ci, float eta, vec3 base)
{
	ci = saturate(fabsf(ci));
	const float ct2 = saturate(1.f - eta * eta * (1.f - ci * ci));
	if (ct2 < 0.0f) return vec3(1.0f);
	const float c = eta > 1.0f ? sqrtf(ct2) : ci;
	const float x = 1 - c, x2 = x * x;
	const float Fc = x2 * x2 * x;
	return vec3(Fc) + (1 - Fc) * base;
}
inline bool refract(vec3 w_i, vec3 N, float ci, float eta, vec3* out, float* F)
{
	if (eta == 1.0f) { *out = -w_i; *F = 0.0f; return true; }
	const float ct2 = 1.f - eta * eta * (1.f - ci * ci);
	if (ct2 < 0.0f) return false;
	const float ct = (ci >= 0.0f ? -1.0f : 1.0f) * sqrtf(ct2);
	*F = fresnel_dielectric(fabsf(ci), fabsf(ct), eta);
	*out = (eta * ci + ct) * N - eta * w_i;
	return true;
}

// ---- the layered Bsdf -----------------------------------------------------------------------
struct Material   // mirrors MeshMaterial's first 112 bytes after texturing
{
	vec3 diffuse, diffuse_trans, specular, emissive, reflectivity;
	float roughness, ior, opacity;
};

struct Bsdf
{
	vec3 diffuse, diffuse_trans;         // Lambert colours (already / pi)
	GGXSmith glossy, glossy_trans;
	vec3 fresnel, reflectivity;
	float ior, opacity, clearcoat_ior;
	const float* table;

	Bsdf(const Material& m, const float* glossy_reflectance_table)
	{
		diffuse = m.diffuse / PI_F;
		diffuse_trans = m.diffuse_trans / PI_F;
		glossy = GGXSmith(fmaxf(m.roughness * 1.0f + 0.0f, 0.0f));    // mollification 1, bias 0, min_roughness 0
		glossy_trans = GGXSmith(m.roughness, true, m.ior, 1.0f);
		fresnel = m.specular / PI_F;
		reflectivity = m.reflectivity;
		ior = m.ior; opacity = m.opacity; table = glossy_reflectance_table;
		const float R0 = fminf(max_comp(reflectivity), 0.95f);
		clearcoat_ior = (1 + sqrtf(R0)) / (1 - sqrtf(R0));
	}

	float glossy_reflectance(float cos_theta) const
	{
		const uint32_t S = 32;
		const float eta = cos_theta > 0.0f ? 1.0f / ior : ior;
		// float -> uint32 with the device's saturating semantics (cvt.rzi.u32.f32): NaN / negative -> 0, overflow -> max
		auto q = [](float v) { const uint32_t i = !(v > 0.0f) ? 0u : (v >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)v); return i < 31u ? i : 31u; };
		const uint32_t ci = q(fabsf(cos_theta) * (S - 1));
		const uint32_t bi = q(max_comp(fresnel) * (S - 1));
		const uint32_t ei = q((eta / 2.0f) * (S - 1));
		const uint32_t ri = q(glossy.roughness * (S - 1));
		return table[ei * S * S * S + bi * S * S + ri * S + ci];
	}

	bool clearcoat_transmission(const Geom& g, vec3 w_i, vec3& H, float& cos_theta_i, vec3& Fc_1, vec3& Tc_1) const
	{
		const float R0 = fminf(max_comp(reflectivity), 0.95f);
		const float eta_c = 1.0f / clearcoat_ior;
		H = g.normal_s;
		cos_theta_i = dot(w_i, H);
		vec3 w_t; float F;
		if (!refract(w_i, H, cos_theta_i, eta_c, &w_t, &F)) { Fc_1 = vec3(1.0f); Tc_1 = vec3(0.0f); return false; }
		Fc_1 = lerp(reflectivity, vec3(1.0f), fmaxf(F - R0, 0.0f) / (1 - R0));
		Tc_1 = 1.0f - Fc_1;
		return true;
	}

	float compression_factor(const Geom& g, vec3 w_i, vec3 w_o) const
	{
		if (ior)   // radiance transport
		{
			const float NoV = dot(w_i, g.normal_s), NoL = dot(w_o, g.normal_s);
			if (NoV * NoL < 0.0f) return sqr(NoV > 0.0f ? ior : 1.0f / ior);
		}
		return 1.0f;
	}

	void fresnel_weights(float VoH, float eta, vec3& r, vec3& t) const
	{
		if (eta == 0.0f) { r = vec3(0.0f); t = vec3(1.0f); }
		else { r = fresnel_schlick(VoH, eta, fresnel); t = vec3(1.0f - max_comp(r)); }
	}

	void sampling_weights(const Geom& g, vec3 V, float w[4]) const
	{
		const float NoV_signed = dot(g.normal_s, V);
		vec3 r, t;
		if (ior == 0) { r = vec3(0.0f); t = vec3(1.0f); }
		else { r = vec3(glossy_reflectance(NoV_signed)); t = vec3(1.0f - max_comp(r)); }
		w[kGR] = max_comp(r);
		w[kGT] = (1 - opacity) * max_comp(t);
		w[kDR] = opacity * max_comp(t * diffuse) * PI_F;
		w[kDT] = opacity * max_comp(t * diffuse_trans) * PI_F;
	}

	void inner_component_weights(const Geom& g, vec3 V, vec3 L, vec3 w[4]) const
	{
		float eta = 0.0f, inv_eta = 0.0f, VoH = 0.0f;
		if (ior)
		{
			const vec3 N = g.normal_s;
			eta = dot(N, V) > 0.0f ? 1.0f / ior : ior;
			inv_eta = dot(N, V) > 0.0f ? ior : 1.0f / ior;
			const vec3 H = microfacet(V, L, N, inv_eta);
			VoH = dot(V, H);
		}
		vec3 r, t;
		fresnel_weights(VoH, eta, r, t);
		const float dw = (1.0f - glossy_reflectance(dot(g.normal_s, V))) * (1.0f - glossy_reflectance(dot(g.normal_s, L)));
		w[kGR] = r;
		w[kGT] = t * (1 - opacity);
		w[kDR] = t * opacity * dw;
		w[kDT] = t * opacity * dw;
	}

	// returns false on clearcoat TIR; NOTE: in that case only w[] is defined (all zero), Fc_1 = 1
	void component_weights(const Geom& g, vec3 w_i, vec3 w_o, vec3& Fc_1, vec3& Tc_1, vec3 w[4]) const
	{
		vec3 H; float ci;
		if (!clearcoat_transmission(g, w_i, H, ci, Fc_1, Tc_1)) { w[0] = w[1] = w[2] = w[3] = vec3(0.0f); return; }
		const vec3 Tc_2 = 1.0f - vec3(0.0f);
		inner_component_weights(g, w_i, w_o, w);
		for (int i = 0; i < 4; ++i) w[i] *= Tc_1 * Tc_2;
	}

	// f[4] (rgb per component) and p[4], projected solid angle, RR = true, all components
	void f_and_p(const Geom& g, vec3 w_i, vec3 w_o, vec3 f[4], float p[4]) const
	{
		vec3 Fc_1, Tc_1, w[4];
		component_weights(g, w_i, w_o, Fc_1, Tc_1, w);
		float coat_r = average(Fc_1);
		float coat_t = 1.0f - coat_r;
		vec3 f_d, f_g, f_dt, f_gt; float p_d, p_g, p_dt, p_gt;
		lambert_f_and_p(false, diffuse, g, w_i, w_o, f_d, p_d);
		lambert_f_and_p(true, diffuse_trans, g, w_i, w_o, f_dt, p_dt);
		glossy.f_and_p(g, w_i, w_o, f_g, p_g);
		glossy_trans.f_and_p(g, w_i, w_o, f_gt, p_gt);
		float w_p[4];
		sampling_weights(g, w_i, w_p);
		for (int i = 0; i < 4; ++i) w_p[i] *= coat_t;     // normalize_sampling_weights with RR = true
		p[kDR] = p_d * w_p[kDR]; p[kDT] = p_dt * w_p[kDT]; p[kGR] = p_g * w_p[kGR]; p[kGT] = p_gt * w_p[kGT];
		const float factor = compression_factor(g, w_i, w_o);
		f[kDR] = f_d * w[kDR] * factor; f[kDT] = f_dt * w[kDT] * factor;
		f[kGR] = f_g * w[kGR] * factor; f[kGT] = f_gt * w[kGT] * factor;
	}

	// Bsdf::sample with RR = true, evaluate_full_bsdf = false, all components
	bool sample(const Geom& g, const float z[3], vec3 in, uint32_t& out_comp, vec3& out, float& out_p, float& out_p_proj, vec3& out_g) const
	{
		vec3 gg(0.0f); float p = 0.0f, p_proj = 0.0f, p_comp = 0.0f;
		vec3 w_i = in, w_o(0.0f);
		vec3 H_c, Fc_1, Tc_1; float cos_theta_i;
		if (!clearcoat_transmission(g, in, H_c, cos_theta_i, Fc_1, Tc_1))
		{
			out = vec3(0.0f); out_p = 0.0f; out_p_proj = 0.0f; out_g = vec3(0.0f); out_comp = cAbsorption;
			return false;
		}
		float coat_r = average(Fc_1);
		float coat_t = 1.0f - coat_r;
		float w_p[4];
		sampling_weights(g, in, w_p);

		// sample the GGX microfacet distribution up-front and blend a-priori / H-dependent lobe weights
		const vec3 V_local = g.to_local(w_i);
		const vec3 H_local = ggx_distribution_sample(glossy.roughness, vec2(z[0], z[1]), V_local);
		const vec3 H = g.from_local(H_local);
		vec3 r, t;
		const float eta = V_local.z > 0.0f ? 1.0f / ior : ior;
		fresnel_weights(dot(V_local, H_local), eta, r, t);
		w_p[kGR] = (w_p[kGR] + max_comp(r)) * 0.5f;
		w_p[kGT] = (w_p[kGT] + (1 - opacity) * max_comp(t)) * 0.5f;
		w_p[kDR] = (w_p[kDR] + opacity * max_comp(t * diffuse) * PI_F) * 0.5f;
		w_p[kDT] = (w_p[kDT] + opacity * max_comp(t * diffuse_trans) * PI_F) * 0.5f;
		for (int i = 0; i < 4; ++i) w_p[i] *= coat_t;

		if (z[2] < w_p[kDR])
		{
			p_comp = w_p[kDR];
			lambert_sample(false, diffuse, vec2(z[0], z[1]), g, w_i, w_o, gg, p, p_proj);
			out_comp = cDiffuseReflection;
		}
		else if (z[2] < w_p[kDR] + w_p[kGR])
		{
			p_comp = w_p[kGR];
			glossy.sample_given_h(g, H, w_i, w_o, gg, p, p_proj);
			out_comp = cGlossyReflection;
		}
		else if (z[2] < w_p[kDR] + w_p[kGR] + w_p[kDT])
		{
			p_comp = w_p[kDT];
			lambert_sample(true, diffuse_trans, vec2(z[0], z[1]), g, w_i, w_o, gg, p, p_proj);
			out_comp = cDiffuseTransmission;
		}
		else if (z[2] < w_p[kDR] + w_p[kGR] + w_p[kDT] + w_p[kGT])
		{
			p_comp = w_p[kGT];
			glossy_trans.sample_given_h(g, H, w_i, w_o, gg, p, p_proj);
			out_comp = cGlossyTransmission;
		}
		else if (z[2] < w_p[kDR] + w_p[kGR] + w_p[kDT] + w_p[kGT] + coat_r)
		{
			p_comp = coat_r;
			out = 2 * cos_theta_i * H_c - in;
			gg = Fc_1 / p_comp;
			p_proj = INFINITY; p = INFINITY;
			out_comp = cClearcoatReflection;
		}
		else out_comp = cAbsorption;

		if (out_comp != cAbsorption && out_comp != cClearcoatReflection)
		{
			const vec3 Tc_2 = 1.0f - vec3(0.0f);
			gg *= Tc_1 * Tc_2;
			out = w_o;
		}
		if (out_comp != cAbsorption)
		{
			if (out_comp != cClearcoatReflection)
			{
				vec3 w[4];
				inner_component_weights(g, in, out, w);
				gg *= (out_comp & cGlossyReflection) ? w[kGR] : (out_comp & cGlossyTransmission) ? w[kGT] : (out_comp & cDiffuseReflection) ? w[kDR] : w[kDT];
				gg /= p_comp;
				p *= p_comp;
				p_proj *= p_comp;
			}
			const float factor = compression_factor(g, in, out);
			out_p = p; out_p_proj = p_proj; out_g = gg * factor;
			return true;
		}
		out = vec3(0.0f); out_p = 0.0f; out_p_proj = 0.0f; out_g = vec3(0.0f);
		return false;
	}
};

} // namespace oracle
