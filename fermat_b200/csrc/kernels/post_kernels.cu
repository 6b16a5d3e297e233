// post_kernels.cu — what RenderingContext::render does AFTER the renderer returns (reference src/renderer.cu:1045-1049):
//
//   filter()            RenderingContextImpl::filter     src/renderer.cu:1099-1160 (the EAW branch that is compiled in)
//     filter_variance   filter_variance_kernel           src/renderer.cu:366-399
//     EAW               EAW_kernel / EAW_mad_kernel      src/eaw.cu:36-247, iteration schedule src/eaw.cu:320-368
//   to_rgba()           to_rgba_kernel                   src/renderer.cu:83-282 (shading modes of renderer_view.h:62-77)
//
// Same arithmetic per pixel, different launch structure. The reference runs 2 x 7 EAW launches per frame and every one
// of the 25 taps of every launch re-derives the tap's normal from its 2x15-bit packing (a sqrt, a sin and a cos) and
// re-clamps/divides its colour by the albedo. Here
//   * the G-buffer is unpacked ONCE per frame into a float4 {normal, miss flag} plane (k_unpack_gbuffer) that all
//     launches read;
//   * the DIFFUSE and SPECULAR channels, which share position/normal/edge-stopping geometry terms, go through the
//     a-trous iterations TOGETHER: one launch per iteration filters both (7 launches instead of 14, the geometry of a
//     tap is fetched and weighted once for both);
//   * taps stream through the read-only path, 32x8 thread tiles keep a warp on one image row (coalesced 512-B rows).
// Every per-pixel value is computed by the same operation sequence as the reference's kernels, so the only
// differences against a CPU evaluation are those of expf / powf / sinf / cosf themselves.
#include "post_kernels.h"

namespace fb {

namespace {

// ---- GBufferView helpers (src/framebuffer.h:84-121) ------------------------------------------------------------
FB_D bool gb_is_miss(const float4 geo) { return (__float_as_uint(geo.w) & (1u << 31)) != 0u; }
FB_D V3 gb_unpack_normal(const float4 geo)
{
	// cugar::unpack_vector<float>(n_i, 15) then uniform_square_to_sphere (contrib/cugar/linalg/vector_inl.h:464-472,
	// spherical/mappings_inline.h:162-172)
	const uint32 n_i = __float_as_uint(geo.w) & ~(1u << 31);
	const float ux = float(n_i & 32767u) / 32767.0f, uy = float(n_i >> 15) / 32767.0f;
	const float cosTheta = uy * 2.0f - 1.0f;
	const float sinTheta = sqrtf(fmaxf(1.0f - cosTheta * cosTheta, 0.0f));
	const float phi = ux * 6.28318530717958647692f;
	return V3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
}

__global__ void __launch_bounds__(256) k_unpack_gbuffer(const float4* __restrict__ geo, float4* __restrict__ normals, const uint32 n)
{
	const uint32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 g = __ldg(geo + i);
	const V3 N = gb_unpack_normal(g);
	normals[i] = make_float4(N.x, N.y, N.z, gb_is_miss(g) ? 1.0f : 0.0f);
}

// ---- filter_variance_kernel (src/renderer.cu:366-390), for C channels at once -----------------------------------
template <int C>
__global__ void __launch_bounds__(256) k_filter_variance(const EawChannels<C> ch, const uint32 res_x, const uint32 res_y, const uint32 FW)
{
	const uint32 x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
	if (x >= res_x || y >= res_y) return;
	const int lx = x > FW ? x - FW : 0, rx = x + FW < res_x ? x + FW : res_x - 1;
	const int ly = y > FW ? y - FW : 0, ry = y + FW < res_y ? y + FW : res_y - 1;
	#pragma unroll
	for (int c = 0; c < C; ++c)
	{
		float variance = 0.0f;
		for (int yy = ly; yy <= ry; yy++)
			for (int xx = lx; xx <= rx; xx++)
				variance += __ldg(ch.img[c] + (size_t)yy * res_x + xx).w;
		variance /= (ry - ly + 1) * (rx - lx + 1);
		ch.var[c][x + y * res_x] = variance;
	}
}

FB_D float4 f4_max(const float4 a, const float m) { return make_float4(fmaxf(a.x, m), fmaxf(a.y, m), fmaxf(a.z, m), fmaxf(a.w, m)); }
FB_D float4 f4_mul(const float4 a, const float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
FB_D float4 f4_div(const float4 a, const float4 b) { return make_float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
FB_D float4 f4_add(const float4 a, const float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ---- one a-trous iteration for C channels ------------------------------------------------------------------------
// MODE 0: EAW_kernel (plain: src -> dst)                                           src/eaw.cu:48-124
// MODE 1: EAW_mad_kernel with kFilterOpDemodulateInput | kFilterOpReplaceMode      src/eaw.cu:128-247, :339-350
// MODE 2: EAW_mad_kernel with kFilterOpModulateOutput  | kFilterOpAddMode          :327-338 (dst += w_img * eaw(src))
// All C channels of MODE 2 add into the SAME dst, channel 0 first — the order in which the reference's two passes
// add the diffuse and then the specular term (src/renderer.cu:1127-1159).
template <int C, int MODE>
__global__ void __launch_bounds__(256) k_eaw(const EawChannels<C> ch, const float4* __restrict__ geo, const float4* __restrict__ normals,
											 const EawParams p, const uint32 res_x, const uint32 res_y, const uint32 step_size)
{
	const uint32 x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
	if (x >= res_x || y >= res_y) return;
	const uint32 pix = x + y * res_x;
	const float kernelWeights[3] = { 1.0f, (float)(2.0 / 3.0), (float)(1.0 / 6.0) };

	float4 weightCenter[C], colorCenter[C];
	#pragma unroll
	for (int c = 0; c < C; ++c)
	{
		const float4 imgCenter = __ldg(ch.src[c] + pix);
		if (MODE != 0) weightCenter[c] = f4_max(__ldg(ch.w_img[c] + pix), p.w_min);
		colorCenter[c] = MODE == 1 ? f4_div(imgCenter, weightCenter[c]) : imgCenter;
	}
	const float4 packed_geo = __ldg(geo + pix);
	const float4 nc = __ldg(normals + pix);
	const V3 normalCenter(nc.x, nc.y, nc.z), positionCenter(packed_geo.x, packed_geo.y, packed_geo.z);
	const V3 U(p.U[0], p.U[1], p.U[2]), Vv(p.V[0], p.V[1], p.V[2]), W(p.W[0], p.W[1], p.W[2]), E(p.E[0], p.E[1], p.E[2]);

	if (nc.w != 0.0f)       // GBufferView::is_miss
	{
		if (MODE == 2)
		{
			float4 r = ch.dst[0][pix];
			#pragma unroll
			for (int c = 0; c < C; ++c) r = f4_add(r, f4_mul(colorCenter[c], weightCenter[c]));
			ch.dst[0][pix] = r;
		}
		else
		{
			#pragma unroll
			for (int c = 0; c < C; ++c) ch.dst[c][pix] = colorCenter[c];
		}
		return;
	}

	// the plain kernel measures the depth from the eye, the mad kernel from the origin (src/eaw.cu:68-70 vs :179-181)
	const float posRadius = 20.0f * fminf(sqrtf(dot(U, U)) / res_x, sqrtf(dot(Vv, Vv)) / res_y) *
							dot(MODE == 0 ? positionCenter - E : positionCenter, W) / dot(W, W);
	const float phiNormal = p.phi_normal * step_size * step_size;
	const float phiPosition = p.phi_position / (posRadius * posRadius);
	float phiColor[C], sumWeight[C];
	V3 sumColor[C];
	#pragma unroll
	for (int c = 0; c < C; ++c)
	{
		const float variance = ch.var[c] ? ch.var[c][pix] : 1.0f;
		phiColor[c] = p.phi_color / fmaxf(1.0e-3f, variance * variance);
		sumWeight[c] = 0.0f; sumColor[c] = V3(0.0f);
	}

	#pragma unroll
	for (int yy = -2; yy <= 2; yy++)
	{
		#pragma unroll
		for (int xx = -2; xx <= 2; xx++)
		{
			const int px = (int)x + xx * (int)step_size, py = (int)y + yy * (int)step_size;
			if (!(px >= 0 && py >= 0 && px < (int)res_x && py < (int)res_y)) continue;
			const uint32 q = (uint32)px + (uint32)py * res_x;
			const float kernel = kernelWeights[xx < 0 ? -xx : xx] * kernelWeights[yy < 0 ? -yy : yy];
			const float4 nP = __ldg(normals + q);
			if (nP.w != 0.0f) continue;
			const float4 geoP = __ldg(geo + q);
			const float d = fmaxf(1e-8f, dot(V3(nP.x, nP.y, nP.z), normalCenter));      // norm_diff (src/eaw.cu:36-42)
			const float wNormal = (1.0f - d) * phiNormal;
			const V3 diffPosition = V3(geoP.x, geoP.y, geoP.z) - positionCenter;
			const float wPosition = dot(diffPosition, diffPosition) * phiPosition;
			#pragma unroll
			for (int c = 0; c < C; ++c)
			{
				const float4 imgP = __ldg(ch.src[c] + q);
				const float4 colorP = MODE == 1 ? f4_div(imgP, f4_max(__ldg(ch.w_img[c] + q), p.w_min)) : imgP;
				const V3 diffCol = V3(colorP.x, colorP.y, colorP.z) - V3(colorCenter[c].x, colorCenter[c].y, colorCenter[c].z);
				const float wColor = dot(diffCol, diffCol) * phiColor[c];
				// (the reference writes `0.0 - ...`: the exponent is summed in double and rounded once)
				const float w = kernel * expf((float)(0.0 - (double)fmaxf(wPosition, 0.0f) - (double)fmaxf(wNormal, 0.0f) - (double)fmaxf(wColor, 0.0f)));
				sumWeight[c] += w;
				sumColor[c] = sumColor[c] + w * V3(colorP.x, colorP.y, colorP.z);
			}
		}
	}

	float4 r = MODE == 2 ? ch.dst[0][pix] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	#pragma unroll
	for (int c = 0; c < C; ++c)
	{
		float4 o = colorCenter[c];
		if (sumWeight[c] != 0.0f) { const V3 m = sumColor[c] / sumWeight[c]; o = make_float4(m.x, m.y, m.z, colorCenter[c].w); }
		if (MODE == 2) r = f4_add(r, f4_mul(o, weightCenter[c]));
		else ch.dst[c][pix] = o;
	}
	if (MODE == 2) ch.dst[0][pix] = r;
}

// ---- to_rgba_kernel (src/renderer.cu:83-273) -------------------------------------------------------------------
FB_D uchar4 quantize_rgba(const float4 c)
{
	return make_uchar4((unsigned char)fminf(c.x * 256.0f, 255.0f), (unsigned char)fminf(c.y * 256.0f, 255.0f),
					   (unsigned char)fminf(c.z * 256.0f, 255.0f), (unsigned char)fminf(c.w * 256.0f, 255.0f));
}
FB_D float4 tonemap(float4 c, const float exposure, const float inv_gamma)
{
	c = make_float4(c.x * exposure, c.y * exposure, c.z * exposure, c.w * exposure);
	c = make_float4(c.x / (c.x + 1.0f), c.y / (c.y + 1.0f), c.z / (c.z + 1.0f), c.w / (c.w + 1.0f));
	return make_float4(powf(c.x, inv_gamma), powf(c.y, inv_gamma), powf(c.z, inv_gamma), powf(c.w, inv_gamma));
}

__global__ void __launch_bounds__(128) k_to_rgba(const FrameBufferView fb, const uint32 mode, const float exposure, const float gamma, uchar4* __restrict__ rgba)
{
	const uint32 idx = threadIdx.x + blockIdx.x * blockDim.x;
	if (idx >= fb.n_pixels) return;
	const float inv_gamma = 1.0f / gamma;
	int tone_channel = -1;
	switch (mode)
	{
	case SHADING_SHADED:          tone_channel = FB_COMPOSITED_C; break;
	case SHADING_FILTERED:        tone_channel = FB_FILTERED_C; break;
	case SHADING_DIFFUSE_COLOR:   tone_channel = FB_DIFFUSE_C; break;
	case SHADING_SPECULAR_COLOR:  tone_channel = FB_SPECULAR_C; break;
	case SHADING_DIRECT_LIGHTING: tone_channel = FB_DIRECT_C; break;
	default: break;
	}
	if (tone_channel >= 0) { rgba[idx] = quantize_rgba(tonemap(fb.channels[tone_channel][idx], exposure, inv_gamma)); return; }
	if (mode == SHADING_ALBEDO) rgba[idx] = quantize_rgba(f4_add(fb.channels[FB_DIFFUSE_A][idx], fb.channels[FB_SPECULAR_A][idx]));
	else if (mode == SHADING_DIFFUSE_ALBEDO) rgba[idx] = quantize_rgba(fb.channels[FB_DIFFUSE_A][idx]);
	else if (mode == SHADING_SPECULAR_ALBEDO) rgba[idx] = quantize_rgba(fb.channels[FB_SPECULAR_A][idx]);
	else if (mode == SHADING_VARIANCE)
	{
		float c = fb.channels[FB_COMPOSITED_C][idx].w;
		c *= exposure; c = c / (c + 1.0f); c = powf(c, inv_gamma);
		rgba[idx] = quantize_rgba(make_float4(c, c, c, c));
	}
	else if (mode == SHADING_UV)
	{
		const float4 c = fb.gb_uv[idx];
		rgba[idx] = quantize_rgba(make_float4(c.z, c.w, 0.5f, 0.0f));
	}
	else if (mode == SHADING_NORMAL)
	{
		const V3 n = gb_unpack_normal(fb.gb_geo[idx]);
		rgba[idx] = make_uchar4((unsigned char)fminf(n.x * 128.0f + 128.0f, 255.0f), (unsigned char)fminf(n.y * 128.0f + 128.0f, 255.0f),
								(unsigned char)fminf(n.z * 128.0f + 128.0f, 255.0f), 0);
	}
	// any other mode (kUVStretch, kCharts without charts, kAux* without auxiliary channels): the reference's kernel writes nothing
}

} // anonymous namespace

cudaError_t launch_unpack_gbuffer(const FrameBufferView& fb, float4* normals, cudaStream_t s)
{
	if (fb.n_pixels == 0) return cudaSuccess;
	k_unpack_gbuffer<<<(fb.n_pixels + 255u) / 256u, 256, 0, s>>>(fb.gb_geo, normals, fb.n_pixels);
	return cudaGetLastError();
}

static inline dim3 grid2d(uint32 rx, uint32 ry) { return dim3((rx + 31u) / 32u, (ry + 7u) / 8u); }

cudaError_t launch_filter_variance2(const EawChannels<2>& ch, uint32 res_x, uint32 res_y, uint32 FW, cudaStream_t s)
{
	if (res_x == 0 || res_y == 0) return cudaSuccess;
	k_filter_variance<2><<<grid2d(res_x, res_y), dim3(32, 8), 0, s>>>(ch, res_x, res_y, FW);
	return cudaGetLastError();
}

cudaError_t launch_eaw2(int mode, const EawChannels<2>& ch, const float4* geo, const float4* normals, const EawParams& p,
						uint32 res_x, uint32 res_y, uint32 step_size, cudaStream_t s)
{
	if (res_x == 0 || res_y == 0) return cudaSuccess;
	const dim3 g = grid2d(res_x, res_y), b(32, 8);
	if (mode == 0) k_eaw<2, 0><<<g, b, 0, s>>>(ch, geo, normals, p, res_x, res_y, step_size);
	else if (mode == 1) k_eaw<2, 1><<<g, b, 0, s>>>(ch, geo, normals, p, res_x, res_y, step_size);
	else if (mode == 2) k_eaw<2, 2><<<g, b, 0, s>>>(ch, geo, normals, p, res_x, res_y, step_size);
	else return cudaErrorInvalidValue;
	return cudaGetLastError();
}

cudaError_t launch_to_rgba(const FrameBufferView& fb, uint32 mode, float exposure, float gamma, uchar4* rgba, cudaStream_t s)
{
	if (fb.n_pixels == 0) return cudaSuccess;
	k_to_rgba<<<(fb.n_pixels + 127u) / 128u, 128, 0, s>>>(fb, mode, exposure, gamma, rgba);
	return cudaGetLastError();
}

} // namespace fb
