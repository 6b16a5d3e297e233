// post_kernels.h — host-callable launchers of the post-render kernels (post_kernels.cu): EAW denoiser and RGBA conversion.
#pragma once
#include "device_scene.h"
#include <cuda_runtime.h>

namespace fb {

// ShadingMode (reference src/renderer_view.h:62-77)
enum ShadingMode
{
	SHADING_SHADED = 0, SHADING_UV = 1, SHADING_UV_STRETCH = 2, SHADING_CHARTS = 3, SHADING_ALBEDO = 4, SHADING_DIFFUSE_ALBEDO = 5,
	SHADING_SPECULAR_ALBEDO = 6, SHADING_DIFFUSE_COLOR = 7, SHADING_SPECULAR_COLOR = 8, SHADING_DIRECT_LIGHTING = 9,
	SHADING_FILTERED = 10, SHADING_VARIANCE = 11, SHADING_NORMAL = 12, SHADING_AUX0 = 13
};

// EAWParams (reference src/eaw.h:43-53) + the weight floor the iteration schedule passes (1e-4, src/eaw.cu:331,344)
struct EawParams
{
	float phi_normal, phi_position, phi_color, w_min;
	float E[3], U[3], V[3], W[3];
};

// the C image channels one launch filters together; per channel: source, destination (for the accumulating last
// iteration every channel adds into dst[0]), the albedo image that de/modulates it, its filtered-variance plane
template <int C>
struct EawChannels
{
	const float4* src[C];
	float4*       dst[C];
	const float4* w_img[C];
	float*        var[C];
	const float4* img[C];         // filter_variance: the channel whose .w holds the variance estimate
};

// normals[i] = {unpacked shading normal of pixel i, 1 if the primary ray missed else 0}
cudaError_t launch_unpack_gbuffer(const FrameBufferView& fb, float4* normals, cudaStream_t s);
cudaError_t launch_filter_variance2(const EawChannels<2>& ch, uint32 res_x, uint32 res_y, uint32 FW, cudaStream_t s);
// mode 0: plain iteration, 1: first iteration (input divided by the albedo), 2: last iteration (dst[0] += albedo * filtered)
cudaError_t launch_eaw2(int mode, const EawChannels<2>& ch, const float4* geo, const float4* normals, const EawParams& p,
						uint32 res_x, uint32 res_y, uint32 step_size, cudaStream_t s);
cudaError_t launch_to_rgba(const FrameBufferView& fb, uint32 mode, float exposure, float gamma, uchar4* rgba, cudaStream_t s);

} // namespace fb
