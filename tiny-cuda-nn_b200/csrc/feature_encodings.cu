// feature_encodings.cu -- parameter-free encodings on their own (general path, encoding tier, nested inside a Composite).
//
//   Identity            encodings/identity.h:46-91            y_j = x_j * scale + offset
//   Frequency           encodings/frequency.h:46-103          y = sin(2^f pi x + (k odd ? pi/2 : 0)), fast-math __sinf like the reference
//   TriangleWave        encodings/triangle_wave.h:46-105      y = |v - floor(v) - 1/2| * 4 - 1, v = 2^(f-1) x + f / 4
//   OneBlob             encodings/oneblob.h:47-160            quartic kernel CDF differences per bin, wrapping around [0, 1)
//   SphericalHarmonics  encodings/spherical_harmonics.h:44-100, common_device.h:476-... (sh_enc / sh_enc_grad): real SH of the
//                       direction 2x - 1 up to degree 8. The reference hard-codes the 64 polynomials; here they come from the
//                       recurrences they were generated from (Sloan, "Stupid Spherical Harmonics Tricks", appendix A1): Legendre factors
//                       as polynomials in z, azimuthal factors as Re / Im of (x + i y)^m -- the same polynomials, valid for non-unit
//                       directions too, evaluated in fp32 (agreement with the reference: fp32 rounding, i.e. within one fp16 ulp).
// One thread per (sample, segment); every output is a function of at most three inputs, so these kernels are pure streaming:
// n * (4 D + 2 width) bytes.
#include "feature_encodings.h"

#include <cmath>

namespace tcnnb {

namespace {

constexpr float PI_F = 3.14159265358979323846f;

__device__ __forceinline__ float quartic_cdf(float x, float inv_radius) {  // common_device.h:1090-1095
	const float u = x * inv_radius;
	const float u2 = u * u;
	const float u4 = u2 * u2;
	return fmaxf(0.0f, fminf(1.0f, (15.0f / 16.0f) * u * (1.0f - (2.0f / 3.0f) * u2 + (1.0f / 5.0f) * u4) + 0.5f));
}

__device__ __forceinline__ float quartic_cdf_deriv(float x, float inv_radius) {  // common_device.h:1080-1088
	const float u = x * inv_radius;
	const float tmp = fmaxf(1.0f - u * u, 0.0f);
	return (15.0f / 16.0f) * tmp * tmp * inv_radius;
}

// wrapped CDF (and its derivative) of the kernel centred on x, evaluated at boundary b (oneblob.h:105-118)
__device__ __forceinline__ float wrapped_cdf(float b, float x, float n_bins) { return quartic_cdf(b - x, n_bins) + quartic_cdf(b - x - 1.0f, n_bins) + quartic_cdf(b - x + 1.0f, n_bins); }
__device__ __forceinline__ float wrapped_cdf_deriv(float b, float x, float n_bins) {
	return quartic_cdf_deriv(b - x, n_bins) + quartic_cdf_deriv(b - x - 1.0f, n_bins) + quartic_cdf_deriv(b - x + 1.0f, n_bins);
}

// Real spherical harmonics Y_l^m, l < degree, index l (l + 1) + m, by recurrence. `emit(index, value, d/dx, d/dy, d/dz)`.
// K_l^m = sqrt((2l + 1) (l - m)! / (4 pi (l + m)!)); P_m^m = (1 - 2m) P_{m-1}^{m-1}; P_{m+1}^m = (2m + 1) z P_m^m;
// P_l^m = ((2l - 1) z P_{l-1}^m - (l + m - 1) P_{l-2}^m) / (l - m); c_m + i s_m = (x + i y)^m.
template <bool GRAD, typename Emit>
__device__ __forceinline__ void spherical_harmonics(uint32_t degree, float x, float y, float z, Emit&& emit) {
	float c = 1.0f, s = 0.0f;        // c_m, s_m
	float c_prev = 0.0f, s_prev = 0.0f;  // c_{m-1}, s_{m-1}
	float pmm = 1.0f;                // P_m^m
	for (uint32_t m = 0; m < degree; ++m) {
		if (m > 0) {
			c_prev = c;
			s_prev = s;
			c = x * c_prev - y * s_prev;
			s = x * s_prev + y * c_prev;
			pmm *= 1.0f - 2.0f * (float)m;
		}
		// derivatives of the azimuthal factors
		const float dc_dx = (float)m * c_prev, dc_dy = -(float)m * s_prev, ds_dx = (float)m * s_prev, ds_dy = (float)m * c_prev;
		float p_lm2 = 0.0f, p_lm1 = 0.0f, dp_lm2 = 0.0f, dp_lm1 = 0.0f;
		for (uint32_t l = m; l < degree; ++l) {
			float p, dp;
			if (l == m) {
				p = pmm;
				dp = 0.0f;
			} else if (l == m + 1) {
				p = (2.0f * (float)m + 1.0f) * z * p_lm1;
				dp = (2.0f * (float)m + 1.0f) * p_lm1;
			} else {
				const float a = 2.0f * (float)l - 1.0f, b = (float)(l + m) - 1.0f, inv = 1.0f / (float)(l - m);
				p = (a * z * p_lm1 - b * p_lm2) * inv;
				dp = (a * (p_lm1 + z * dp_lm1) - b * dp_lm2) * inv;
			}
			p_lm2 = p_lm1;
			dp_lm2 = dp_lm1;
			p_lm1 = p;
			dp_lm1 = dp;
			// K_l^m: (l - m)! / (l + m)! = 1 / prod_{k = l - m + 1}^{l + m} k
			float ratio = 1.0f;
			for (uint32_t k = l - m + 1; k <= l + m; ++k) ratio /= (float)k;
			float K = sqrtf((2.0f * (float)l + 1.0f) * ratio * (1.0f / (4.0f * PI_F)));
			const uint32_t base = l * (l + 1);
			if (m == 0) {
				emit(base, K * p, 0.0f, 0.0f, GRAD ? K * dp : 0.0f);
			} else {
				K *= 1.41421356237309504880f;
				emit(base + m, K * p * c, GRAD ? K * p * dc_dx : 0.0f, GRAD ? K * p * dc_dy : 0.0f, GRAD ? K * dp * c : 0.0f);
				emit(base - m, K * p * s, GRAD ? K * p * ds_dx : 0.0f, GRAD ? K * p * ds_dy : 0.0f, GRAD ? K * dp * s : 0.0f);
			}
		}
	}
}

__global__ void feature_forward_kernel(const FeatureSegments segs, uint32_t n, const float* __restrict__ x, uint32_t x_stride, __half* __restrict__ encoded, uint32_t row_stride) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const FeatureSegment sg = segs.s[blockIdx.y];
	if (sg.type == FEAT_GRID) return;
	const float* __restrict__ xin = x + (size_t)i * x_stride + sg.in_begin;
	__half* __restrict__ out = encoded + (size_t)i * row_stride + sg.out_begin;
	const __half one = __float2half_rn(1.0f);
	if (sg.type == FEAT_SPHERICAL_HARMONICS) {
		for (uint32_t j = 0; j < sg.n_pad; ++j) out[j] = one;  // padding FIRST (spherical_harmonics.h:56-60)
		__half* __restrict__ sh = out + sg.n_pad;
		const float dx = __ldg(xin) * 2.0f - 1.0f, dy = __ldg(xin + 1) * 2.0f - 1.0f, dz = __ldg(xin + 2) * 2.0f - 1.0f;
		spherical_harmonics<false>(sg.param, dx, dy, dz, [&](uint32_t idx, float v, float, float, float) { sh[idx] = __float2half_rn(v); });
		return;
	}
	if (sg.type == FEAT_IDENTITY) {
		for (uint32_t d = 0; d < sg.n_in; ++d) out[d] = __float2half_rn(__fmaf_rn(__ldg(xin + d), sg.scale, sg.offset));
	} else if (sg.type == FEAT_FREQUENCY) {
		const uint32_t F = sg.param;
		for (uint32_t d = 0; d < sg.n_in; ++d) {
			const float xd = __ldg(xin + d);
			for (uint32_t f = 0; f < F; ++f) {
				const float arg = scalbnf(xd, (int)f) * PI_F;
				out[d * 2 * F + 2 * f] = __float2half_rn(__sinf(arg));
				out[d * 2 * F + 2 * f + 1] = __float2half_rn(__sinf(arg + PI_F / 2));
			}
		}
	} else if (sg.type == FEAT_TRIANGLE_WAVE) {
		const uint32_t F = sg.param;
		for (uint32_t d = 0; d < sg.n_in; ++d) {
			const float xd = __ldg(xin + d);
			for (uint32_t f = 0; f < F; ++f) {
				const float val = scalbnf(xd, (int)f - 1) + (float)f * 0.25f;
				out[d * F + f] = __float2half_rn(fabsf(val - floorf(val) - 0.5f) * 4.0f - 1.0f);
			}
		}
	} else if (sg.type == FEAT_ONEBLOB) {
		const uint32_t log2_bins = sg.param, n_bins = 1u << log2_bins;
		for (uint32_t d = 0; d < sg.n_in; ++d) {
			const float xd = __ldg(xin + d);
			float left = wrapped_cdf(0.0f, xd, (float)n_bins);
			const float first = left;
			for (uint32_t b = 0; b < n_bins; ++b) {
				// the right boundary of the last bin wraps to the first bin's left boundary plus one whole kernel (oneblob.h:60-66)
				const float right = b + 1 == n_bins ? first + 1.0f : wrapped_cdf(scalbnf((float)(b + 1), -(int)log2_bins), xd, (float)n_bins);
				out[d * n_bins + b] = __float2half_rn(right - left);
				left = right;
			}
		}
	}
	for (uint32_t j = 0; j < sg.n_pad; ++j) out[sg.n_out + j] = one;
}

__global__ void feature_input_gradient_kernel(const FeatureSegments segs, uint32_t n, const float* __restrict__ x, uint32_t x_stride, const __half* __restrict__ dL_dy, uint32_t row_stride,
                                              float* __restrict__ dL_dx) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const FeatureSegment sg = segs.s[blockIdx.y];
	if (sg.type == FEAT_GRID) return;
	const float* __restrict__ xin = x + (size_t)i * x_stride + sg.in_begin;
	const __half* __restrict__ dy = dL_dy + (size_t)i * row_stride + sg.out_begin;
	float* __restrict__ dx = dL_dx + (size_t)i * x_stride + sg.in_begin;
	if (sg.type == FEAT_SPHERICAL_HARMONICS) {
		const __half* __restrict__ g = dy + sg.n_pad;
		const float vx = __ldg(xin) * 2.0f - 1.0f, vy = __ldg(xin + 1) * 2.0f - 1.0f, vz = __ldg(xin + 2) * 2.0f - 1.0f;
		float gx = 0.0f, gy = 0.0f, gz = 0.0f;
		spherical_harmonics<true>(sg.param, vx, vy, vz, [&](uint32_t idx, float, float ddx, float ddy, float ddz) {
			const float w = __half2float(g[idx]);
			gx += w * ddx;
			gy += w * ddy;
			gz += w * ddz;
		});
		dx[0] = 2.0f * gx;  // [0, 1]^3 -> [-1, 1]^3 (spherical_harmonics.h:95-98)
		dx[1] = 2.0f * gy;
		dx[2] = 2.0f * gz;
	} else if (sg.type == FEAT_IDENTITY) {
		for (uint32_t d = 0; d < sg.n_in; ++d) dx[d] = __half2float(dy[d]) * sg.scale;
	} else if (sg.type == FEAT_FREQUENCY) {
		const uint32_t F = sg.param;
		for (uint32_t d = 0; d < sg.n_in; ++d) {
			const float xd = __ldg(xin + d);
			float acc = 0.0f;
			for (uint32_t f = 0; f < F; ++f) {
				const float arg = scalbnf(xd, (int)f) * PI_F, k = scalbnf(1.0f, (int)f) * PI_F;
				acc += __half2float(dy[d * 2 * F + 2 * f]) * (k * __cosf(arg));
				acc += __half2float(dy[d * 2 * F + 2 * f + 1]) * (k * __cosf(arg + PI_F / 2));
			}
			dx[d] = acc;
		}
	} else if (sg.type == FEAT_TRIANGLE_WAVE) {
		const uint32_t F = sg.param;
		for (uint32_t d = 0; d < sg.n_in; ++d) {
			const float xd = __ldg(xin + d);
			float acc = 0.0f;
			for (uint32_t f = 0; f < F; ++f) {
				const float val = scalbnf(xd, (int)f - 1) + (float)f * 0.25f;
				const float slope = scalbnf(((int)floorf(val * 2.0f) % 2 == 0) ? -1.0f : 1.0f, (int)f + 1);  // triangle_wave.h:80
				acc += __half2float(dy[d * F + f]) * slope;
			}
			dx[d] = acc;
		}
	} else if (sg.type == FEAT_ONEBLOB) {
		const uint32_t log2_bins = sg.param, n_bins = 1u << log2_bins;
		for (uint32_t d = 0; d < sg.n_in; ++d) {
			const float xd = __ldg(xin + d);
			float left = wrapped_cdf_deriv(0.0f, xd, (float)n_bins);
			float acc = 0.0f;
			for (uint32_t b = 0; b < n_bins; ++b) {
				const float right = wrapped_cdf_deriv(scalbnf((float)(b + 1), -(int)log2_bins), xd, (float)n_bins);
				acc += __half2float(dy[d * n_bins + b]) * (left - right);  // d/dx of cdf(b - x) is -pdf: oneblob.h:139-148
				left = right;
			}
			dx[d] = acc;
		}
	}
}

}  // namespace

cudaError_t launch_feature_forward(cudaStream_t stream, const FeatureSegments& segs, uint32_t n, const float* x, uint32_t x_stride, __half* encoded, uint32_t row_stride) {
	if (segs.n == 0 || n == 0) return cudaSuccess;
	if (segs.n > MAX_FEATURE_SEGMENTS || !x || !encoded) return cudaErrorInvalidValue;
	feature_forward_kernel<<<dim3((n + 127) / 128, segs.n), 128, 0, stream>>>(segs, n, x, x_stride, encoded, row_stride);
	return cudaGetLastError();
}

cudaError_t launch_feature_input_gradient(cudaStream_t stream, const FeatureSegments& segs, uint32_t n, const float* x, uint32_t x_stride, const __half* dL_dy, uint32_t row_stride,
                                          float* dL_dx) {
	if (segs.n == 0 || n == 0) return cudaSuccess;
	if (segs.n > MAX_FEATURE_SEGMENTS || !x || !dL_dy || !dL_dx) return cudaErrorInvalidValue;
	feature_input_gradient_kernel<<<dim3((n + 127) / 128, segs.n), 128, 0, stream>>>(segs, n, x, x_stride, dL_dy, row_stride, dL_dx);
	return cudaGetLastError();
}

}  // namespace tcnnb
