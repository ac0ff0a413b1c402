// misc_kernels.cu -- the streaming kernels around the fused step: parameter initialisation (pcg32 jump-ahead fill),
// fp32 -> fp16 parameter cast, per-level scale evaluation, and the Adam step.
//
// Compiled with --use_fast_math like the reference (CMakeLists.txt:288-290) so that exp2f / powf / sqrtf / '/' lower
// to the same approximate instructions as in the reference's kernels.
#include "common.cuh"
#include "misc_kernels.h"

namespace tcnnb {

// ------------------------------------------------------------------------------------------------------------------
// pcg32 (PCG-XSH-RR 64/32, O'Neill) on the device: next_uint / next_float / advance as specified in
// dependencies/pcg32/pcg32.h:62-69,103-112,145-166.
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr uint64_t PCG32_MULT = 0x5851f42d4c957f2dULL;

__device__ __forceinline__ uint32_t pcg_next_uint(Pcg32& r) {
	const uint64_t old = r.state;
	r.state = old * PCG32_MULT + r.inc;
	const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
	const uint32_t rot = (uint32_t)(old >> 59u);
	return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}

__device__ __forceinline__ float pcg_next_float(Pcg32& r) {
	return __uint_as_float((pcg_next_uint(r) >> 9) | 0x3f800000u) - 1.0f;
}

__device__ __forceinline__ void pcg_advance(Pcg32& r, uint64_t delta) {
	uint64_t cur_mult = PCG32_MULT, cur_plus = r.inc, acc_mult = 1u, acc_plus = 0u;
	while (delta > 0) {
		if (delta & 1) {
			acc_mult *= cur_mult;
			acc_plus = acc_plus * cur_mult + cur_plus;
		}
		cur_plus = (cur_mult + 1) * cur_plus;
		cur_mult *= cur_mult;
		delta /= 2;
	}
	r.state = acc_mult * r.state + acc_plus;
}

// generate_random_kernel (random.h:40-53): thread i jumps ahead 4i draws and writes elements i, i+n_thr, i+2n_thr, i+3n_thr.
__global__ void random_uniform_kernel(uint64_t n_elements, Pcg32 rng, float* __restrict__ out, float lower, float upper) {
	const uint64_t i = threadIdx.x + (uint64_t)blockIdx.x * blockDim.x;
	const uint64_t n_threads = (uint64_t)blockDim.x * gridDim.x;
	pcg_advance(rng, i * 4);
#pragma unroll
	for (uint64_t j = 0; j < 4; ++j) {
		const uint64_t idx = i + n_threads * j;
		if (idx >= n_elements) return;
		out[idx] = pcg_next_float(rng) * (upper - lower) + lower;  // contracted to one FMA, as in random.h:69
	}
}

__global__ void cast_params_kernel(uint64_t n, const float* __restrict__ in, __half* __restrict__ out) {
	const uint64_t i = threadIdx.x + (uint64_t)blockIdx.x * blockDim.x;
	if (i < n) out[i] = (__half)in[i];
}

// grid_scale evaluated on the device exactly as the reference's kernels do (common_device.h:886-891 called from
// grid.h:97 with log2_per_level_scale = std::log2(per_level_scale) computed on the host, grid.h:782).
__global__ void level_scales_kernel(uint32_t n_levels, float log2_per_level_scale, uint32_t base_resolution, float* __restrict__ scales) {
	const uint32_t level = threadIdx.x;
	if (level < n_levels) {
		scales[level] = exp2f(level * log2_per_level_scale) * base_resolution - 1.0f;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Adam (optimizers/adam.h:48-129). One thread per parameter, same expression order as the reference.
// Differences in data flow only:
//   * matrix (MLP) weight gradients arrive as fp32 sums in dw_accum (written by the fused kernel with red.add.f32);
//     they are rounded to fp16 here -- which is what the reference's gradient buffer holds -- stored to `gradients`
//     for API visibility, and the accumulator is re-armed to zero for the next step.
//   * optimizer state streams use evict-first loads/stores so that the fp16 table + gradient table stay L2 resident.
// ------------------------------------------------------------------------------------------------------------------
__global__ void adam_step_kernel(const AdamParams a, const uint32_t n_elements, const uint32_t n_matrix_weights, const float loss_scale,
                                 float* __restrict__ weights_full_precision, __half* __restrict__ weights, __half* __restrict__ gradients,
                                 float* __restrict__ dw_accum, float* __restrict__ first_moments, float* __restrict__ second_moments,
                                 uint32_t* __restrict__ param_steps) {
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	if (i >= n_elements) return;

	__half grad_h;
	if (i < n_matrix_weights && dw_accum != nullptr) {
		grad_h = (__half)dw_accum[i];
		dw_accum[i] = 0.0f;
		gradients[i] = grad_h;
	} else {
		grad_h = gradients[i];
	}

	float gradient = (float)grad_h / loss_scale;
	if (i >= n_matrix_weights) {
		if (!a.optimize_non_matrix_params || (gradient == 0 && a.skip_zero_grad_non_matrix_params)) return;
	} else {
		if (!a.optimize_matrix_params) return;
	}

	const float weight_fp = __ldcs(weights_full_precision + i);

	if (i < n_matrix_weights) {
		gradient += a.l2_reg * weight_fp;
	} else {
		gradient += a.non_matrix_l2_reg * weight_fp;
	}

	if (a.gradient_clipping_magnitude != 0.0f) {
		gradient = copysignf(fminf(fabsf(gradient), a.gradient_clipping_magnitude), gradient);
	}

	const float gradient_sq = gradient * gradient;

	const float first_moment = a.beta1 * __ldcs(first_moments + i) + (1 - a.beta1) * gradient;
	__stcs(first_moments + i, first_moment);
	const float second_moment = a.beta2 * __ldcs(second_moments + i) + (1 - a.beta2) * gradient_sq;
	__stcs(second_moments + i, second_moment);

	float learning_rate = a.learning_rate;
	if (i >= n_matrix_weights) learning_rate *= a.non_matrix_learning_rate_factor;

	const uint32_t current_step = __ldcs(param_steps + i) + 1;
	__stcs(param_steps + i, current_step);
	learning_rate *= sqrtf(1 - powf(a.beta2, (float)current_step)) / (1 - powf(a.beta1, (float)current_step));

	const float effective_learning_rate = fminf(fmaxf(learning_rate / (sqrtf(second_moment) + a.epsilon), a.lower_lr_bound), a.upper_lr_bound);

	// weight_decay (common_device.h:1045-1048)
	const float decayed_weight = (1 - a.relative_decay * learning_rate) * weight_fp - copysignf(a.absolute_decay * learning_rate, weight_fp);
	float new_weight = decayed_weight - effective_learning_rate * first_moment;

	if (a.clipping_magnitude != 0.0f) {
		new_weight = fminf(fmaxf(new_weight, -a.clipping_magnitude), a.clipping_magnitude);
	}

	__stcs(weights_full_precision + i, new_weight);
	weights[i] = (__half)new_weight;
}

__global__ void mlp_grad_finalize_kernel(uint32_t n, float* __restrict__ dw_accum, __half* __restrict__ gradients) {
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	if (i < n) {
		gradients[i] = (__half)dw_accum[i];
		dw_accum[i] = 0.0f;
	}
}

}  // namespace

static inline uint32_t blocks_for(uint64_t n, uint32_t threads) { return (uint32_t)((n + threads - 1) / threads); }

cudaError_t launch_random_uniform(cudaStream_t stream, Pcg32 rng, uint64_t n_elements, float* out, float lower, float upper) {
	if (n_elements == 0) return cudaSuccess;
	const uint64_t n_threads = (n_elements + 3) / 4;
	random_uniform_kernel<<<blocks_for(n_threads, 128), 128, 0, stream>>>(n_elements, rng, out, lower, upper);
	return cudaGetLastError();
}

cudaError_t launch_cast_params(cudaStream_t stream, uint64_t n, const float* in, __half* out) {
	if (n == 0) return cudaSuccess;
	cast_params_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(n, in, out);
	return cudaGetLastError();
}

cudaError_t launch_level_scales(cudaStream_t stream, uint32_t n_levels, float log2_per_level_scale, uint32_t base_resolution, float* scales_dev) {
	level_scales_kernel<<<1, 128, 0, stream>>>(n_levels, log2_per_level_scale, base_resolution, scales_dev);
	return cudaGetLastError();
}

cudaError_t launch_adam_step(cudaStream_t stream, const AdamParams& a, uint32_t n_elements, uint32_t n_matrix_weights, float loss_scale,
                             float* weights_full_precision, __half* weights, __half* gradients, float* dw_accum, float* first_moments,
                             float* second_moments, uint32_t* param_steps) {
	if (n_elements == 0) return cudaSuccess;
	adam_step_kernel<<<blocks_for(n_elements, 256), 256, 0, stream>>>(a, n_elements, n_matrix_weights, loss_scale, weights_full_precision, weights,
	                                                                gradients, dw_accum, first_moments, second_moments, param_steps);
	return cudaGetLastError();
}

cudaError_t launch_mlp_grad_finalize(cudaStream_t stream, uint32_t n, float* dw_accum, __half* gradients) {
	if (n == 0) return cudaSuccess;
	mlp_grad_finalize_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(n, dw_accum, gradients);
	return cudaGetLastError();
}

}  // namespace tcnnb
