// misc_kernels.h -- launch interface of misc_kernels.cu.
#pragma once
#include "common.cuh"

namespace tcnnb {

// Hyper-parameters of optimizers/adam.h:221-303 with the reference's defaults (adam.h:337-355).
struct AdamParams {
	float learning_rate = 1e-3f;
	float beta1 = 0.9f;
	float beta2 = 0.999f;
	float epsilon = 1e-8f;
	float l2_reg = 1e-8f;
	float relative_decay = 0.0f;
	float absolute_decay = 0.0f;
	float clipping_magnitude = 0.0f;
	float gradient_clipping_magnitude = 0.0f;
	float non_matrix_learning_rate_factor = 1.0f;
	float non_matrix_l2_reg = 0.0f;
	float lower_lr_bound = 0.0f;          // AdaBound bounds, recomputed per step on the host (adam.h:161-168)
	float upper_lr_bound = 3.402823466e+38f;
	int adabound = 0;
	int optimize_matrix_params = 1;
	int optimize_non_matrix_params = 1;
	int skip_zero_grad_non_matrix_params = 1;
};

cudaError_t launch_random_uniform(cudaStream_t stream, Pcg32 rng, uint64_t n_elements, float* out, float lower, float upper);
cudaError_t launch_cast_params(cudaStream_t stream, uint64_t n, const float* in, __half* out);
cudaError_t launch_level_scales(cudaStream_t stream, uint32_t n_levels, float log2_per_level_scale, uint32_t base_resolution, float* scales_dev);
cudaError_t launch_adam_step(cudaStream_t stream, const AdamParams& a, uint32_t n_elements, uint32_t n_matrix_weights, float loss_scale,
                             float* weights_full_precision, __half* weights, __half* gradients, float* dw_accum, float* first_moments,
                             float* second_moments, uint32_t* param_steps);
cudaError_t launch_mlp_grad_finalize(cudaStream_t stream, uint32_t n, float* dw_accum, __half* gradients);

}  // namespace tcnnb
