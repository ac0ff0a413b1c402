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

// General (unfused) training path: loss + dL/d(output) from fp16 prediction rows [batch][stride]; loss_values / loss_sum may be null.
cudaError_t launch_loss(cudaStream_t stream, uint32_t loss_type, uint32_t output_activation, uint32_t batch, uint32_t n_out, uint32_t stride, float loss_scale, uint32_t n_total,
                        const __half* prediction, const float* targets, __half* dL_dy, float* loss_values, float* loss_sum);
// out = grad * f'(.) of the output activation, expressed through the forward output (element-wise, fp16).
cudaError_t launch_activation_backward_output(cudaStream_t stream, uint32_t activation, uint64_t n, const __half* grad, const __half* forward_output, __half* out);

// Ema wrapper (optimizers/ema.h:46-75): weights_ema = (weights_ema * decay * debias_old + weights * (1 - decay)) * debias_new; tmp optional (fp32 average)
cudaError_t launch_ema_step(cudaStream_t stream, uint32_t n, float decay, float debias_old, float debias_new, const __half* weights, __half* weights_ema, float* tmp);
// Identity encoding (encodings/identity.h:46-91) as stand-alone kernels: rows [n][width] fp16 with ones as padding; dL/dx in fp32.
cudaError_t launch_identity_encode(cudaStream_t stream, uint32_t n, uint32_t n_dims, uint32_t width, float scale, float offset, const float* x, __half* out);
cudaError_t launch_identity_backward(cudaStream_t stream, uint32_t n, uint32_t n_dims, uint32_t width, float scale, const __half* dL_dy, float* dL_dx);

// ---- data parallelism over peer memory (NVLink / NVSwitch): one-pass "reduce + Adam + publish" on this rank's slice ----------------
// Every rank holds [fp16 params | fp16 gradients | flags] at the SAME offsets of a symmetric allocation that all peers have
// mapped (rendezvous by the host framework). peers.* are this rank's views of every rank's copy; *_mc are the NVLS multicast
// views of the same regions (null when the fabric has no multicast: the kernels then loop over the peer pointers).
constexpr uint32_t DP_MAX_RANKS = 8;
struct DpPeers {
	uint32_t world, rank;
	__half* params[DP_MAX_RANKS];
	__half* grads[DP_MAX_RANKS];
	uint32_t* flags[DP_MAX_RANKS];  // [DP_MAX_RANKS] monotonic epoch counters per rank: flags[r][q] = last barrier rank q has reached, as seen by rank r
	__half* params_mc;
	__half* grads_mc;
};
// Cross-GPU barrier on `stream`: everything enqueued before it on every rank's stream is complete and visible to all ranks before
// anything enqueued after it starts. `epoch` must increase by one per call, identically on all ranks.
cudaError_t launch_dp_barrier(cudaStream_t stream, const DpPeers& peers, uint32_t epoch);
// Adam over the parameters [first, first + count) (multiples of 8) of the padded vector: the gradient of a parameter is the SUM over
// ranks of grads[r][i] (multimem.ld_reduce in the switch, or peer loads), evaluated once by the slice's owner; the updated fp16
// weight is published into every rank's params (multimem.st, or peer stores); the reduced gradient is left in the local gradient
// buffer. State arrays (fp32 master, moments, steps) are local and indexed by the global parameter index.
cudaError_t launch_adam_step_dp(cudaStream_t stream, const AdamParams& a, const DpPeers& peers, uint64_t first, uint64_t count, uint32_t n_matrix_weights, uint64_t n_params,
                                float loss_scale, float* weights_full_precision, float* first_moments, float* second_moments, uint32_t* param_steps);

}  // namespace tcnnb
