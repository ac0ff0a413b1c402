// mlp_fused.h -- launch interface of the stand-alone FullyFusedMLP forward kernel (mlp_fused.cu).
#pragma once
#include "common.cuh"

namespace tcnnb {

struct MlpForwardParams {
	// network (fully_fused_mlp.cu:635-672): W_0 [width][in_width], (n_hidden_layers - 1) x [width][width], W_out [out_width][width],
	// row-major fp16, contiguous in this order
	uint32_t width;              // 16 / 32 / 64 / 128
	uint32_t in_width;           // multiple of 16, <= max(width, 64) ... see mlp_forward_supported()
	uint32_t out_width;          // PADDED output width: multiple of 16, <= width rounded up to 64
	uint32_t n_hidden_layers;    // >= 1
	uint32_t activation;         // hidden activation (Activation enum)
	uint32_t output_activation;
	const __half* weights;
	// batch
	uint32_t batch_size;         // multiple of 128
	// input: exactly one of the two
	const __half* input_fp16;    // [batch][in_width] fp16 (Network<T>::inference_mixed_precision, column-major in_width x batch)
	const float* input_fp32;     // [batch][n_input_dims] fp32 through the Identity encoding (encodings/identity.h:46-67):
	uint32_t n_input_dims;       //   feature j < n_input_dims = (half)x_j, features n_input_dims .. in_width-1 = 1 (padding with ones)
	// output: either / both may be null
	__half* output_fp16;         // [batch][out_width]
	float* output_fp32;          // [batch][n_output_dims] (network->inference: trimmed + cast, object.h:214-282)
	uint32_t n_output_dims;
	// optional: post-activation hidden layers [n_hidden_layers][batch][width] fp16 (forward pass kept for a backward pass)
	__half* hidden_out;
	// profiling only (scripts/mlp_timeline.py): clock64 stamps [cta][slot + 1 (0 = issuer)][64 events][8], null in production
	long long* dbg_clock;
};

// Maximum number of weight matrices that stay resident in shared memory for a width (more layers stream through a ring).
uint32_t mlp_forward_resident_layers(uint32_t width);
bool mlp_forward_supported(const MlpForwardParams& p, const char** why);
cudaError_t launch_mlp_forward(const MlpForwardParams& p, uint32_t n_sms, cudaStream_t stream);

}  // namespace tcnnb
