// grid_kernels.h -- launch interface of the stand-alone multiresolution-grid kernels (grid_kernels.cu): the encoding on its own
// (tcnn::cpp::create_encoding, cpp_api.h:124) and the input-position gradient of the module tier.
#pragma once
#include "common.cuh"

namespace tcnnb {

struct GridKernelArgs {
	uint32_t n_pos_dims;             // 2, 3, 4
	uint32_t n_features_per_level;   // 1, 2, 4, 8
	uint32_t n_levels;
	uint32_t interpolation;          // InterpolationType
	float max_level;                 // fraction of the levels that is active, as GridEncoding::set_max_level (grid.h:69-92); 1 = all
	const LevelInfo* levels_dev;     // [n_levels] in device memory
	uint32_t n_elements;
	const float* positions;          // [n][D] fp32
	uint32_t row_stride;             // fp16 elements per row of encoded / dL_dy (>= n_levels * F)
};

// encoded [n][row_stride] fp16 (row = sample; columns level * F + f; columns beyond n_levels * F are zeroed)        grid.h:49-169
cudaError_t launch_grid_forward(cudaStream_t stream, const GridKernelArgs& a, const __half* table, __half* encoded);
// grad_table (fp16, n_params, accumulated INTO: the caller zeroes it) += scatter of dL_dy [n][row_stride] fp16       grid.h:215-320
// F == 1 accumulates in `tmp_fp32` (n_params floats, zeroed by the caller) and casts at the end, as the reference (grid.h:858-894).
cudaError_t launch_grid_backward(cudaStream_t stream, const GridKernelArgs& a, const __half* dL_dy, __half* grad_table, float* tmp_fp32, uint32_t n_params);
// dL_dx [n][D] fp32 = sum over features of dL_dy * d(encoded)/d(position)                                           grid.h:170-212,322-350
cudaError_t launch_grid_input_gradient(cudaStream_t stream, const GridKernelArgs& a, const __half* table, const __half* dL_dy, float* dL_dx);

}  // namespace tcnnb
