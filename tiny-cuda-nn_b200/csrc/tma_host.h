// tma_host.h -- host-side construction of TMA descriptors (CUtensorMap) without linking libcuda: the encoder is fetched through
// cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>  // CUtensorMap (types only)
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace tcnnb {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
	static EncodeTiledFn fn = [] {
		void* f = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
		return (EncodeTiledFn)f;
	}();
	return fn;
}

// fp16 matrix [rows][cols] row-major -> 2-D tensor map with a box of 64 columns x box_rows rows, SWIZZLE_128B, zero fill outside
// the matrix (a box wider than the matrix still occupies box_rows x 128 bytes of shared memory).
inline bool make_fp16_matrix_map(CUtensorMap* map, const __half* base, uint64_t rows, uint32_t cols, uint32_t box_rows) {
	EncodeTiledFn fn = encode_tiled_fn();
	if (!fn) return false;
	const cuuint64_t dims[2] = {cols, rows};
	const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(__half)};
	const cuuint32_t box[2] = {64, box_rows};
	const cuuint32_t elem[2] = {1, 1};
	return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tcnnb
