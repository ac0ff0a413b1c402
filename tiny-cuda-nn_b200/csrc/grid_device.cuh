// grid_device.cuh -- per-sample multiresolution grid arithmetic shared by the standalone encoding kernels and the
// fused training kernel. Integer results (cells, indices) are bit-exact with the reference:
//   pos_fract            common_device.h:1031-1043
//   grid_index           common_device.h:847-884 (dense stride accumulation, coherent-prime hash, modulo)
//   coherent_prime_hash  common_device.h:787-791
//   N-linear weights     grid.h:142-163 (fp32 products in corner order, bit d of the corner index <-> dim d)
#pragma once
#include "common.cuh"

namespace tcnnb {

enum : uint32_t { LEVEL_DENSE = 0, LEVEL_HASH = 1, LEVEL_DEGENERATE = 2 };

__device__ __forceinline__ float smoothstep_f(float x) { return x * x * (3.0f - 2.0f * x); }

template <uint32_t D>
struct CellPos {
	uint32_t cell[D];
	float frac[D];
};

template <uint32_t D>
__device__ __forceinline__ void pos_fract(const float (&x)[D], float scale, uint32_t interpolation, CellPos<D>& out) {
#pragma unroll
	for (uint32_t d = 0; d < D; ++d) {
		float p = __fmaf_rn(scale, x[d], 0.5f);
		const float t = floorf(p);
		out.cell[d] = (uint32_t)(int)t;
		p -= t;
		out.frac[d] = interpolation == INTERP_SMOOTHSTEP ? smoothstep_f(p) : p;
	}
}

// Entry index of one corner inside its level.
template <uint32_t D>
__device__ __forceinline__ uint32_t corner_index(const LevelInfo& lv, const uint32_t (&c)[D]) {
	uint32_t index = 0;
	if (lv.use_hash == LEVEL_HASH) {
		index = c[0];
		if (D > 1) index ^= c[1] * 2654435761u;
		if (D > 2) index ^= c[2] * 805459861u;
		if (D > 3) index ^= c[3] * 3674653429u;
	} else if (lv.use_hash == LEVEL_DENSE) {
		uint32_t stride = 1;
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) {
			index += c[d] * stride;
			stride *= lv.resolution;
		}
	}
	return lv.pow2_mask ? (index & lv.pow2_mask) : (index % lv.size);
}

// Corner `idx` of the cell: coordinates and fp32 interpolation weight, multiplied in dimension order like the reference.
template <uint32_t D>
__device__ __forceinline__ float corner(const CellPos<D>& p, uint32_t idx, uint32_t (&c)[D]) {
	float w = 1.0f;
#pragma unroll
	for (uint32_t d = 0; d < D; ++d) {
		if ((idx & (1u << d)) == 0) {
			w *= 1.0f - p.frac[d];
			c[d] = p.cell[d];
		} else {
			w *= p.frac[d];
			c[d] = p.cell[d] + 1;
		}
	}
	return w;
}

}  // namespace tcnnb
