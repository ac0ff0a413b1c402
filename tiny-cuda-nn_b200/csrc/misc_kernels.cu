// misc_kernels.cu -- the streaming kernels around the fused step: parameter initialisation (pcg32 jump-ahead fill),
// fp32 -> fp16 parameter cast, per-level scale evaluation, and the Adam step.
//
// Compiled with --use_fast_math like the reference (CMakeLists.txt:288-290) so that exp2f / powf / sqrtf / '/' lower
// to the same approximate instructions as in the reference's kernels.
#include "common.cuh"
#include "misc_kernels.h"

#include "fused_common.cuh"

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
// One Adam update, shared by the scalar and the 4-wide kernels. Returns false if the parameter is skipped.
__device__ __forceinline__ bool adam_update(const AdamParams& a, const bool is_matrix, const float loss_scale, const __half grad_h,
                                            float& weight_fp, float& first_moment, float& second_moment, uint32_t& step) {
	float gradient = (float)grad_h / loss_scale;
	if (!is_matrix) {
		if (!a.optimize_non_matrix_params || (gradient == 0 && a.skip_zero_grad_non_matrix_params)) return false;
	} else {
		if (!a.optimize_matrix_params) return false;
	}
	if (is_matrix) {
		gradient += a.l2_reg * weight_fp;
	} else {
		gradient += a.non_matrix_l2_reg * weight_fp;
	}
	if (a.gradient_clipping_magnitude != 0.0f) {
		gradient = copysignf(fminf(fabsf(gradient), a.gradient_clipping_magnitude), gradient);
	}
	const float gradient_sq = gradient * gradient;
	first_moment = a.beta1 * first_moment + (1 - a.beta1) * gradient;
	second_moment = a.beta2 * second_moment + (1 - a.beta2) * gradient_sq;
	float learning_rate = a.learning_rate;
	if (!is_matrix) learning_rate *= a.non_matrix_learning_rate_factor;
	const uint32_t current_step = ++step;
	learning_rate *= sqrtf(1 - powf(a.beta2, (float)current_step)) / (1 - powf(a.beta1, (float)current_step));
	const float effective_learning_rate = fminf(fmaxf(learning_rate / (sqrtf(second_moment) + a.epsilon), a.lower_lr_bound), a.upper_lr_bound);
	const float decayed_weight = (1 - a.relative_decay * learning_rate) * weight_fp - copysignf(a.absolute_decay * learning_rate, weight_fp);
	float new_weight = decayed_weight - effective_learning_rate * first_moment;
	if (a.clipping_magnitude != 0.0f) {
		new_weight = fminf(fmaxf(new_weight, -a.clipping_magnitude), a.clipping_magnitude);
	}
	weight_fp = new_weight;
	return true;
}

// One parameter per thread: the tail / unaligned path (caller-provided arrays that are not 16-byte aligned, lengths that are not a
// multiple of four); same update as the 4-wide kernel.
__global__ void adam_step_kernel(const AdamParams a, const uint32_t n_elements, const uint32_t n_matrix_weights, const float loss_scale,
                                 float* __restrict__ weights_full_precision, __half* __restrict__ weights, __half* __restrict__ gradients,
                                 float* __restrict__ dw_accum, float* __restrict__ first_moments, float* __restrict__ second_moments,
                                 uint32_t* __restrict__ param_steps) {
	pdl_wait();  // no early pdl_launch_dependents(): a long bandwidth-bound kernel should not share its SMs with parked CTAs
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
	float w = __ldcs(weights_full_precision + i), m = __ldcs(first_moments + i), v = __ldcs(second_moments + i);
	uint32_t st = __ldcs(param_steps + i);
	if (!adam_update(a, i < n_matrix_weights, loss_scale, grad_h, w, m, v, st)) return;
	__stcs(weights_full_precision + i, w);
	__stcs(first_moments + i, m);
	__stcs(second_moments + i, v);
	__stcs(param_steps + i, st);
	weights[i] = (__half)w;
}

// 4 parameters per thread: 128-bit streaming accesses to the fp32 state, 64-bit to the fp16 params / gradients.
// Groups whose four gradients are all zero (untouched hash-table entries) touch nothing but the 8 gradient bytes.
__global__ void __launch_bounds__(256) adam_step_vec4_kernel(const AdamParams a, const uint32_t n_groups, const uint32_t n_matrix_weights, const float loss_scale,
                                      float4* __restrict__ weights_full_precision, uint2* __restrict__ weights, uint2* __restrict__ gradients,
                                      float4* __restrict__ dw_accum, float4* __restrict__ first_moments, float4* __restrict__ second_moments,
                                      uint4* __restrict__ param_steps) {
	pdl_wait();
	const uint32_t gidx = threadIdx.x + blockIdx.x * blockDim.x;
	if (gidx >= n_groups) return;
	const uint32_t i0 = gidx * 4;

	// Issue every load of this group up front (independent 128-bit streams); the zero-gradient early-out below only
	// saves the write-back -- untouched groups are ~2 % of a step, memory-level parallelism matters more.
	const bool from_accum = i0 + 3 < n_matrix_weights && dw_accum != nullptr;
	const uint2 raw = from_accum ? make_uint2(0, 0) : gradients[gidx];
	const float4 w4 = __ldcs(weights_full_precision + gidx);
	const float4 m4 = __ldcs(first_moments + gidx);
	const float4 v4 = __ldcs(second_moments + gidx);
	const uint4 s4 = __ldcs(param_steps + gidx);

	__half g[4];
	if (from_accum) {
		const float4 acc = dw_accum[gidx];
		dw_accum[gidx] = make_float4(0.f, 0.f, 0.f, 0.f);
		g[0] = (__half)acc.x; g[1] = (__half)acc.y; g[2] = (__half)acc.z; g[3] = (__half)acc.w;
		gradients[gidx] = *reinterpret_cast<const uint2*>(g);
	} else {
		*reinterpret_cast<uint2*>(g) = raw;
		if (i0 >= n_matrix_weights && a.skip_zero_grad_non_matrix_params && ((raw.x | raw.y) & 0x7FFF7FFFu) == 0) return;  // all four are +-0
	}

	float w[4] = {w4.x, w4.y, w4.z, w4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
	uint32_t st[4] = {s4.x, s4.y, s4.z, s4.w};
	bool any = false;
#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) {
		any |= adam_update(a, i0 + k < n_matrix_weights, loss_scale, g[k], w[k], m[k], v[k], st[k]);
	}
	if (!any) return;
	__stcs(weights_full_precision + gidx, make_float4(w[0], w[1], w[2], w[3]));
	__stcs(first_moments + gidx, make_float4(m[0], m[1], m[2], m[3]));
	__stcs(second_moments + gidx, make_float4(v[0], v[1], v[2], v[3]));
	__stcs(param_steps + gidx, make_uint4(st[0], st[1], st[2], st[3]));
	__half h[4] = {(__half)w[0], (__half)w[1], (__half)w[2], (__half)w[3]};
	weights[gidx] = *reinterpret_cast<const uint2*>(h);
}

__global__ void mlp_grad_finalize_kernel(uint32_t n, float* __restrict__ dw_accum, __half* __restrict__ gradients) {
	pdl_wait();
	pdl_launch_dependents();
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	if (i < n) {
		gradients[i] = (__half)dw_accum[i];
		dw_accum[i] = 0.0f;
	}
}

// ---- loss on its own (general, unfused training path): prediction rows [batch][stride] fp16 + targets [batch][n_out] fp32 ->
// dL/d(output) rows (x loss_scale, through the output activation's transfer, padding columns zero), per-element values, sum.
// Same expressions as the fused kernel's output epilogue (fused_common.cuh loss_element).
__global__ void loss_kernel(uint32_t loss_type, uint32_t out_act, uint32_t batch, uint32_t n_out, uint32_t stride, float loss_scale, float n_total, const __half* __restrict__ prediction,
                            const float* __restrict__ targets, __half* __restrict__ dL_dy, float* __restrict__ loss_values, float* __restrict__ loss_sum) {
	pdl_wait();
	pdl_launch_dependents();
	// one thread per (sample, group of 8 output columns): 16-byte loads / stores of the fp16 rows
	const uint32_t groups = stride / 8;
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	float value_sum = 0.0f;
	if (i < batch * groups) {
		const uint32_t sample = i / groups, q0 = (i % groups) * 8;
		const uint4 raw = *reinterpret_cast<const uint4*>(prediction + (size_t)sample * stride + q0);
		const __half* y = reinterpret_cast<const __half*>(&raw);
		__half dy[8];
		const float luminance = loss_type == LOSS_RELATIVE_L2_LUMINANCE ? fused::row_luminance(prediction + (size_t)sample * stride, n_out) : 0.0f;
#pragma unroll
		for (uint32_t k = 0; k < 8; ++k) {
			const uint32_t q = q0 + k;
			float gq = 0.0f;
			if (q < n_out) {
				float value, grad;
				fused::loss_element(loss_type, __half2float(y[k]), targets[(size_t)sample * n_out + q], n_total, luminance, value, grad);
				gq = loss_scale * grad / n_total;
				value_sum += value;
				if (loss_values) loss_values[(size_t)sample * n_out + q] = value;
			}
			dy[k] = fused::act_bwd_h(out_act, __float2half_rn(gq), y[k]);
		}
		*reinterpret_cast<uint4*>(dL_dy + (size_t)sample * stride + q0) = *reinterpret_cast<const uint4*>(dy);
	}
	// block sum -> ONE atomic per block (a per-warp atomic on the single loss word serialises the whole kernel)
	__shared__ float warp_sums[8];
#pragma unroll
	for (uint32_t o = 16; o > 0; o >>= 1) value_sum += __shfl_xor_sync(0xFFFFFFFFu, value_sum, o);
	if ((threadIdx.x & 31u) == 0) warp_sums[threadIdx.x >> 5] = value_sum;
	__syncthreads();
	if (threadIdx.x == 0 && loss_sum) {
		float total = 0.0f;
		for (uint32_t w = 0; w < blockDim.x / 32; ++w) total += warp_sums[w];
		if (total != 0.0f) atomicAdd(loss_sum, total);
	}
}

// dL/d(output) of a caller (Network::backward, fully_fused_mlp.cu:755-762) through the output activation's transfer.
__global__ void activation_backward_output_kernel(uint32_t act, uint64_t n, const __half* __restrict__ grad, const __half* __restrict__ fwd, __half* __restrict__ out) {
	pdl_wait();
	pdl_launch_dependents();
	const uint64_t i = threadIdx.x + (uint64_t)blockIdx.x * blockDim.x;
	if (i < n) out[i] = fused::act_bwd_h(act, grad[i], fwd[i]);
}

// optimizers/ema.h:46-75: debiased exponential moving average of the working weights; `tmp` (fp32 copy of the average) when the
// wrapper runs in full precision.
__global__ void ema_step_kernel(uint32_t n, float decay, float debias_old, float debias_new, const __half* __restrict__ weights, __half* __restrict__ weights_ema, float* __restrict__ tmp) {
	pdl_wait();
	pdl_launch_dependents();
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	if (i >= n) return;
	const float prev = tmp ? tmp[i] : __half2float(weights_ema[i]);
	const float filtered = (prev * decay * debias_old + __half2float(weights[i]) * (1.0f - decay)) * debias_new;
	if (tmp) tmp[i] = filtered;
	weights_ema[i] = __float2half_rn(filtered);
}

// Identity encoding on its own (encodings/identity.h:46-67,69-91): rows [n][width] fp16, feature j < n_dims = x_j * scale + offset,
// padding features 1; and its backward, dL/dx_j = dL/d(feature j) * scale (fp32 rows [n][n_dims]).
__global__ void identity_encode_kernel(uint64_t n_total, uint32_t n_dims, uint32_t width, float scale, float offset, const float* __restrict__ x, __half* __restrict__ out) {
	pdl_wait();
	pdl_launch_dependents();
	const uint64_t i = threadIdx.x + (uint64_t)blockIdx.x * blockDim.x;
	if (i >= n_total) return;
	const uint64_t sample = i / width;
	const uint32_t j = (uint32_t)(i % width);
	out[i] = j < n_dims ? __float2half_rn(fmaf(x[sample * n_dims + j], scale, offset)) : __float2half_rn(1.0f);
}

__global__ void identity_backward_kernel(uint64_t n_total, uint32_t n_dims, uint32_t width, float scale, const __half* __restrict__ dL_dy, float* __restrict__ dL_dx) {
	pdl_wait();
	pdl_launch_dependents();
	const uint64_t i = threadIdx.x + (uint64_t)blockIdx.x * blockDim.x;
	if (i >= n_total) return;
	const uint64_t sample = i / n_dims;
	const uint32_t j = (uint32_t)(i % n_dims);
	dL_dx[i] = __half2float(dL_dy[sample * width + j]) * scale;
}

// ------------------------------------------------------------------------------------------------------------------
// Data parallelism over peer memory. See misc_kernels.h for the protocol.
// ------------------------------------------------------------------------------------------------------------------
__global__ void dp_barrier_kernel(const DpPeers peers, const uint32_t epoch) {
	const uint32_t t = threadIdx.x;
	if (t >= peers.world) return;
	__threadfence_system();  // this rank's earlier writes (gradients, published weights) before the signal
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.flags[t] + peers.rank), "r"(epoch) : "memory");
	const uint32_t* mine = peers.flags[peers.rank] + t;
	const long long t0 = clock64();
	for (;;) {
		uint32_t seen;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
		if ((int32_t)(seen - epoch) >= 0) break;
		if (clock64() - t0 > 40000000000LL) __trap();  // ~20 s: a peer died; fail loudly instead of hanging the GPU
	}
}

template <bool MULTICAST>
__global__ void __launch_bounds__(256) adam_step_dp_kernel(const AdamParams a, const DpPeers peers, const uint64_t first, const uint32_t n_groups, const uint32_t n_matrix_weights,
                                                            const uint64_t n_params, const float loss_scale, float* __restrict__ weights_full_precision,
                                                            float* __restrict__ first_moments, float* __restrict__ second_moments, uint32_t* __restrict__ param_steps) {
	const uint32_t gidx = threadIdx.x + blockIdx.x * blockDim.x;
	if (gidx >= n_groups) return;
	const uint64_t i0 = first + (uint64_t)gidx * 8;
	if (i0 >= n_params) return;  // padding of the parameter vector: never touched

	// ---- the reduced gradient of 8 consecutive parameters: one 16-byte access per rank, or one in-switch reduction
	uint32_t gw[4];
	if (MULTICAST) {
		asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];" : "=r"(gw[0]), "=r"(gw[1]), "=r"(gw[2]), "=r"(gw[3]) : "l"(peers.grads_mc + i0) : "memory");
	} else {
		float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (uint32_t r = 0; r < peers.world; ++r) {
			uint32_t w[4];
			asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(peers.grads[r] + i0) : "memory");
#pragma unroll
			for (uint32_t k = 0; k < 4; ++k) {
				const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
				acc[2 * k] += f.x;
				acc[2 * k + 1] += f.y;
			}
		}
#pragma unroll
		for (uint32_t k = 0; k < 4; ++k) {
			const __half2 h = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
			gw[k] = *reinterpret_cast<const uint32_t*>(&h);
		}
	}
	// the reduced gradient stays in the local gradient buffer (what trainer->param_gradients() shows for the owned slice)
	*reinterpret_cast<uint4*>(peers.grads[peers.rank] + i0) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
	if (i0 >= n_matrix_weights && a.skip_zero_grad_non_matrix_params && ((gw[0] | gw[1] | gw[2] | gw[3]) & 0x7FFF7FFFu) == 0) return;  // untouched entries (adam.h:79-82)

	const float4* w4p = reinterpret_cast<const float4*>(weights_full_precision + i0);
	const float4* m4p = reinterpret_cast<const float4*>(first_moments + i0);
	const float4* v4p = reinterpret_cast<const float4*>(second_moments + i0);
	const uint4* s4p = reinterpret_cast<const uint4*>(param_steps + i0);
	const float4 wa = __ldcs(w4p), wb = __ldcs(w4p + 1), ma = __ldcs(m4p), mb = __ldcs(m4p + 1), va = __ldcs(v4p), vb = __ldcs(v4p + 1);
	const uint4 sa = __ldcs(s4p), sb = __ldcs(s4p + 1);
	float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w}, m[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w}, v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
	uint32_t st[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
	const __half* g = reinterpret_cast<const __half*>(gw);
	bool any = false;
#pragma unroll
	for (uint32_t k = 0; k < 8; ++k) any |= adam_update(a, i0 + k < n_matrix_weights, loss_scale, g[k], w[k], m[k], v[k], st[k]);
	if (!any) return;
	float4* w4o = reinterpret_cast<float4*>(weights_full_precision + i0);
	float4* m4o = reinterpret_cast<float4*>(first_moments + i0);
	float4* v4o = reinterpret_cast<float4*>(second_moments + i0);
	uint4* s4o = reinterpret_cast<uint4*>(param_steps + i0);
	__stcs(w4o, make_float4(w[0], w[1], w[2], w[3]));
	__stcs(w4o + 1, make_float4(w[4], w[5], w[6], w[7]));
	__stcs(m4o, make_float4(m[0], m[1], m[2], m[3]));
	__stcs(m4o + 1, make_float4(m[4], m[5], m[6], m[7]));
	__stcs(v4o, make_float4(v[0], v[1], v[2], v[3]));
	__stcs(v4o + 1, make_float4(v[4], v[5], v[6], v[7]));
	__stcs(s4o, make_uint4(st[0], st[1], st[2], st[3]));
	__stcs(s4o + 1, make_uint4(st[4], st[5], st[6], st[7]));
	// ---- publish the working-precision weights to every replica
	uint32_t hw[4];
#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) {
		const __half2 h = __floats2half2_rn(w[2 * k], w[2 * k + 1]);
		hw[k] = *reinterpret_cast<const uint32_t*>(&h);
	}
	if (MULTICAST) {
		asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(peers.params_mc + i0), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
	} else {
		for (uint32_t r = 0; r < peers.world; ++r) {
			asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(peers.params[r] + i0), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
		}
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
	auto aligned = [](const void* p, size_t a) { return ((uintptr_t)p % a) == 0; };
	// A skipped lane of a partially updated group is written back with its loaded (unchanged) value, which is only
	// equivalent to "not touched" if the weights' fp16 copy equals (half)fp32 master -- true for trainer-owned buffers.
	const bool vec_ok = n_elements % 4 == 0 && n_matrix_weights % 4 == 0 && aligned(weights_full_precision, 16) && aligned(weights, 8) && aligned(gradients, 8) &&
	                    aligned(first_moments, 16) && aligned(second_moments, 16) && aligned(param_steps, 16) && (dw_accum == nullptr || aligned(dw_accum, 16));
	if (vec_ok) {
		const uint32_t n_groups = n_elements / 4;
		return launch_pdl(adam_step_vec4_kernel, blocks_for(n_groups, 256), 256, 0, stream, a, n_groups, n_matrix_weights, loss_scale, (float4*)weights_full_precision,
		                  (uint2*)weights, (uint2*)gradients, (float4*)dw_accum, (float4*)first_moments, (float4*)second_moments, (uint4*)param_steps);
	} else {
		return launch_pdl(adam_step_kernel, blocks_for(n_elements, 256), 256, 0, stream, a, n_elements, n_matrix_weights, loss_scale, weights_full_precision, weights,
		                  gradients, dw_accum, first_moments, second_moments, param_steps);
	}
	return cudaGetLastError();
}

cudaError_t launch_mlp_grad_finalize(cudaStream_t stream, uint32_t n, float* dw_accum, __half* gradients) {
	if (n == 0) return cudaSuccess;
	return launch_pdl(mlp_grad_finalize_kernel, blocks_for(n, 256), 256, 0, stream, n, dw_accum, gradients);
}

cudaError_t launch_loss(cudaStream_t stream, uint32_t loss_type, uint32_t output_activation, uint32_t batch, uint32_t n_out, uint32_t stride, float loss_scale, uint32_t n_total,
                        const __half* prediction, const float* targets, __half* dL_dy, float* loss_values, float* loss_sum) {
	if ((uint64_t)batch * stride >= (1ull << 32) || stride % 8 != 0) return cudaErrorInvalidValue;
	return launch_pdl(loss_kernel, blocks_for(batch * (stride / 8), 256), 256, 0, stream, loss_type, output_activation, batch, n_out, stride, loss_scale, (float)n_total, prediction, targets,
	                  dL_dy, loss_values, loss_sum);
}

cudaError_t launch_activation_backward_output(cudaStream_t stream, uint32_t activation, uint64_t n, const __half* grad, const __half* forward_output, __half* out) {
	if (n == 0) return cudaSuccess;
	return launch_pdl(activation_backward_output_kernel, (uint32_t)((n + 255) / 256), 256, 0, stream, activation, n, grad, forward_output, out);
}

cudaError_t launch_ema_step(cudaStream_t stream, uint32_t n, float decay, float debias_old, float debias_new, const __half* weights, __half* weights_ema, float* tmp) {
	if (n == 0) return cudaSuccess;
	return launch_pdl(ema_step_kernel, blocks_for(n, 256), 256, 0, stream, n, decay, debias_old, debias_new, weights, weights_ema, tmp);
}

cudaError_t launch_identity_encode(cudaStream_t stream, uint32_t n, uint32_t n_dims, uint32_t width, float scale, float offset, const float* x, __half* out) {
	const uint64_t total = (uint64_t)n * width;
	if (total == 0) return cudaSuccess;
	return launch_pdl(identity_encode_kernel, (uint32_t)((total + 255) / 256), 256, 0, stream, total, n_dims, width, scale, offset, x, out);
}

cudaError_t launch_identity_backward(cudaStream_t stream, uint32_t n, uint32_t n_dims, uint32_t width, float scale, const __half* dL_dy, float* dL_dx) {
	const uint64_t total = (uint64_t)n * n_dims;
	if (total == 0) return cudaSuccess;
	return launch_pdl(identity_backward_kernel, (uint32_t)((total + 255) / 256), 256, 0, stream, total, n_dims, width, scale, dL_dy, dL_dx);
}

cudaError_t launch_dp_barrier(cudaStream_t stream, const DpPeers& peers, uint32_t epoch) {
	if (peers.world < 1 || peers.world > DP_MAX_RANKS) return cudaErrorInvalidValue;
	dp_barrier_kernel<<<1, 32, 0, stream>>>(peers, epoch);
	return cudaGetLastError();
}

cudaError_t launch_adam_step_dp(cudaStream_t stream, const AdamParams& a, const DpPeers& peers, uint64_t first, uint64_t count, uint32_t n_matrix_weights, uint64_t n_params,
                                float loss_scale, float* weights_full_precision, float* first_moments, float* second_moments, uint32_t* param_steps) {
	if (count == 0) return cudaSuccess;
	if (first % 8 != 0 || count % 8 != 0 || n_params % 8 != 0 || n_matrix_weights % 8 != 0) return cudaErrorInvalidValue;
	const uint32_t n_groups = (uint32_t)(count / 8);
	if (peers.grads_mc && peers.params_mc) {
		adam_step_dp_kernel<true><<<blocks_for(n_groups, 256), 256, 0, stream>>>(a, peers, first, n_groups, n_matrix_weights, n_params, loss_scale, weights_full_precision, first_moments,
		                                                                          second_moments, param_steps);
	} else {
		adam_step_dp_kernel<false><<<blocks_for(n_groups, 256), 256, 0, stream>>>(a, peers, first, n_groups, n_matrix_weights, n_params, loss_scale, weights_full_precision, first_moments,
		                                                                           second_moments, param_steps);
	}
	return cudaGetLastError();
}

}  // namespace tcnnb
