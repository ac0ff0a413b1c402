// fused_ws.cu -- the hot path in ONE warp-specialised kernel: HashGrid gather + N-linear blend -> FullyFusedMLP forward (tcgen05)
// -> loss -> MLP backward (tcgen05: dgrad and wgrad) -> hash-grid gradient scatter (f16x2 reductions); without its backward half
// it is the inference kernel.
//
// Replaces, for one training step, the reference's kernel_grid (grid.h:49), kernel_mlp_fused (fully_fused_mlp.cu:500),
// relative_l2_loss / l2_loss (losses/*.h:40), kernel_mlp_fused_backward (fully_fused_mlp.cu:151), the three CUTLASS split-K
// weight-gradient GEMMs and the dL/d(encoded) GEMM (fully_fused_mlp.cu:784-836) and kernel_grid_backward (grid.h:215). No
// activation ever leaves the SM: the only HBM/L2 traffic is positions, targets, the fp16 tables (gather), their fp16 gradient
// tables (red.f16x2) and 7 K fp32 weight-gradient partial sums per CTA.
//
// Data layout on chip: operand tiles are [128 rows][64 fp16] in the canonical SWIZZLE_128B layout; the SAME bytes are consumed as
// a K-major A operand by the forward / dgrad MMAs (M = samples, K = neurons) and as an MN-major operand by the wgrad MMAs
// (M/N = neurons, K = samples), so no transposes are ever materialised. Weights are staged once per CTA: W_l [out][in] row-major
// is the K-major B operand of the forward MMA and, read MN-major, the B operand of the dgrad MMA (W_l^T). Accumulators live in
// TMEM: ACC (64 columns, reused by every layer) + one 64x64 fp32 wgrad accumulator per weight matrix that persists across all
// tiles of the CTA and is flushed once with red.global.add.f32. The backward activations g_l overwrite h_l in place.
//
// Why warp-specialised: in a bulk-synchronous structure (round 1's fused_step.cu, removed) every MMA batch ends in a CTA-wide
// barrier, so the long and uneven latencies of the table gathers / gradient reductions that fill the wait slots end up on the
// critical path of the MLP chain (12 % barrier stalls, 45 % long-scoreboard; 25 % slower for training, 43 % for inference).
// Here the two kinds of work never wait for each other inside a tile:
//
//   warps 0..3   "MLP group" (128 threads, thread t <-> tile row t <-> TMEM lane t): issues the tcgen05 MMAs and runs the
//                epilogues of one 128-sample tile at a time, tile after tile.
//   warps 4..19  "memory group": two sub-groups of 8 warps (256 threads = two threads per sample). Sub-group g owns the
//                tiles k = g, g+2, ... of this CTA: it gathers + blends tile k into enc[g], then scatters the parked
//                dL/d(enc) of its previous tile k-2, continuously, while the MLP group works on tile k-1.
//                (Running two MLP chains ping-pong in the MLP group was tried and is SLOWER -- it phase-locks the two
//                memory sub-groups; scripts/experiments/fused_ws_dual_chain.cu.txt)
//
// Hand-offs are mbarriers only:  enc_full[g]  (memory -> MLP: encoded tile ready)
//                                enc_free[g]  (tcgen05.commit -> memory: last MMA that reads enc[g] has finished)
//                                park_full[g] (MLP -> memory: dL/d(enc) of the tile parked)
//                                park_free[g] (memory -> MLP: parked gradients consumed)
// One persistent CTA per SM, 640 threads (4 MLP warps + two memory sub-groups), grid = #SMs. (A second shape -- two 384-thread
// CTAs per SM with one memory sub-group each -- was measured slower, 0.265 vs 0.213 ms on the headline configuration, and
// was removed; DESIGN.md section 3.1 keeps the number.)
#include "common.cuh"
#include "fused_common.cuh"
#include "fused_step.h"
#include "grid_device.cuh"
#include "ptx.cuh"

namespace tcnnb {

// Phase timestamps for pipeline analysis (ablation builds only: make ABLATION=1; scripts/ws_timeline.py reads them).
#ifdef TCNNB_ENABLE_ABLATION
#define WS_STAMP(role, tile_idx, slot)                                                                                   \
	do {                                                                                                                \
		if (p.dbg_clock && (tile_idx) < 16u) p.dbg_clock[((blockIdx.x * 3u + (role)) * 16u + (tile_idx)) * 16u + (slot)] = clock64(); \
	} while (0)
#else
#define WS_STAMP(role, tile_idx, slot) \
	do {                               \
	} while (0)
#endif

using namespace ptx;
using namespace fused;

namespace {

// enc buffers per CTA (tile k uses buffer k % 2). Four -- so that a gather never waits for an earlier tile's backward pass -- was
// measured 1.6 % SLOWER: the 32 KB come out of the L1 that the table gathers live on.
constexpr uint32_t WS_ENC_BUFFERS = 2;
constexpr uint32_t WS_MLP_THREADS = 128;
constexpr uint32_t WS_SUB_THREADS = 256;
constexpr uint32_t SUBS = 2;  // memory sub-groups per CTA
constexpr uint32_t WS_THREADS = WS_MLP_THREADS + SUBS * WS_SUB_THREADS;  // 640

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads) {
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

}  // namespace

template <uint32_t D, uint32_t F, bool TRAIN, bool GENERIC_ACT>
__global__ void __launch_bounds__(WS_THREADS, 1) fused_ws_kernel(const FusedStepParams p) {
	// ReLU hidden / linear output (the reference's default and the benchmark configuration) fold to the packed fast path at compile time
	const uint32_t hid_act = GENERIC_ACT ? p.activation : (uint32_t)ACT_RELU;
	const uint32_t out_act = GENERIC_ACT ? p.output_activation : (uint32_t)ACT_NONE;
	static_assert(F == 2, "fused path: F == 2");
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	const uint32_t tid = threadIdx.x;
	const uint32_t warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);  // broadcast: the compiler can treat it as warp-uniform
	const uint32_t NH = p.n_hidden_layers;
	const uint32_t in_w = p.grid.padded_width;

	// ---- shared memory: [ enc_0 .. enc_{NE-1} | h_0 .. h_{NH-1} | dy | park_0 | park_1 | W_0 .. W_{NH-1} | W_out ] barriers
	// The dynamic segment starts 1024-byte aligned (declared so above; there is no static shared memory in this kernel), so no
	// alignment slack is requested: the last kilobyte decides which shared-memory carve-out the SM uses, i.e. how much L1 is left.
	const uint32_t smem_base = smem_u32(smem_raw);
	if (smem_base & 1023u) __trap();
	const uint32_t s_enc = smem_base;
	constexpr uint32_t NE = WS_ENC_BUFFERS;  // tile k uses enc buffer k % NE
	const uint32_t s_h0 = s_enc + NE * TILE_BYTES;
	const uint32_t s_dy = s_h0 + NH * TILE_BYTES;
	const uint32_t s_park = s_dy + (TRAIN ? TILE_BYTES : 0);
	// The parked dL/d(enc) rows need enc_width columns; with an encoding of at most 32 features they fit the unused upper half
	// of the two enc tiles (tile k parks into enc[k & 1], NE == 2) and the kernel does without park tiles of its own -- 32 KB of
	// shared memory that the hardware hands to the L1 instead, which the table gathers make good use of.
	const bool park_in_enc = NE == 2 && in_w <= 32;
	const uint32_t s_w0 = s_park + (TRAIN && !park_in_enc ? 2 * TILE_BYTES : 0);
	// parked dL/d(enc) row `row`, 16-byte chunk `chunk` of buffer g: chunks 4..7 of enc[g], or park tiles of their own
	auto park_addr = [&](uint32_t g, uint32_t row, uint32_t chunk) {
		return park_in_enc ? s_enc + g * TILE_BYTES + sw128(row, 4u + chunk) : s_park + g * TILE_BYTES + sw128(row, chunk);
	};
	const uint32_t s_wout = s_w0 + NH * (WIDTH * 128);
	const uint32_t s_bars = s_wout + 16 * 128;  // 9 mbarriers
	const uint32_t bar_mma = s_bars;
	const uint32_t bar_enc_full = s_bars + 8;    // [4]
	const uint32_t bar_enc_free = s_bars + 40;   // [4]
	const uint32_t bar_park_full = s_bars + 72;  // [2]
	const uint32_t bar_park_free = s_bars + 88;  // [2]
	const uint32_t s_tmem_slot = s_bars + 104;

	const uint32_t tmem_cols = TRAIN ? ((NH + 2) * 64 <= 256 ? 256u : 512u) : 64u;
	if (tid == 0) {
		mbar_init(bar_mma, 1);
		for (uint32_t e = 0; e < NE; ++e) {
			mbar_init(bar_enc_full + 8 * e, WS_SUB_THREADS / 32);   // one arrival per memory warp
			mbar_init(bar_enc_free + 8 * e, 1);                      // tcgen05.commit
		}
		for (uint32_t g = 0; g < 2; ++g) {
			mbar_init(bar_park_full + 8 * g, WS_MLP_THREADS / 32);  // one arrival per MLP warp
			mbar_init(bar_park_free + 8 * g, WS_SUB_THREADS / 32);
		}
		fence_mbar_init();
	}
	if (warp == 0) {
		__syncwarp();
		tmem_alloc(s_tmem_slot, tmem_cols);
		tmem_relinquish();
	}

	pdl_wait();  // everything above touched only shared / tensor memory; from here on the previous kernel's writes are needed
	{
		// W_l [out][in] row-major -> 128-byte tile rows (K-major, SWIZZLE_128B). Rows / columns beyond the network's width
		// and the encoding's width are zero, which makes a 16- or 32-wide network an exact sub-problem of the 64-wide tiles.
		const __half* __restrict__ w = p.params;  // MLP weights come first in the parameter buffer
		const uint32_t NW = p.width;
		auto stage = [&](uint32_t tile, const __half* __restrict__ src, uint32_t rows, uint32_t cols, uint32_t tile_rows) {
			for (uint32_t i = tid; i < tile_rows * 8; i += WS_THREADS) {
				const uint32_t r = i >> 3, c = i & 7;
				uint4 v = make_uint4(0, 0, 0, 0);
				if (r < rows && c * 8 < cols) v = __ldg(reinterpret_cast<const uint4*>(src + r * cols + c * 8));
				st_shared_v4(tile + sw128(r, c), v.x, v.y, v.z, v.w);
			}
		};
		stage(s_w0, w, NW, in_w, WIDTH);
		w += NW * in_w;
		for (uint32_t l = 1; l < NH; ++l) {
			stage(s_w0 + l * (WIDTH * 128), w, NW, NW, WIDTH);
			w += NW * NW;
		}
		stage(s_wout, w, 16, NW, 16);
	}
	fence_proxy_async_smem();
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();

	uint32_t tmem_base;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem_slot));
	const uint32_t n_tiles = p.batch_size / TILE_M;
	// Tiles of this CTA: a CONTIGUOUS range [cta_first, cta_end) of the (spatially binned) batch. With the round-robin assignment of
	// round 1 (tile = blockIdx + k * gridDim) all 148 CTAs worked on 148 neighbouring tiles at any moment, i.e. on the same few
	// cells of the coarse dense levels, and their reductions queued up on the same L2 addresses (ablation: the 4 dense levels -- a
	// quarter of the reductions -- cost 0.047 of the kernel's 0.217 ms). Contiguous ranges put concurrent CTAs in different regions.
	const uint32_t cta_first = (uint32_t)(((uint64_t)blockIdx.x * n_tiles) / gridDim.x);
	const uint32_t cta_end = (uint32_t)(((uint64_t)(blockIdx.x + 1) * n_tiles) / gridDim.x);

	if (warp >= WS_MLP_THREADS / 32) {
		// =========================================================================================== memory group
		const uint32_t mt = tid - WS_MLP_THREADS;
		const uint32_t sub = mt / WS_SUB_THREADS;   // sub-group; with SUBS == 2 it owns the tiles (and buffers) of parity `sub`
		const uint32_t lt = mt % WS_SUB_THREADS;
		const uint32_t row = lt & 127u;
		const uint32_t hsel = lt >> 7;

		constexpr uint32_t LEVELS_PER_CHUNK = 8 / F;
		const uint32_t n_chunks = in_w / 8;
		const uint32_t level_begin = hsel * (n_chunks / 2) * LEVELS_PER_CHUNK;
		const uint32_t level_end = min(p.grid.n_levels, level_begin + (n_chunks / 2) * LEVELS_PER_CHUNK);
		const __half* __restrict__ table = p.params + p.n_mlp_params;
		__half* __restrict__ grad_table = p.grads + p.n_mlp_params;

		// The level descriptor is indexed by a warp-uniform loop counter: it is read straight from the kernel parameters
		// (constant bank, uniform registers) and costs the load/store unit nothing.
		auto load_level = [&](uint32_t level) -> const LevelInfo& { return p.grid.levels[level]; };

		float x_prev[D], x_cur[D];
		uint32_t os_cur = 0;
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) x_prev[d] = x_cur[d] = 0.0f;

		auto scatter_prev = [&](uint32_t gp) {
#pragma unroll 1
			for (uint32_t level = level_begin; level < level_end; ++level) {
				const LevelInfo& lv = load_level(level);
				LevelCorners<D> lc;
				level_corners<D>(lv, x_prev, p.grid.interpolation, lc);
				uint32_t gbits;
				const uint32_t feat = level * F;
				asm volatile("ld.shared.b32 %0, [%1];" : "=r"(gbits) : "r"(park_addr(gp, row, feat >> 3) + (feat & 7u) * 2u));
				const __half2 grad = *reinterpret_cast<const __half2*>(&gbits);
				uint32_t* __restrict__ ltab = reinterpret_cast<uint32_t*>(grad_table + (size_t)lv.offset * F);
#pragma unroll
				for (uint32_t pr = 0; pr < (1u << (D - 1)); ++pr) {
					// (GRAD_T)weight * grad -> __hmul2, then atomic f16x2 add (grid.h:252-255, vec.h:328-336)
					const __half2 a0 = __hmul2(__float2half2_rn(lc.w[2 * pr]), grad);
					const __half2 a1 = __hmul2(__float2half2_rn(lc.w[2 * pr + 1]), grad);
					const bool paired = (lc.paired >> pr) & 1u;
					if (!(TCNNB_ABLATE(ABLATE_SCATTER)) && !(TCNNB_ABLATE(ABLATE_SCATTER_DENSE) && lv.use_hash == 0) && !(TCNNB_ABLATE(ABLATE_SCATTER_HASH) && lv.use_hash != 0)) {
						scatter_pair_f16x2(ltab, lc.idx[2 * pr], lc.idx[2 * pr + 1], paired && !(TCNNB_ABLATE(ABLATE_PAIRING)), lv.wide_ok != 0, *reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1));
					}
				}
			}
		};

		auto sample_of = [&](uint32_t tile) {
			const uint32_t s = tile * TILE_M + row;
			return p.perm ? __ldg(p.perm + s) : s;
		};
		// tiles of this sub-group: k = sub, sub + 2, ... (SUBS == 2) or every tile of the CTA (SUBS == 1)
		const uint32_t k_first = SUBS == 2 ? sub : 0u, k_step = SUBS;
		const uint32_t tile_first = cta_first + k_first, tile_stride = k_step;
		uint32_t os_next = tile_first < cta_end ? sample_of(tile_first) : 0;
		uint32_t os_next2 = tile_first + tile_stride < cta_end ? sample_of(tile_first + tile_stride) : 0;
		float x_next[D];
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) x_next[d] = tile_first < cta_end ? __ldg(p.positions + (size_t)os_next * D + d) : 0.0f;
		bool have_prev = false;
		uint32_t k_prev = 0;
		for (uint32_t k = k_first, tile = tile_first; tile < cta_end; k += k_step, tile += tile_stride) {
			const uint32_t e = k % NE, je = k / NE;  // enc buffer and its use count
			const uint32_t enc_tile = s_enc + e * TILE_BYTES;
			// ---- position of this thread's sample (fetched one tile ahead, its index two tiles ahead: no dependent global
			//      load latency in front of the gather)
			os_cur = os_next;
			os_next = os_next2;
#pragma unroll
			for (uint32_t d = 0; d < D; ++d) x_cur[d] = x_next[d];
			if (tile + tile_stride < cta_end) {
#pragma unroll
				for (uint32_t d = 0; d < D; ++d) x_next[d] = __ldg(p.positions + (size_t)os_next * D + d);
			}
			if (tile + 2 * tile_stride < cta_end) os_next2 = sample_of(tile + 2 * tile_stride);

			// ---- gather tile k into enc[k % NE] once the MMAs of tile k - NE have released it
			if (lt == 0) WS_STAMP(1 + sub, k, 0);
			if (je >= 1) mbar_wait(bar_enc_free + 8 * e, (je - 1) & 1u);
			if (lt == 0) WS_STAMP(1 + sub, k, 1);
			// zero this thread's half of the row (padding features are zero, grid.h:759-766). With park_in_enc the upper half of
			// the tile holds parked gradients instead: the forward MMA reads only the first in_w columns, and what the
			// weight-gradient MMA makes of the rest lands in accumulator columns that are never flushed.
			// With park_in_enc only this thread's chunks BELOW in_w are zeroed (the parked upper half must be left alone), and only
			// when the levels do not fill them (n_levels * F < in_w, e.g. 12 levels -> 24 features padded to 32): the L0 MMA and the
			// dW0 weight-gradient MMA read all in_w columns.
			if (!park_in_enc || p.grid.n_features < in_w) {
#pragma unroll
				for (uint32_t c = 0; c < 4; ++c) {
					if (c >= n_chunks / 2 && park_in_enc) break;
					const uint32_t chunk = c < n_chunks / 2 ? hsel * (n_chunks / 2) + c : n_chunks + hsel * ((8 - n_chunks) / 2) + (c - n_chunks / 2);
					st_shared_v4(enc_tile + sw128(row, chunk), 0, 0, 0, 0);
				}
			}
			if (p.enc_identity) {
				// Identity encoding (encodings/identity.h:46-67): feature j = (half)(x_j * scale + offset) for j < D, ONE for the padding
				// features up to the network's input width; this thread writes its half of the row's 16-byte chunks.
				for (uint32_t c = 0; c < n_chunks / 2; ++c) {
					const uint32_t chunk = hsel * (n_chunks / 2) + c;
					__half f[8];
#pragma unroll
					for (uint32_t i = 0; i < 8; ++i) {
						const uint32_t jf = chunk * 8 + i;
						float v = 1.0f;
#pragma unroll
						for (uint32_t d = 0; d < D; ++d) v = jf == d ? __fmaf_rn(x_cur[d], p.identity_scale, p.identity_offset) : v;
						f[i] = __float2half_rn(v);
					}
					const uint4 q = *reinterpret_cast<const uint4*>(f);
					st_shared_v4(enc_tile + sw128(row, chunk), q.x, q.y, q.z, q.w);
					if (p.dbg_enc) *reinterpret_cast<uint4*>(p.dbg_enc + (size_t)os_cur * 64 + chunk * 8) = q;
				}
			}
			{
				// Three levels in flight: the loads of levels l+1 and l+2 are issued before the values of level l are consumed.
				// What travels with the loads is the fractional position (D floats), not the 2^D weights: they are rebuilt at
				// consumption, which is what makes the third level fit the register budget.
				struct InFlight {
					uint32_t vals[1u << D];
					float frac[D];
				};
				auto issue = [&](uint32_t level, InFlight& f) {
					const LevelInfo& lv = load_level(level);
					LevelCorners<D> lc;
					level_corners<D>(lv, x_cur, p.grid.interpolation, lc);
					const uint32_t* __restrict__ ltab = reinterpret_cast<const uint32_t*>(table + (size_t)lv.offset * F);
#pragma unroll
					for (uint32_t pr = 0; pr < (1u << (D - 1)); ++pr) {
						const bool paired = (lc.paired >> pr) & 1u;
						if (TCNNB_ABLATE(ABLATE_GATHER) || (TCNNB_ABLATE(ABLATE_GATHER_DENSE) && lv.use_hash == 0) || (TCNNB_ABLATE(ABLATE_GATHER_HASH) && lv.use_hash != 0)) {
							f.vals[2 * pr] = lc.idx[2 * pr];
							f.vals[2 * pr + 1] = lc.idx[2 * pr + 1];
						} else {
							gather_pair_f16x2(ltab, lc.idx[2 * pr], lc.idx[2 * pr + 1], paired && !(TCNNB_ABLATE(ABLATE_PAIRING)), f.vals[2 * pr], f.vals[2 * pr + 1]);
						}
					}
#pragma unroll
					for (uint32_t d = 0; d < D; ++d) f.frac[d] = lc.frac[d];
				};
				auto consume = [&](uint32_t level, const InFlight& f) {
					float w[1u << D];
					corner_weights<D>(f.frac, w);
					__half2 result = __float2half2_rn(0.0f);
#pragma unroll
					for (uint32_t idx = 0; idx < (1u << D); ++idx) {
						// fma((T)weight, grid_val, result) with T = __half -> __hfma2 (grid.h:162, vec.h:372-378)
						result = __hfma2(__float2half2_rn(w[idx]), *reinterpret_cast<const __half2*>(&f.vals[idx]), result);
					}
					const uint32_t feat = level * F;
					asm volatile("st.shared.b32 [%0], %1;" ::"r"(enc_tile + sw128(row, feat >> 3) + (feat & 7u) * 2u), "r"(*reinterpret_cast<uint32_t*>(&result)) : "memory");
					if (p.dbg_enc) *reinterpret_cast<uint32_t*>(p.dbg_enc + (size_t)os_cur * 64 + feat) = *reinterpret_cast<uint32_t*>(&result);
				};
				if (level_begin < level_end) {
					InFlight f0, f1, f2;
					issue(level_begin, f0);
					if (level_begin + 1 < level_end) issue(level_begin + 1, f1);
#pragma unroll 1
					for (uint32_t level = level_begin; level < level_end; ++level) {
						if (level + 2 < level_end) issue(level + 2, f2);
						consume(level, f0);
						f0 = f1;
						f1 = f2;
					}
				}
			}
			fence_proxy_async_smem();  // the tile is read by tcgen05.mma (async proxy)
			__syncwarp();
			if ((tid & 31u) == 0) mbar_arrive(bar_enc_full + 8 * e);
			if (lt == 0) WS_STAMP(1 + sub, k, 2);

			// ---- scatter the previous tile of this sub-group while the MLP group chews on the tiles in between
			if (TRAIN && have_prev) {
				const uint32_t gp = k_prev & 1u, jp = k_prev >> 1;
				mbar_wait(bar_park_full + 8 * gp, jp & 1u);
				if (lt == 0) WS_STAMP(1 + sub, k, 3);
				scatter_prev(gp);
				__syncwarp();
				if ((tid & 31u) == 0) mbar_arrive(bar_park_free + 8 * gp);
				if (lt == 0) WS_STAMP(1 + sub, k, 4);
			}
#pragma unroll
			for (uint32_t d = 0; d < D; ++d) x_prev[d] = x_cur[d];
			have_prev = true;
			k_prev = k;
		}
		pdl_launch_dependents();  // the optimizer's CTAs may start arriving (they block until this grid has completed)
		if (TRAIN && have_prev) {  // drain: the last tile of this sub-group
			const uint32_t gp = k_prev & 1u, jp = k_prev >> 1;
			mbar_wait(bar_park_full + 8 * gp, jp & 1u);
			scatter_prev(gp);
		}
	} else {
		// =========================================================================================== MLP group
		const uint32_t row = tid;  // 0..127
		const uint32_t lane_field = (warp * 32u) << 16;
		const uint32_t tmem_acc = tmem_base;
		uint32_t phase = 0;
		float loss_acc = 0.0f;
		bool dw_started = false;

		constexpr uint32_t IDESC_FWD_N64 = umma_idesc_f16(128, 64, 0, 0);
		constexpr uint32_t IDESC_FWD_N16 = umma_idesc_f16(128, 16, 0, 0);
		constexpr uint32_t IDESC_DGRAD = umma_idesc_f16(128, 64, 0, 1);
		constexpr uint32_t IDESC_WGRAD = umma_idesc_f16(64, 64, 1, 1);
		auto kmaj = [](uint32_t tile, uint32_t jj) { return umma_desc_sw128(tile + jj * 32u, 16u, 1024u); };
		auto mnmaj = [](uint32_t tile, uint32_t jj) { return umma_desc_sw128(tile + jj * 2048u, TILE_BYTES, 1024u); };
		auto stage_sync = [&]() {
			tmem_ld_wait();
			tc_fence_before_sync();
			fence_proxy_async_smem();
			named_bar_sync(1, WS_MLP_THREADS);
		};
		auto wait_mma = [&]() {
			mbar_wait(bar_mma, phase);
			phase ^= 1u;
			tc_fence_after_sync();
		};

		const uint32_t n_batches = TRAIN ? 2 * NH + 2 : NH + 1;
		// The sample index (through the binning permutation) and the targets are global loads whose latency, under the memory
		// group's traffic, would otherwise sit in the middle of this group's dependent MMA / epilogue chain: the index is
		// fetched one tile ahead, the targets at the top of the tile (the loss needs them three MMA stages later).
		constexpr uint32_t N_TGT_PREFETCH = 4;
		auto sample_of = [&](uint32_t tile) {
			const uint32_t s = tile * TILE_M + row;
			return p.perm ? __ldg(p.perm + s) : s;
		};
		uint32_t osample_next = cta_first < cta_end ? sample_of(cta_first) : 0;
		uint32_t k = 0;
		for (uint32_t tile = cta_first; tile < cta_end; ++tile, ++k) {
			const uint32_t g = k & 1u, j = k >> 1;
			const uint32_t e = k % NE, je = k / NE;
			const uint32_t enc_cur = s_enc + e * TILE_BYTES;
			const uint32_t osample = osample_next;
			float tgt[N_TGT_PREFETCH];
			uint4 edy_lo = make_uint4(0, 0, 0, 0), edy_hi = make_uint4(0, 0, 0, 0);
			if (TRAIN) {
				if (p.ext_dy) {  // module-tier backward: the caller's dL/d(output) row instead of targets
					edy_lo = __ldg(reinterpret_cast<const uint4*>(p.ext_dy + (size_t)osample * 16));
					edy_hi = __ldg(reinterpret_cast<const uint4*>(p.ext_dy + (size_t)osample * 16) + 1);
				} else {
#pragma unroll
					for (uint32_t q = 0; q < N_TGT_PREFETCH; ++q) tgt[q] = q < p.n_out ? __ldg(p.targets + (size_t)osample * p.n_out + q) : 0.0f;
				}
			}
			if (tile + 1 < cta_end) osample_next = sample_of(tile + 1);
			if (tid == 0) WS_STAMP(0, k, 0);
			mbar_wait(bar_enc_full + 8 * e, je & 1u);
			if (tid == 0) WS_STAMP(0, k, 1);

#pragma unroll 1
			for (uint32_t b = 0; b < n_batches; ++b) {
				if (tid == 0 && (b == 1 || b == NH + 1)) WS_STAMP(0, k, b == 1 ? 6 : 11);
				if (tid == 0 && (b == 2 || b == NH + 2)) WS_STAMP(0, k, b == 2 ? 10 : 15);
				stage_sync();
				if (tid == 0 && b == NH + 1) WS_STAMP(0, k, 2);
				if (tid == 0 && (b == 1 || b == NH + 1)) WS_STAMP(0, k, b == 1 ? 7 : 12);
				if (warp == 0 && elect_one_sync()) {  // warp-uniform branch + one elected lane: operands stay in uniform registers
					tc_fence_after_sync();
					if (b < NH) {
						const uint32_t a_tile = b == 0 ? enc_cur : s_h0 + (b - 1) * TILE_BYTES;
						const uint32_t b_tile = s_w0 + b * (WIDTH * 128);
						const uint32_t ksteps = b == 0 ? in_w / 16 : WIDTH / 16;
						for (uint32_t jj = 0; jj < ksteps; ++jj) umma_f16_ss(tmem_acc, kmaj(a_tile, jj), kmaj(b_tile, jj), IDESC_FWD_N64, jj > 0);
						if (!TRAIN && b == 0) umma_commit(bar_enc_free + 8 * e);  // inference: L0 is the only reader of enc
					} else if (b == NH) {
						const uint32_t a_tile = s_h0 + (NH - 1) * TILE_BYTES;
						for (uint32_t jj = 0; jj < WIDTH / 16; ++jj) umma_f16_ss(tmem_acc, kmaj(a_tile, jj), kmaj(s_wout, jj), IDESC_FWD_N16, jj > 0);
					} else if (b <= 2 * NH) {
						const uint32_t l = 2 * NH + 1 - b;
						const uint32_t h_prev = s_h0 + (l - 1) * TILE_BYTES;
						if (l == NH) {
							umma_f16_ss(tmem_acc, kmaj(s_dy, 0), mnmaj(s_wout, 0), IDESC_DGRAD, 0);
							const uint32_t dw = tmem_base + 64u * (1 + NH);
							for (uint32_t jj = 0; jj < TILE_M / 16; ++jj) umma_f16_ss(dw, mnmaj(h_prev, jj), mnmaj(s_dy, jj), IDESC_WGRAD, dw_started || jj > 0);
						} else {
							const uint32_t g_tile = s_h0 + l * TILE_BYTES;
							const uint32_t w_tile = s_w0 + l * (WIDTH * 128);
							for (uint32_t jj = 0; jj < WIDTH / 16; ++jj) umma_f16_ss(tmem_acc, kmaj(g_tile, jj), mnmaj(w_tile, jj), IDESC_DGRAD, jj > 0);
							const uint32_t dw = tmem_base + 64u * (1 + l);
							for (uint32_t jj = 0; jj < TILE_M / 16; ++jj) umma_f16_ss(dw, mnmaj(g_tile, jj), mnmaj(h_prev, jj), IDESC_WGRAD, dw_started || jj > 0);
						}
					} else {
						const uint32_t g_tile = s_h0;
						for (uint32_t jj = 0; jj < WIDTH / 16; ++jj) umma_f16_ss(tmem_acc, kmaj(g_tile, jj), mnmaj(s_w0, jj), IDESC_DGRAD, jj > 0);
						const uint32_t dw = tmem_base + 64u;
						for (uint32_t jj = 0; jj < TILE_M / 16; ++jj) umma_f16_ss(dw, mnmaj(g_tile, jj), mnmaj(enc_cur, jj), IDESC_WGRAD, dw_started || jj > 0);
						umma_commit(bar_enc_free + 8 * e);  // last reader of this enc buffer: release it to the memory group
					}
					umma_commit(bar_mma);
					if (b == 1 || b == NH + 1) WS_STAMP(0, k, b == 1 ? 8 : 13);
				}
				__syncwarp();
				wait_mma();
				if (tid == 0 && (b == 1 || b == NH + 1)) WS_STAMP(0, k, b == 1 ? 9 : 14);

				// ---- epilogue of batch b: this thread owns row `row`, all 64 accumulator columns (two passes of 32)
				if (b < NH) {
					const uint32_t h_tile = s_h0 + b * TILE_BYTES;
					// both halves of the accumulator row are requested before the single wait (one TMEM round trip, not two)
					uint32_t racc[2][32];
					tmem_ld_32x32b_x32(tmem_acc + lane_field, racc[0]);
					tmem_ld_32x32b_x32(tmem_acc + lane_field + 32, racc[1]);
					tmem_ld_wait();
#pragma unroll
					for (uint32_t half = 0; half < 2; ++half) {
						const uint32_t(&r)[32] = racc[half];
#pragma unroll
						for (uint32_t c = 0; c < 4; ++c) {
							const uint32_t v0 = act_pack(hid_act, r[c * 8 + 0], r[c * 8 + 1]), v1 = act_pack(hid_act, r[c * 8 + 2], r[c * 8 + 3]);
							const uint32_t v2 = act_pack(hid_act, r[c * 8 + 4], r[c * 8 + 5]), v3 = act_pack(hid_act, r[c * 8 + 6], r[c * 8 + 7]);
							st_shared_v4(h_tile + sw128(row, half * 4 + c), v0, v1, v2, v3);
							if (p.dbg_hidden) *reinterpret_cast<uint4*>(p.dbg_hidden + ((size_t)b * p.batch_size + osample) * 64 + (half * 4 + c) * 8) = make_uint4(v0, v1, v2, v3);
						}
					}
				} else if (b == NH) {
					uint32_t r[16];
					tmem_ld_32x32b_x16(tmem_acc + lane_field, r);
					tmem_ld_wait();
					__half y16[16];
#pragma unroll
					for (uint32_t q = 0; q < 16; ++q) y16[q] = act_fwd_h(out_act, __float2half_rn(__uint_as_float(r[q])));
					if (p.out_fp16) {
						uint4* dst = reinterpret_cast<uint4*>(p.out_fp16 + (size_t)osample * 16);
						dst[0] = *reinterpret_cast<uint4*>(&y16[0]);
						dst[1] = *reinterpret_cast<uint4*>(&y16[8]);
					}
					if (p.out_fp32) {
						for (uint32_t q = 0; q < p.n_out; ++q) p.out_fp32[(size_t)osample * p.n_out + q] = __half2float(y16[q]);
					}
					if (TRAIN) {
						// relative_l2_loss / l2_loss (losses/relative_l2.h:56-75, l2.h:56-74); pad lanes give 0.
						__half dy[16];
						const float n_total = (float)(p.loss_batch_size * p.n_out);
						const float luminance = p.loss_type == LOSS_RELATIVE_L2_LUMINANCE ? row_luminance(y16, p.n_out) : 0.0f;
						if (p.ext_dy) {
							// Module::backward (cpp_api.cu:115-124): dL/d(output) comes from the caller and, like the loss gradient below,
							// passes through the output activation's transfer (fully_fused_mlp.cu:758-762)
							*reinterpret_cast<uint4*>(&dy[0]) = edy_lo;
							*reinterpret_cast<uint4*>(&dy[8]) = edy_hi;
#pragma unroll
							for (uint32_t q = 0; q < 16; ++q) dy[q] = act_bwd_h(out_act, dy[q], y16[q]);
						} else
#pragma unroll
						for (uint32_t q = 0; q < 16; ++q) {
							float gq = 0.0f;
							if (q < p.n_out) {
								const float pred = __half2float(y16[q]);
								const float target = q < N_TGT_PREFETCH ? tgt[q] : __ldg(p.targets + (size_t)osample * p.n_out + q);
								float value, grad;
								loss_element(p.loss_type, pred, target, n_total, luminance, value, grad);
								gq = p.loss_scale * grad / n_total;
								loss_acc += value;
								if (p.loss_values) p.loss_values[(size_t)osample * p.n_out + q] = value;
							}
							// activation_backward_output (fully_fused_mlp.cu:755-759): dL/dy through the output activation, in fp16
						dy[q] = act_bwd_h(out_act, __float2half_rn(gq), y16[q]);
						}
						const uint4 lo = *reinterpret_cast<uint4*>(&dy[0]), hi = *reinterpret_cast<uint4*>(&dy[8]);
						st_shared_v4(s_dy + sw128(row, 0), lo.x, lo.y, lo.z, lo.w);
						st_shared_v4(s_dy + sw128(row, 1), hi.x, hi.y, hi.z, hi.w);
						if (p.dbg_dy) {
							uint4* dst = reinterpret_cast<uint4*>(p.dbg_dy + (size_t)osample * 16);
							dst[0] = lo;
							dst[1] = hi;
						}
					}
				} else if (b <= 2 * NH) {
					// g overwrites h in place: each thread rewrites the chunk of its own row it has just read
					const uint32_t l = 2 * NH + 1 - b;
					const uint32_t h_tile = s_h0 + (l - 1) * TILE_BYTES;
					uint32_t racc[2][32];
					tmem_ld_32x32b_x32(tmem_acc + lane_field, racc[0]);
					tmem_ld_32x32b_x32(tmem_acc + lane_field + 32, racc[1]);
					tmem_ld_wait();
#pragma unroll
					for (uint32_t half = 0; half < 2; ++half) {
						const uint32_t(&r)[32] = racc[half];
#pragma unroll
						for (uint32_t c = 0; c < 4; ++c) {
							uint32_t f0, f1, f2, f3;
							ld_shared_v4(h_tile + sw128(row, half * 4 + c), f0, f1, f2, f3);
							const uint32_t v0 = act_bwd_pack(hid_act, r[c * 8 + 0], r[c * 8 + 1], f0), v1 = act_bwd_pack(hid_act, r[c * 8 + 2], r[c * 8 + 3], f1);
							const uint32_t v2 = act_bwd_pack(hid_act, r[c * 8 + 4], r[c * 8 + 5], f2), v3 = act_bwd_pack(hid_act, r[c * 8 + 6], r[c * 8 + 7], f3);
							st_shared_v4(h_tile + sw128(row, half * 4 + c), v0, v1, v2, v3);
							if (p.dbg_grad_hidden) *reinterpret_cast<uint4*>(p.dbg_grad_hidden + ((size_t)(l - 1) * p.batch_size + osample) * 64 + (half * 4 + c) * 8) = make_uint4(v0, v1, v2, v3);
						}
					}
				} else {
					// dL/d(encoded): round once to fp16 (fully_fused_mlp.cu:835) and park the row for the memory group
					dw_started = true;
					if (tid == 0) WS_STAMP(0, k, 3);
					if (j >= 1) mbar_wait(bar_park_free + 8 * g, (j - 1) & 1u);  // scatter of tile k-2 has consumed park[g]
					if (tid == 0) WS_STAMP(0, k, 4);
					uint32_t racc[2][32];
					tmem_ld_32x32b_x32(tmem_acc + lane_field, racc[0]);
					tmem_ld_32x32b_x32(tmem_acc + lane_field + 32, racc[1]);
					tmem_ld_wait();
#pragma unroll
					for (uint32_t half = 0; half < 2; ++half) {
						const uint32_t(&r)[32] = racc[half];
#pragma unroll
						for (uint32_t c = 0; c < 4; ++c) {
							const uint32_t v0 = pack_half2(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1]));
							const uint32_t v1 = pack_half2(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
							const uint32_t v2 = pack_half2(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
							const uint32_t v3 = pack_half2(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
							if ((half * 4 + c) * 8 < in_w) st_shared_v4(park_addr(g, row, half * 4 + c), v0, v1, v2, v3);
							if (p.dbg_denc) *reinterpret_cast<uint4*>(p.dbg_denc + (size_t)osample * 64 + (half * 4 + c) * 8) = make_uint4(v0, v1, v2, v3);
						}
					}
					__syncwarp();
					if ((tid & 31u) == 0) mbar_arrive(bar_park_full + 8 * g);
					if (tid == 0) WS_STAMP(0, k, 5);
				}
			}
		}

		pdl_launch_dependents();
		// ---- flush the weight-gradient accumulators and the loss
		if (TRAIN) {
			tmem_ld_wait();
			tc_fence_before_sync();
			named_bar_sync(1, WS_MLP_THREADS);
			tc_fence_after_sync();
			if (dw_started) {
				// M = 64 accumulators: row m lives in TMEM lane (m % 16) + 32 * (m / 16) -> lanes 0..15 of each lane quadrant
				const uint32_t lane = tid & 31u;
				const uint32_t m = warp * 16 + lane;
				for (uint32_t l = 0; l <= NH; ++l) {
					const uint32_t dw = tmem_base + 64u * (1 + l) + lane_field;
#pragma unroll
					for (uint32_t half = 0; half < 2; ++half) {
						uint32_t r[32];
						tmem_ld_32x32b_x32(dw + half * 32, r);
						tmem_ld_wait();
						if (lane < 16 && m < p.width) {
							if (l == 0) {
								float* dst = p.dw_accum + m * in_w;
#pragma unroll
								for (uint32_t q = 0; q < 32; q += 4) {
									const uint32_t n = half * 32 + q;
									if (n < in_w) red_add_v4_f32(dst + n, __uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
								}
							} else if (l < NH) {
								float* dst = p.dw_accum + p.width * in_w + (l - 1) * p.width * p.width + m * p.width + half * 32;
#pragma unroll
								for (uint32_t q = 0; q < 32; q += 4) if (half * 32 + q < p.width) red_add_v4_f32(dst + q, __uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
							} else if (half == 0) {
								float* dst = p.dw_accum + p.width * in_w + (NH - 1) * p.width * p.width;
#pragma unroll
								for (uint32_t n = 0; n < 16; ++n) red_add_f32(dst + n * p.width + m, __uint_as_float(r[n]));
							}
						}
					}
				}
			}
#pragma unroll
			for (uint32_t o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xFFFFFFFFu, loss_acc, o);
			if ((tid & 31u) == 0 && p.loss_sum) atomicAdd(p.loss_sum, loss_acc);
		}
		tmem_ld_wait();
		tc_fence_before_sync();
	}

	__syncthreads();
	if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
size_t fused_ws_smem_bytes(uint32_t n_hidden_layers, uint32_t enc_width, bool train) {
	const size_t enc_tiles = WS_ENC_BUFFERS;
	const bool park_in_enc = enc_tiles == 2 && enc_width <= 32;
	const size_t tiles = enc_tiles + n_hidden_layers + (train ? (park_in_enc ? 1 : 3) : 0);
	return tiles * TILE_BYTES + n_hidden_layers * (WIDTH * 128) + 16 * 128 + 128 /* barriers, TMEM slot */;
}

template <uint32_t D, bool TRAIN, bool GENERIC>
static cudaError_t launch_ws_impl(const FusedStepParams& p, uint32_t n_ctas, cudaStream_t stream) {
	auto kernel = fused_ws_kernel<D, 2, TRAIN, GENERIC>;
	const size_t smem = fused_ws_smem_bytes(p.n_hidden_layers, p.grid.padded_width, TRAIN);
	cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (err != cudaSuccess) return err;
	// ask for the smallest shared-memory carve-out that holds the CTA (+1 KB the system reserves): the rest is L1
	const size_t per_sm = smem + 1024;
	err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)(per_sm * 100 / (228 * 1024))  /* rounded DOWN: the driver still has to fit the kernel, so it takes the first configuration that does (rounding up skipped the 100 KB one) */);
	if (err != cudaSuccess) return err;
	return launch_pdl(kernel, n_ctas, WS_THREADS, smem, stream, p);
}

template <uint32_t D, bool TRAIN>
static cudaError_t launch_ws_act(const FusedStepParams& p, uint32_t n_ctas, cudaStream_t stream) {
	const bool generic = p.activation != ACT_RELU || p.output_activation != ACT_NONE;
	return generic ? launch_ws_impl<D, TRAIN, true>(p, n_ctas, stream) : launch_ws_impl<D, TRAIN, false>(p, n_ctas, stream);
}

cudaError_t launch_fused_ws(const FusedStepParams& p, uint32_t n_pos_dims, bool train, uint32_t n_ctas, cudaStream_t stream) {
	if (n_pos_dims == 3) return train ? launch_ws_act<3, true>(p, n_ctas, stream) : launch_ws_act<3, false>(p, n_ctas, stream);
	if (n_pos_dims == 2) return train ? launch_ws_act<2, true>(p, n_ctas, stream) : launch_ws_act<2, false>(p, n_ctas, stream);
	return cudaErrorInvalidValue;
}

}  // namespace tcnnb
