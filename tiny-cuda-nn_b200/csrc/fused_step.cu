// fused_step.cu -- the hot path in ONE kernel: HashGrid gather + trilinear blend -> FullyFusedMLP forward (tcgen05)
// -> loss -> MLP backward (tcgen05: dgrad and wgrad) -> hash-grid gradient scatter (f16x2 reductions).
//
// Replaces, for one training step, the reference's kernel_grid (grid.h:49), kernel_mlp_fused (fully_fused_mlp.cu:500),
// relative_l2_loss / l2_loss (losses/*.h:40), kernel_mlp_fused_backward (fully_fused_mlp.cu:151), the three CUTLASS
// split-K weight-gradient GEMMs and the dL/d(encoded) GEMM (fully_fused_mlp.cu:784-836) and kernel_grid_backward
// (grid.h:215). No activation ever leaves the SM: the only HBM/L2 traffic is positions, targets, the fp16 tables
// (gather), their fp16 gradient tables (red.f16x2) and 7 K fp32 weight-gradient partial sums per CTA.
//
// Structure (one CTA = 256 threads = one 128-sample tile at a time, 2 CTAs co-resident per SM = 16 warps):
//   two threads per sample: thread t (t < 128) and thread t + 128 share row t of every smem operand tile and TMEM lane t;
//   they split the resolution levels (gather / scatter) and the accumulator columns (epilogues) between them.
//   Operand tiles are [128 rows][64 fp16] in the canonical SWIZZLE_128B layout; the SAME bytes are consumed as a
//   K-major A operand by the forward/dgrad MMAs (M = samples, K = neurons) and as an MN-major operand by the wgrad
//   MMAs (M/N = neurons, K = samples), so no transposes are ever materialised.
//   Weights are staged once per CTA: W_l [out][in] row-major is the K-major B operand of the forward MMA and, read
//   MN-major, the B operand of the dgrad MMA (W_l^T).
//   Accumulators live in TMEM: ACC (64 columns, reused by every layer) + one 64x64 fp32 wgrad accumulator per weight
//   matrix that persists across all tiles of the CTA and is flushed once with red.global.add.f32.
#include "common.cuh"
#include "fused_step.h"
#include "fused_common.cuh"
#include "grid_device.cuh"
#include "ptx.cuh"

namespace tcnnb {

using namespace ptx;

using namespace fused;
using Smem = fused::SmemSync;


// ------------------------------------------------------------------------------------------------------------------
template <uint32_t D, uint32_t F, bool TRAIN>
__global__ void __launch_bounds__(256, 2) fused_step_kernel(const FusedStepParams p) {
	static_assert(F == 2, "fused path: F == 2");
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	const uint32_t tid = threadIdx.x;
	const uint32_t warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);  // broadcast: the compiler can treat it as warp-uniform
	const uint32_t row = tid & 127u;   // sample within the tile == smem tile row == TMEM lane
	const uint32_t hsel = tid >> 7;    // which half of the levels / accumulator columns this thread owns (warp-uniform)
	const uint32_t NH = p.n_hidden_layers;
	const uint32_t in_w = p.grid.padded_width;

	// ---- shared memory carve-up (all tile bases 1024-aligned)
	//   [ enc_0 | enc_1 | h_0 .. h_{NH-1} | dy (+ parked dL/d(enc)) | (park tile if in_w > 48) | W_0 .. W_{NH-1} | W_out ] barriers
	// The backward activations g_l overwrite h_l in place, so there are no separate gradient tiles.
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	Smem s;
	s.enc = smem_base;                                  // two buffers: tile t and tile t+1
	s.h0 = s.enc + 2 * TILE_BYTES;
	s.dy = s.h0 + NH * TILE_BYTES;
	const bool park_in_dy = in_w <= 48;               // 96 spare bytes per dy row hold 2 x 12 levels of parked gradients
	s.park = park_in_dy ? s.dy : s.dy + TILE_BYTES;
	s.w0 = s.dy + (TRAIN ? (park_in_dy ? 1 : 2) : 0) * TILE_BYTES;
	s.w_out = s.w0 + NH * (WIDTH * 128);
	s.bar = s.w_out + 16 * 128;
	s.tmem_slot = s.bar + 8;
	s.levels = s.bar + 16;  // LevelInfo[MAX_LEVELS] copy: dynamic indexing of kernel parameters costs a constant-cache round trip per level
	const uint32_t park_f0 = park_in_dy ? 16 + hsel * 24 : hsel * 32;  // first fp16 column of this thread's parking slots

	// ---- one-time setup: barrier, TMEM, weights
	const uint32_t tmem_cols = TRAIN ? (NH + 2) * 64 <= 256 ? 256u : 512u : 64u;
	if (tid == 0) {
		mbar_init(s.bar, 1);
		fence_mbar_init();
	}
	if (warp == 0) {
		__syncwarp();
		tmem_alloc(s.tmem_slot, tmem_cols);
		tmem_relinquish();
	}

	for (uint32_t i = tid; i < p.grid.n_levels * (uint32_t)(sizeof(LevelInfo) / 4); i += 256) {
		const uint32_t v = reinterpret_cast<const uint32_t*>(p.grid.levels)[i];
		asm volatile("st.shared.b32 [%0], %1;" ::"r"(s.levels + i * 4), "r"(v) : "memory");
	}

	pdl_wait();  // first access to global memory below: the previous kernel on the stream must have completed (common.cuh)
	// Stage the fp16 weights: W_l rows are 128-byte tile rows (K-major, SWIZZLE_128B); unused columns are zeroed.
	{
		// W_l [out][in] row-major -> 128-byte tile rows (K-major, SWIZZLE_128B). Rows / columns beyond the network's width
		// and the encoding's width are zero, which makes a 16- or 32-wide network an exact sub-problem of the 64-wide tiles.
		const __half* __restrict__ w = p.params;  // MLP weights come first in the parameter buffer
		const uint32_t NW = p.width;
		auto stage = [&](uint32_t tile, const __half* __restrict__ src, uint32_t rows, uint32_t cols, uint32_t tile_rows) {
			for (uint32_t i = tid; i < tile_rows * 8; i += 256) {
				const uint32_t r = i >> 3, c = i & 7;
				uint4 v = make_uint4(0, 0, 0, 0);
				if (r < rows && c * 8 < cols) v = __ldg(reinterpret_cast<const uint4*>(src + r * cols + c * 8));
				st_shared_v4(tile + sw128(r, c), v.x, v.y, v.z, v.w);
			}
		};
		stage(s.w0, w, NW, in_w, WIDTH);
		w += NW * in_w;
		for (uint32_t l = 1; l < NH; ++l) {
			stage(s.w0 + l * (WIDTH * 128), w, NW, NW, WIDTH);
			w += NW * NW;
		}
		stage(s.w_out, w, 16, NW, 16);
	}
	fence_proxy_async_smem();
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();

	uint32_t tmem_base;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s.tmem_slot));
	const uint32_t tmem_acc = tmem_base;                          // 64 columns
	const uint32_t lane_field = ((warp & 3u) * 32u) << 16;        // this warp's TMEM lane quadrant
	uint32_t phase = 0;

	// Instruction descriptors.
	constexpr uint32_t IDESC_FWD_N64 = umma_idesc_f16(128, 64, 0, 0);   // A K-major, B K-major
	constexpr uint32_t IDESC_FWD_N16 = umma_idesc_f16(128, 16, 0, 0);
	constexpr uint32_t IDESC_DGRAD = umma_idesc_f16(128, 64, 0, 1);     // A K-major, B MN-major (W^T)
	constexpr uint32_t IDESC_WGRAD = umma_idesc_f16(64, 64, 1, 1);      // A, B MN-major (K = samples)

	// K-major operand, k-step j: +32 bytes. MN-major operand, k-step j (16 rows): +2048 bytes.
	auto kmaj = [](uint32_t tile, uint32_t j) { return umma_desc_sw128(tile + j * 32u, 16u, 1024u); };
	auto mnmaj = [](uint32_t tile, uint32_t j) { return umma_desc_sw128(tile + j * 2048u, TILE_BYTES, 1024u); };

	// All threads: make this thread's smem tile writes / TMEM reads visible, then rendezvous.
	auto stage_sync = [&]() {
		tmem_ld_wait();
		tc_fence_before_sync();
		fence_proxy_async_smem();
		__syncthreads();
	};
	auto wait_mma = [&]() {
		mbar_wait(s.bar, phase);
		phase ^= 1u;
		tc_fence_after_sync();
	};

	// ---- this thread's share of the resolution levels: the level-bearing 16-byte chunks of a row are split evenly
	constexpr uint32_t LEVELS_PER_CHUNK = 8 / F;
	const uint32_t n_chunks = in_w / 8;  // even: in_w is a multiple of 16
	const uint32_t level_begin = hsel * (n_chunks / 2) * LEVELS_PER_CHUNK;
	const uint32_t level_end = min(p.grid.n_levels, level_begin + (n_chunks / 2) * LEVELS_PER_CHUNK);
	const uint32_t n_my_levels = level_end > level_begin ? level_end - level_begin : 0;
	const __half* __restrict__ table = p.params + p.n_mlp_params;

	auto load_level = [&](uint32_t level) {
		LevelInfo lv;
		uint32_t* w = reinterpret_cast<uint32_t*>(&lv);
		const uint32_t base = s.levels + level * (uint32_t)sizeof(LevelInfo);
		static_assert(sizeof(LevelInfo) == 32, "LevelInfo layout");
		asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(base));
		asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(base + 16));
		return lv;
	};

	// Gather + N-linear blend of part `part` of `parts` of this thread's levels for the sample at `x` into tile `enc_tile`.
	// Loops are deliberately not unrolled: memory-system throughput, not load latency, bounds the gather, and an unrolled
	// body overflows the instruction caches (profiles/).
	auto gather_part = [&](const float (&x)[D], uint32_t enc_tile, uint32_t sample, uint32_t part, uint32_t parts) {
		if (part == 0) {
			// zero this thread's 4 chunks of the row first (padding features are zero, grid.h:759-766)
#pragma unroll
			for (uint32_t c = 0; c < 4; ++c) {
				const uint32_t chunk = c < n_chunks / 2 ? hsel * (n_chunks / 2) + c : n_chunks + hsel * ((8 - n_chunks) / 2) + (c - n_chunks / 2);
				st_shared_v4(enc_tile + sw128(row, chunk), 0, 0, 0, 0);
			}
		}
		const uint32_t lb = level_begin + (n_my_levels * part) / parts, le = level_begin + (n_my_levels * (part + 1)) / parts;
		// Two levels in flight: the loads of level l+1 are issued before the values of level l are consumed, so the
		// (long, loaded-L2) latency of one level hides behind the index arithmetic and the loads of the next.
		struct InFlight {
			uint32_t vals[1u << D];
			uint32_t w16[1u << D];  // (half)weight duplicated into both halves
		};
		auto issue = [&](uint32_t level, InFlight& f) {
			const LevelInfo lv = load_level(level);
			LevelCorners<D> lc;
			level_corners<D>(lv, x, p.grid.interpolation, lc);
			const uint32_t* __restrict__ lt = reinterpret_cast<const uint32_t*>(table + (size_t)lv.offset * F);
#pragma unroll
			for (uint32_t pr = 0; pr < (1u << (D - 1)); ++pr) {
				const bool paired = (lc.paired >> pr) & 1u;
				if (TCNNB_ABLATE(ABLATE_GATHER)) {
					f.vals[2 * pr] = lc.idx[2 * pr];
					f.vals[2 * pr + 1] = lc.idx[2 * pr + 1];
				} else {
					gather_pair_f16x2(lt, lc.idx[2 * pr], lc.idx[2 * pr + 1], paired && !(TCNNB_ABLATE(ABLATE_PAIRING)), f.vals[2 * pr], f.vals[2 * pr + 1]);
				}
			}
#pragma unroll
			for (uint32_t i = 0; i < (1u << D); ++i) {
				const __half2 h = __float2half2_rn(lc.w[i]);
				f.w16[i] = *reinterpret_cast<const uint32_t*>(&h);
			}
		};
		auto consume = [&](uint32_t level, const InFlight& f) {
			__half2 result = __float2half2_rn(0.0f);
#pragma unroll
			for (uint32_t idx = 0; idx < (1u << D); ++idx) {
				// fma((T)weight, grid_val, result) with T = __half -> __hfma2 (grid.h:162, vec.h:372-378)
				result = __hfma2(*reinterpret_cast<const __half2*>(&f.w16[idx]), *reinterpret_cast<const __half2*>(&f.vals[idx]), result);
			}
			const uint32_t feat = level * F;
			asm volatile("st.shared.b32 [%0], %1;" ::"r"(enc_tile + sw128(row, feat >> 3) + (feat & 7u) * 2u), "r"(*reinterpret_cast<uint32_t*>(&result)) : "memory");
			if (p.dbg_enc) *reinterpret_cast<uint32_t*>(p.dbg_enc + (size_t)sample * 64 + feat) = *reinterpret_cast<uint32_t*>(&result);
		};
		if (lb < le) {
			InFlight cur, nxt;
			issue(lb, cur);
#pragma unroll 1
			for (uint32_t level = lb; level < le; ++level) {
				if (level + 1 < le) issue(level + 1, nxt);
				consume(level, cur);
				cur = nxt;
			}
		}
	};

	// Scatter part `part` of `parts` of this thread's levels: the parked fp16 dL/d(enc) of the sample at `x`.
	auto scatter_part = [&](const float (&x)[D], uint32_t part, uint32_t parts) {
		__half* __restrict__ grad_table = p.grads + p.n_mlp_params;
		const uint32_t lb = level_begin + (n_my_levels * part) / parts, le = level_begin + (n_my_levels * (part + 1)) / parts;
#pragma unroll 1
		for (uint32_t level = lb; level < le; ++level) {
			const LevelInfo lv = load_level(level);
			LevelCorners<D> lc;
			level_corners<D>(lv, x, p.grid.interpolation, lc);
			uint32_t gbits;
			const uint32_t feat = park_f0 + (level - level_begin) * F;  // 2 features = one 32-bit word
			asm volatile("ld.shared.b32 %0, [%1];" : "=r"(gbits) : "r"(s.park + sw128(row, feat >> 3) + (feat & 7u) * 2u));
			const __half2 grad = *reinterpret_cast<const __half2*>(&gbits);
			uint32_t* __restrict__ lt = reinterpret_cast<uint32_t*>(grad_table + (size_t)lv.offset * F);
#pragma unroll
			for (uint32_t pr = 0; pr < (1u << (D - 1)); ++pr) {
				// (GRAD_T)weight * grad -> __hmul2, then atomic f16x2 add (grid.h:252-255, vec.h:328-336)
				const __half2 a0 = __hmul2(__float2half2_rn(lc.w[2 * pr]), grad);
				const __half2 a1 = __hmul2(__float2half2_rn(lc.w[2 * pr + 1]), grad);
				const bool paired = (lc.paired >> pr) & 1u;
				if (!(TCNNB_ABLATE(ABLATE_SCATTER))) {
					scatter_pair_f16x2(lt, lc.idx[2 * pr], lc.idx[2 * pr + 1], paired && !(TCNNB_ABLATE(ABLATE_PAIRING)), lv.wide_ok != 0, *reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1));
				}
			}
		}
	};

	// Tile row -> caller's sample index (identity unless the batch was spatially binned) and its position.
	auto load_x = [&](uint32_t tile, float (&x)[D], uint32_t& osample) {
		osample = tile * TILE_M + row;
		if (p.perm) osample = __ldg(p.perm + osample);
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(p.positions + (size_t)osample * D + d);
	};

	// ---- software pipeline over this CTA's tiles: while the tensor core works on tile t (2*NH + 2 MMA batches, each followed
	// by an mbarrier wait), the wait slots are filled with the gather of tile t+1 (slots 0,1) and the scatter of tile t-1
	// (slots 2..). The gather / scatter of a tile therefore never sit on the critical path of its own MLP chain.
	constexpr uint32_t GATHER_PARTS = 2;
	const uint32_t n_slots = 2 * NH + 2;
	const uint32_t scatter_parts = TRAIN ? min(4u, n_slots - GATHER_PARTS) : 0u;

	float loss_acc = 0.0f;
	bool dw_started = false;
	bool have_prev = false;
	float x_prev[D], x_cur[D], x_next[D];
	uint32_t os_cur = 0, os_next = 0;  // caller's sample index of this thread's row in the current / next tile
#pragma unroll
	for (uint32_t d = 0; d < D; ++d) x_prev[d] = x_cur[d] = x_next[d] = 0.0f;

	const uint32_t n_tiles = p.batch_size / TILE_M;
	uint32_t tile = blockIdx.x;
	if (tile < n_tiles) {
		load_x(tile, x_cur, os_cur);
		for (uint32_t part = 0; part < GATHER_PARTS; ++part) gather_part(x_cur, s.enc, os_cur, part, GATHER_PARTS);
	}
	uint32_t buf = 0;
	for (; tile < n_tiles; tile += gridDim.x, buf ^= 1u) {
		const uint32_t osample = os_cur;  // caller's sample index: targets and per-sample outputs
		const uint32_t enc_cur = s.enc + buf * TILE_BYTES, enc_nxt = s.enc + (buf ^ 1u) * TILE_BYTES;
		const uint32_t next_tile = tile + gridDim.x;
		const bool have_next = next_tile < n_tiles;
		if (have_next) load_x(next_tile, x_next, os_next);
		// One tile = n_batches MMA batches. Batch b: stage_sync -> one thread issues the MMAs -> every thread fills the wait
		// slot with memory work of the neighbouring tiles -> wait for the tensor core -> epilogue of batch b.
		//   b <  NH          forward hidden layer b
		//   b == NH          output layer + loss
		//   NH < b <= 2NH    backward through layer l = 2NH + 1 - b (l = NH is the output layer): dgrad + wgrad
		//   b == 2NH + 1     first layer: dL/d(encoded) + dW_0
		const uint32_t n_batches = TRAIN ? 2 * NH + 2 : NH + 1;
#pragma unroll 1
		for (uint32_t b = 0; b < n_batches; ++b) {
			stage_sync();
			if (warp == 0 && elect_one_sync()) {  // warp-uniform branch + one elected lane: operands stay in uniform registers
				tc_fence_after_sync();
				if (b < NH) {
					const uint32_t a_tile = b == 0 ? enc_cur : s.h0 + (b - 1) * TILE_BYTES;
					const uint32_t b_tile = s.w0 + b * (WIDTH * 128);
					const uint32_t ksteps = b == 0 ? in_w / 16 : WIDTH / 16;
					for (uint32_t j = 0; j < ksteps; ++j) umma_f16_ss(tmem_acc, kmaj(a_tile, j), kmaj(b_tile, j), IDESC_FWD_N64, j > 0);
				} else if (b == NH) {
					const uint32_t a_tile = s.h0 + (NH - 1) * TILE_BYTES;
					for (uint32_t j = 0; j < WIDTH / 16; ++j) umma_f16_ss(tmem_acc, kmaj(a_tile, j), kmaj(s.w_out, j), IDESC_FWD_N16, j > 0);
				} else if (b <= 2 * NH) {
					// g_{NH-1} = (dy . W_out) * act'(h_{NH-1});  dW_out^T += h_{NH-1}^T . dy   (fully_fused_mlp.cu:192-240, :784-787)
					// g_{l-1}  = (g_l . W_l)   * act'(h_{l-1});   dW_l     += g_l^T . h_{l-1}   (fully_fused_mlp.cu:248-250, :815-830)
					const uint32_t l = 2 * NH + 1 - b;
					const uint32_t h_prev = s.h0 + (l - 1) * TILE_BYTES;
					if (l == NH) {
						umma_f16_ss(tmem_acc, kmaj(s.dy, 0), mnmaj(s.w_out, 0), IDESC_DGRAD, 0);
						const uint32_t dw = tmem_base + 64u * (1 + NH);
						for (uint32_t j = 0; j < TILE_M / 16; ++j) umma_f16_ss(dw, mnmaj(h_prev, j), mnmaj(s.dy, j), IDESC_WGRAD, dw_started || j > 0);
					} else {
						const uint32_t g_tile = s.h0 + l * TILE_BYTES;  // g_l lives in h_l's tile
						const uint32_t w_tile = s.w0 + l * (WIDTH * 128);
						for (uint32_t j = 0; j < WIDTH / 16; ++j) umma_f16_ss(tmem_acc, kmaj(g_tile, j), mnmaj(w_tile, j), IDESC_DGRAD, j > 0);
						const uint32_t dw = tmem_base + 64u * (1 + l);
						for (uint32_t j = 0; j < TILE_M / 16; ++j) umma_f16_ss(dw, mnmaj(g_tile, j), mnmaj(h_prev, j), IDESC_WGRAD, dw_started || j > 0);
					}
				} else {
					// d_enc = g_0 . W_0 (fully_fused_mlp.cu:833-836);  dW_0 += g_0^T . enc (:827-830)
					const uint32_t g_tile = s.h0;
					for (uint32_t j = 0; j < WIDTH / 16; ++j) umma_f16_ss(tmem_acc, kmaj(g_tile, j), mnmaj(s.w0, j), IDESC_DGRAD, j > 0);
					const uint32_t dw = tmem_base + 64u;
					for (uint32_t j = 0; j < TILE_M / 16; ++j) umma_f16_ss(dw, mnmaj(g_tile, j), mnmaj(enc_cur, j), IDESC_WGRAD, dw_started || j > 0);
				}
				umma_commit(s.bar);
			}
			__syncwarp();

			// ---- wait slot: gather of tile t+1 (slots 0, 1), scatter of tile t-1 (following slots)
			if (b < GATHER_PARTS) {
				if (have_next) gather_part(x_next, enc_nxt, os_next, b, GATHER_PARTS);
			} else if (TRAIN && have_prev && b < GATHER_PARTS + scatter_parts) {
				scatter_part(x_prev, b - GATHER_PARTS, scatter_parts);
			}
			wait_mma();

			// ---- epilogue of batch b
			if (b < NH) {
				const uint32_t h_tile = s.h0 + b * TILE_BYTES;
				uint32_t r[32];
				tmem_ld_32x32b_x32(tmem_acc + lane_field + hsel * 32, r);  // this thread's 32 accumulator columns
				tmem_ld_wait();
#pragma unroll
				for (uint32_t c = 0; c < 4; ++c) {
					const uint32_t v0 = act_pack(p.activation, r[c * 8 + 0], r[c * 8 + 1]), v1 = act_pack(p.activation, r[c * 8 + 2], r[c * 8 + 3]);
					const uint32_t v2 = act_pack(p.activation, r[c * 8 + 4], r[c * 8 + 5]), v3 = act_pack(p.activation, r[c * 8 + 6], r[c * 8 + 7]);
					st_shared_v4(h_tile + sw128(row, hsel * 4 + c), v0, v1, v2, v3);
					if (p.dbg_hidden) *reinterpret_cast<uint4*>(p.dbg_hidden + ((size_t)b * p.batch_size + osample) * 64 + (hsel * 4 + c) * 8) = make_uint4(v0, v1, v2, v3);
				}
			} else if (b == NH) {
				if (hsel == 0) {
					uint32_t r[16];
					tmem_ld_32x32b_x16(tmem_acc + lane_field, r);
					tmem_ld_wait();
					// The reference's network output is fp16 (fully_fused_mlp.cu:421-476); everything downstream reads that rounding.
					__half y16[16];
#pragma unroll
					for (uint32_t j = 0; j < 16; ++j) y16[j] = act_fwd_h(p.output_activation, __float2half_rn(__uint_as_float(r[j])));
					if (p.out_fp16) {
						uint4* dst = reinterpret_cast<uint4*>(p.out_fp16 + (size_t)osample * 16);
						dst[0] = *reinterpret_cast<uint4*>(&y16[0]);
						dst[1] = *reinterpret_cast<uint4*>(&y16[8]);
					}
					if (p.out_fp32) {
						for (uint32_t j = 0; j < p.n_out; ++j) p.out_fp32[(size_t)osample * p.n_out + j] = __half2float(y16[j]);
					}
					if (TRAIN) {
						// relative_l2_loss / l2_loss (losses/relative_l2.h:56-75, l2.h:56-74); pad lanes give 0.
						__half dy[16];
						const float n_total = (float)(p.loss_batch_size * p.n_out);
						if (p.ext_dy) {
							// Module::backward (cpp_api.cu:115-124): the caller's dL/d(output), through the output activation's transfer
							const uint4* src = reinterpret_cast<const uint4*>(p.ext_dy + (size_t)osample * 16);
							*reinterpret_cast<uint4*>(&dy[0]) = __ldg(src);
							*reinterpret_cast<uint4*>(&dy[8]) = __ldg(src + 1);
#pragma unroll
							for (uint32_t j = 0; j < 16; ++j) dy[j] = act_bwd_h(p.output_activation, dy[j], y16[j]);
						} else
#pragma unroll
						for (uint32_t j = 0; j < 16; ++j) {
							float g = 0.0f;
							if (j < p.n_out) {
								const float pred = __half2float(y16[j]);
								const float diff = pred - __ldg(p.targets + (size_t)osample * p.n_out + j);
								float value, grad;
								if (p.loss_type == LOSS_RELATIVE_L2) {
									const float psq = pred * pred + 0.01f;
									value = diff * diff / psq / n_total;
									grad = 2.0f * diff / psq;
								} else {
									value = diff * diff / n_total;
									grad = 2.0f * diff;
								}
								g = p.loss_scale * grad / n_total;
								loss_acc += value;
								if (p.loss_values) p.loss_values[(size_t)osample * p.n_out + j] = value;
							}
							dy[j] = act_bwd_h(p.output_activation, __float2half_rn(g), y16[j]);
						}
						const uint4 lo = *reinterpret_cast<uint4*>(&dy[0]), hi = *reinterpret_cast<uint4*>(&dy[8]);
						st_shared_v4(s.dy + sw128(row, 0), lo.x, lo.y, lo.z, lo.w);
						st_shared_v4(s.dy + sw128(row, 1), hi.x, hi.y, hi.z, hi.w);
						if (p.dbg_dy) {
							uint4* dst = reinterpret_cast<uint4*>(p.dbg_dy + (size_t)osample * 16);
							dst[0] = lo;
							dst[1] = hi;
						}
					}
				}
			} else if (b <= 2 * NH) {
				// g overwrites h IN PLACE: every thread rewrites only chunks of its own row that it has just read, and all MMAs
				// that read this tile (forward A operand, wgrad B operand of the batch just completed) are finished.
				const uint32_t l = 2 * NH + 1 - b;
				const uint32_t h_tile = s.h0 + (l - 1) * TILE_BYTES;
				uint32_t r[32];
				tmem_ld_32x32b_x32(tmem_acc + lane_field + hsel * 32, r);
				tmem_ld_wait();
#pragma unroll
				for (uint32_t c = 0; c < 4; ++c) {
					uint32_t f0, f1, f2, f3;
					ld_shared_v4(h_tile + sw128(row, hsel * 4 + c), f0, f1, f2, f3);
					const uint32_t v0 = act_bwd_pack(p.activation, r[c * 8 + 0], r[c * 8 + 1], f0), v1 = act_bwd_pack(p.activation, r[c * 8 + 2], r[c * 8 + 3], f1);
					const uint32_t v2 = act_bwd_pack(p.activation, r[c * 8 + 4], r[c * 8 + 5], f2), v3 = act_bwd_pack(p.activation, r[c * 8 + 6], r[c * 8 + 7], f3);
					st_shared_v4(h_tile + sw128(row, hsel * 4 + c), v0, v1, v2, v3);
					if (p.dbg_grad_hidden) *reinterpret_cast<uint4*>(p.dbg_grad_hidden + ((size_t)(l - 1) * p.batch_size + osample) * 64 + (hsel * 4 + c) * 8) = make_uint4(v0, v1, v2, v3);
				}
			} else {
				// dL/d(encoded) is an fp16 matrix in the reference (output of fc_multiply, fully_fused_mlp.cu:835): round the fp32
				// accumulator once and park this thread's levels; they are scattered during the NEXT tile's wait slots.
				dw_started = true;
				uint32_t r[32];
				tmem_ld_32x32b_x32(tmem_acc + lane_field + hsel * (in_w / 2), r);
				tmem_ld_wait();
#pragma unroll
				for (uint32_t k = 0; k < 16; ++k) {
					if (k < n_my_levels) {
						const uint32_t v = pack_half2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
						const uint32_t feat = park_f0 + k * F;
						asm volatile("st.shared.b32 [%0], %1;" ::"r"(s.park + sw128(row, feat >> 3) + (feat & 7u) * 2u), "r"(v) : "memory");
						if (p.dbg_denc) *reinterpret_cast<uint32_t*>(p.dbg_denc + (size_t)osample * 64 + (level_begin + k) * F) = v;
					}
				}
			}
		}
		if (TRAIN) {
#pragma unroll
			for (uint32_t d = 0; d < D; ++d) x_prev[d] = x_cur[d];
			have_prev = true;
		}
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) x_cur[d] = x_next[d];
		os_cur = os_next;
	}

	// ---- drain: scatter of the last tile
	if (TRAIN && have_prev) {
		for (uint32_t part = 0; part < scatter_parts; ++part) scatter_part(x_prev, part, scatter_parts);
	}

	// ================================================================ flush weight gradients, loss, teardown
	if (TRAIN) {
		tmem_ld_wait();
		tc_fence_before_sync();
		__syncthreads();
		tc_fence_after_sync();
		if (dw_started) {
			// M = 64 accumulators: row m lives in TMEM lane (m % 16) + 32 * (m / 16) -> lanes 0..15 of each lane quadrant.
			// Warp w reads quadrant w & 3 and column half w >> 2.
			const uint32_t lane = tid & 31u;
			const uint32_t m = (warp & 3u) * 16 + lane;
			const uint32_t half = hsel;
			for (uint32_t l = 0; l <= NH; ++l) {
				const uint32_t dw = tmem_base + 64u * (1 + l) + lane_field;
				uint32_t r[32];
				tmem_ld_32x32b_x32(dw + half * 32, r);
				tmem_ld_wait();
				if (lane < 16 && m < p.width) {
					if (l == 0) {
						// dW_0[out = m][in = n], n < in_w
						float* dst = p.dw_accum + m * in_w;
#pragma unroll
						for (uint32_t k = 0; k < 32; k += 4) {
							const uint32_t n = half * 32 + k;
							if (n < in_w) red_add_v4_f32(dst + n, __uint_as_float(r[k]), __uint_as_float(r[k + 1]), __uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
						}
					} else if (l < NH) {
						float* dst = p.dw_accum + p.width * in_w + (l - 1) * p.width * p.width + m * p.width + half * 32;
#pragma unroll
						for (uint32_t k = 0; k < 32; k += 4) if (half * 32 + k < p.width) red_add_v4_f32(dst + k, __uint_as_float(r[k]), __uint_as_float(r[k + 1]), __uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
					} else if (half == 0) {
						// accumulator holds dW_out^T[in = m][out = n], n < 16
						float* dst = p.dw_accum + p.width * in_w + (NH - 1) * p.width * p.width;
#pragma unroll
						for (uint32_t n = 0; n < 16; ++n) red_add_f32(dst + n * p.width + m, __uint_as_float(r[n]));
					}
				}
			}
		}
		// loss: warp reduce, one atomic per warp
#pragma unroll
		for (uint32_t o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xFFFFFFFFu, loss_acc, o);
		if ((tid & 31u) == 0 && p.loss_sum) atomicAdd(p.loss_sum, loss_acc);
	}

	tmem_ld_wait();
	tc_fence_before_sync();
	__syncthreads();
	if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
size_t fused_step_smem_bytes(uint32_t n_hidden_layers, uint32_t in_w, bool train) {
	const size_t tiles = 2 + n_hidden_layers + (train ? (in_w <= 48 ? 1 : 2) : 0);
	return tiles * TILE_BYTES + n_hidden_layers * (WIDTH * 128) + 16 * 128 + 16 + MAX_LEVELS * sizeof(LevelInfo) + 1024 /* alignment slack */;
}

template <uint32_t D, bool TRAIN>
static cudaError_t launch_impl(const FusedStepParams& p, uint32_t n_ctas, cudaStream_t stream) {
	auto kernel = fused_step_kernel<D, 2, TRAIN>;
	const size_t smem = fused_step_smem_bytes(p.n_hidden_layers, p.grid.padded_width, TRAIN);
	cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (err != cudaSuccess) return err;
	return launch_pdl(kernel, n_ctas, 256, smem, stream, p);
}

cudaError_t launch_fused_step(const FusedStepParams& p, uint32_t n_pos_dims, bool train, uint32_t n_ctas, cudaStream_t stream) {
	if (train) return cudaErrorInvalidValue;  // the training step is fused_ws.cu; only the inference instantiation of this kernel is built
	if (n_pos_dims == 3) return launch_impl<3, false>(p, n_ctas, stream);
	if (n_pos_dims == 2) return launch_impl<2, false>(p, n_ctas, stream);
	return cudaErrorInvalidValue;
}

}  // namespace tcnnb
