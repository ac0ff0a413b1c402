// encoding.cu -- the stand-alone encoding tier of the C ABI (include/tcnn_b200.h, tcnnb_encoding_*).
//
// Mirrors tcnn::cpp::create_encoding (cpp_api.h:124, src/cpp_api.cu:165-174: a DifferentiableObject<__half> around
// create_encoding<__half>(n_input_dims, json, alignment 0)) for the grid encodings (src/encoding.cu:69-75: "Grid", "HashGrid",
// "TiledGrid", "DenseGrid"), i.e. what the PyTorch extension's tcnn.Encoding binds (bindings.cpp:284-343):
//   forward / inference   fp32 positions [n][n_input_dims] -> fp16 features [n][n_levels * F]      (kernel_grid, grid.h:49)
//   backward              dL_dparams (fp16, OVERWRITTEN) and / or dL_dinput fp32 from dL_doutput   (grid.h:215-358)
//   initialize_params     U(-1e-4, 1e-4) * scale from pcg32{seed}, the reference's jump-ahead pattern (grid.h:1076-1079)
// Parameters are CALLER-owned. Only the fp16 build of the reference is mirrored (Precision::Fp32 is rejected).
#include "../../include/tcnn_b200.h"

#include "grid_config.h"
#include "grid_kernels.h"
#include "host_common.h"
#include "json_mini.h"
#include "misc_kernels.h"

#include <memory>
#include <vector>

namespace tcnnb {

struct Encoding {
	GridConfig grid;
	DeviceBuffer<LevelInfo> levels_dev;
	DeviceBuffer<float> scratch;        // level scales; fp32 gradient accumulator when F == 1
	DeviceBuffer<__half> replicas;      // private copies of the coarse levels' gradients (grid_kernels.h plan_grid_scatter)
	float max_level = 1.0f;
	std::string hyperparams_json;

	GridKernelArgs args(uint32_t n, const float* positions) const {
		GridKernelArgs a{};
		a.n_pos_dims = grid.n_pos_dims;
		a.n_features_per_level = grid.n_features_per_level;
		a.n_levels = grid.n_levels;
		a.interpolation = grid.interpolation;
		a.max_level = max_level;
		a.levels_dev = levels_dev.ptr;
		a.n_elements = n;
		a.positions = positions;
		a.row_stride = grid.n_levels * grid.n_features_per_level;  // alignment 0: no padding (src/cpp_api.cu:165-174)
		return a;
	}
};

static void build_encoding(Encoding& e, uint32_t n_input_dims, const json::Value& cfg) {
	int device = 0;
	TCNNB_CUDA_CHECK(cudaGetDevice(&device));
	cudaDeviceProp prop;
	TCNNB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) throw std::runtime_error("tcnn_b200 requires an sm_100-class GPU (B200); found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
	e.grid = parse_grid(n_input_dims, cfg);
	if (e.grid.stochastic_interpolation) throw std::runtime_error("tcnn_b200: stochastic_interpolation is not built");
	e.grid.padded_width = e.grid.n_levels * e.grid.n_features_per_level;
	e.scratch.resize(128);
	evaluate_level_scales(e.grid, e.scratch.ptr);
	std::vector<LevelInfo> levels(e.grid.n_levels);
	for (uint32_t l = 0; l < e.grid.n_levels; ++l) levels[l] = make_level_info(e.grid, l);
	e.levels_dev.resize(levels.size());
	TCNNB_CUDA_CHECK(cudaMemcpy(e.levels_dev.ptr, levels.data(), sizeof(LevelInfo) * levels.size(), cudaMemcpyHostToDevice));
}

static void check_n(uint32_t n) {
	if (n == 0 || n % BATCH_GRANULARITY != 0) throw std::runtime_error("batch size " + std::to_string(n) + " must be a non-zero multiple of " + std::to_string(BATCH_GRANULARITY));
}

static void check_ptr(const void* p, const char* what, size_t align) {
	if (!p) throw std::runtime_error(std::string("encoding: ") + what + " is null.");
	if ((uintptr_t)p % align != 0) throw std::runtime_error(std::string("encoding: ") + what + " must be " + std::to_string(align) + "-byte aligned.");
}

}  // namespace tcnnb

using namespace tcnnb;

struct tcnnb_encoding {
	Encoding impl;
};

extern "C" {

int tcnnb_encoding_create(uint32_t n_input_dims, const char* encoding_json, tcnnb_encoding** out) {
	TCNNB_API_BEGIN
	if (!out) throw std::runtime_error("tcnnb_encoding_create: out is null");
	*out = nullptr;
	auto e = std::make_unique<tcnnb_encoding>();
	build_encoding(e->impl, n_input_dims, json::parse(encoding_json ? encoding_json : "{}"));
	*out = e.release();
	TCNNB_API_END
}

void tcnnb_encoding_destroy(tcnnb_encoding* e) { delete e; }
uint64_t tcnnb_encoding_n_params(const tcnnb_encoding* e) { return e->impl.grid.n_params; }
uint32_t tcnnb_encoding_n_input_dims(const tcnnb_encoding* e) { return e->impl.grid.n_pos_dims; }
uint32_t tcnnb_encoding_n_output_dims(const tcnnb_encoding* e) { return e->impl.grid.n_levels * e->impl.grid.n_features_per_level; }

int tcnnb_encoding_grid_levels(const tcnnb_encoding* e, uint32_t* n_levels, uint32_t* offsets, float* scales, uint32_t* resolutions) {
	TCNNB_API_BEGIN
	const GridConfig& g = e->impl.grid;
	if (n_levels) *n_levels = g.n_levels;
	for (uint32_t l = 0; l < g.n_levels; ++l) {
		if (offsets) offsets[l] = g.offsets[l];
		if (scales) scales[l] = g.scales[l];
		if (resolutions) resolutions[l] = (uint32_t)ceilf(g.scales[l]) + 1;
	}
	if (offsets) offsets[g.n_levels] = g.offsets[g.n_levels];
	TCNNB_API_END
}

int tcnnb_encoding_set_max_level(tcnnb_encoding* e, float max_level) {
	TCNNB_API_BEGIN
	e->impl.max_level = max_level;  // GridEncoding::set_max_level: fraction of the levels that is active
	TCNNB_API_END
}

int tcnnb_encoding_initialize_params(tcnnb_encoding* e, uint64_t seed, float* params_full_precision_dev, float scale) {
	TCNNB_API_BEGIN
	if (!params_full_precision_dev) throw std::runtime_error("encoding: params_full_precision is null.");
	HostPcg32 rng{seed};
	TCNNB_CUDA_CHECK(launch_random_uniform(nullptr, rng.device(), e->impl.grid.n_params, params_full_precision_dev, -1e-4f * scale, 1e-4f * scale));
	++g_kernel_launches;
	TCNNB_CUDA_CHECK(cudaDeviceSynchronize());
	TCNNB_API_END
}

int tcnnb_encoding_forward(tcnnb_encoding* e, tcnnb_stream stream, uint32_t n_elements, const float* input_dev, void* output_dev, const void* params_dev) {
	TCNNB_API_BEGIN
	check_n(n_elements);
	const uint32_t F = e->impl.grid.n_features_per_level;
	check_ptr(input_dev, "input", 4);
	check_ptr(output_dev, "output", 2 * F);
	check_ptr(params_dev, "params", 2 * F);
	TCNNB_CUDA_CHECK(launch_grid_forward((cudaStream_t)stream, e->impl.args(n_elements, input_dev), (const __half*)params_dev, (__half*)output_dev));
	++g_kernel_launches;
	TCNNB_API_END
}

int tcnnb_encoding_backward(tcnnb_encoding* e, tcnnb_stream stream, uint32_t n_elements, float* dL_dinput_dev, const void* dL_doutput_dev, void* dL_dparams_dev, const float* input_dev,
                            const void* params_dev) {
	TCNNB_API_BEGIN
	Encoding& enc = e->impl;
	check_n(n_elements);
	const uint32_t F = enc.grid.n_features_per_level;
	check_ptr(input_dev, "input", 4);
	check_ptr(dL_doutput_dev, "dL_doutput", 2 * F);
	cudaStream_t s = (cudaStream_t)stream;
	GridKernelArgs a = enc.args(n_elements, input_dev);
	if (dL_dparams_dev && enc.max_level >= 1.0f) {  // contended coarse levels scatter into private copies
		std::vector<LevelInfo> levels(enc.grid.n_levels);
		for (uint32_t l = 0; l < enc.grid.n_levels; ++l) levels[l] = make_level_info(enc.grid, l);
		const GridScatterPlan plan = plan_grid_scatter(levels.data(), enc.grid.n_levels, F, enc.grid.n_pos_dims, n_elements);
		if (plan.n_replicas > 1) {
			if (enc.replicas.n < plan.scratch_halfs) {
				enc.replicas.resize(plan.scratch_halfs);
				enc.replicas.zero(s);
			}
			a.replica_scratch = enc.replicas.ptr;
			a.n_replicas = plan.n_replicas;
			a.replica_entries = plan.replica_entries;
		}
	}
	if (dL_dparams_dev) {  // GradientMode::Overwrite (src/cpp_api.cu:115)
		check_ptr(dL_dparams_dev, "dL_dparams", 2 * F);
		TCNNB_CUDA_CHECK(cudaMemsetAsync(dL_dparams_dev, 0, sizeof(__half) * enc.grid.n_params, s));
		float* tmp = nullptr;
		if (F == 1) {
			enc.scratch.resize(std::max<size_t>(enc.scratch.n, enc.grid.n_params));
			tmp = enc.scratch.ptr;
			TCNNB_CUDA_CHECK(cudaMemsetAsync(tmp, 0, sizeof(float) * enc.grid.n_params, s));
		}
		TCNNB_CUDA_CHECK(launch_grid_backward(s, a, (const __half*)dL_doutput_dev, (__half*)dL_dparams_dev, tmp, enc.grid.n_params));
		g_kernel_launches += F == 1 ? 2 : 1;
	}
	if (dL_dinput_dev) {
		check_ptr(params_dev, "params", 2 * F);
		TCNNB_CUDA_CHECK(launch_grid_input_gradient(s, a, (const __half*)params_dev, (const __half*)dL_doutput_dev, dL_dinput_dev));
		++g_kernel_launches;
	}
	TCNNB_API_END
}

}  // extern "C"
