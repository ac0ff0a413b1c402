// encoding_plan.h -- an encoding configuration (src/encoding.cu:60-150) resolved into a table of segments: which input dimensions each
// (nested) encoding reads, which output columns it writes, how it is padded, where its parameters live. Shared by the general path of
// the model (model.cu) and the encoding tier (encoding.cu).
//
// Composite (encodings/composite.h:135-215): "nested" array, each entry with "n_dims_to_encode" (one entry may leave it out and takes
// the remaining dimensions); outputs are concatenated; every nested encoding but the last is padded so that the NEXT one starts at a
// multiple of its required alignment (grids: n_features_per_level, everything else 1), the last one takes the padding up to the
// alignment the consumer asks for (16 in front of a network). Padding columns hold ONE (zero behind a grid); SphericalHarmonics puts its
// padding IN FRONT of its coefficients (spherical_harmonics.h:56-60). Parameters of nested encodings are concatenated in order.
// Not built: "reduction" other than Concatenation, "dims_to_encode_begin", Empty.
#pragma once
#include "feature_encodings.h"
#include "grid_config.h"
#include "grid_kernels.h"
#include "host_common.h"
#include "json_mini.h"

#include <memory>
#include <vector>

namespace tcnnb {

struct EncodingPlan {
	struct Grid {
		GridConfig cfg;
		uint32_t segment = 0;       // index into segs
		size_t param_offset = 0;    // first parameter of this grid inside the encoding's parameter vector
		std::vector<LevelInfo> levels;
		DeviceBuffer<LevelInfo> levels_dev;
	};
	uint32_t n_in = 0;
	uint32_t width = 0;             // padded output width
	uint32_t n_features = 0;        // output width without the final alignment padding
	size_t n_params = 0;
	bool composite = false;
	FeatureSegments segs{};
	std::vector<std::unique_ptr<Grid>> grids;
	bool has_plain_features() const {
		for (uint32_t i = 0; i < segs.n; ++i) if (segs.s[i].type != FEAT_GRID) return true;
		return false;
	}
};

namespace plan_detail {

inline bool is_grid_otype(const std::string& lower) { return lower == "grid" || lower == "hashgrid" || lower == "tiledgrid" || lower == "densegrid"; }

// One nested (or top-level) encoding over `n_dims` input dimensions -> segment (out_begin / n_pad filled in by the caller).
inline void add_segment(EncodingPlan& plan, uint32_t in_begin, uint32_t n_dims, const json::Value& e) {
	if (plan.segs.n >= MAX_FEATURE_SEGMENTS) throw std::runtime_error("tcnn_b200: a Composite encoding may nest at most 8 encodings");
	const std::string otype = e.value("otype", "OneBlob");  // src/encoding.cu:133
	const std::string lower = to_lower(otype);
	FeatureSegment sg{};
	sg.in_begin = in_begin;
	sg.n_in = n_dims;
	sg.scale = 1.0f;
	if (is_grid_otype(lower)) {
		auto g = std::make_unique<EncodingPlan::Grid>();
		g->cfg = parse_grid(n_dims, e);
		if (g->cfg.stochastic_interpolation) throw std::runtime_error("tcnn_b200: stochastic_interpolation is not built");
		g->segment = plan.segs.n;
		sg.type = FEAT_GRID;
		sg.n_out = g->cfg.n_levels * g->cfg.n_features_per_level;
		sg.param = g->cfg.n_features_per_level;
		plan.grids.push_back(std::move(g));
	} else if (lower == "identity") {
		sg.type = FEAT_IDENTITY;
		sg.n_out = n_dims;
		sg.scale = (float)e.value("scale", 1.0);
		sg.offset = (float)e.value("offset", 0.0);
	} else if (lower == "frequency") {
		sg.type = FEAT_FREQUENCY;
		sg.param = (uint32_t)e.value("n_frequencies", 12.0);
		sg.n_out = n_dims * sg.param * 2;
	} else if (lower == "trianglewave") {
		sg.type = FEAT_TRIANGLE_WAVE;
		sg.param = (uint32_t)e.value("n_frequencies", 12.0);
		sg.n_out = n_dims * sg.param;
	} else if (lower == "oneblob") {
		const uint32_t n_bins = (uint32_t)e.value("n_bins", 16.0);
		if (n_bins == 0 || (n_bins & (n_bins - 1)) != 0) throw std::runtime_error("Number of bins must be a power of 2");  // oneblob.h:170-172
		sg.type = FEAT_ONEBLOB;
		uint32_t log2_bins = 0;
		while ((1u << log2_bins) < n_bins) ++log2_bins;
		sg.param = log2_bins;
		sg.n_out = n_dims * n_bins;
	} else if (lower == "sphericalharmonics") {
		sg.type = FEAT_SPHERICAL_HARMONICS;
		sg.param = (uint32_t)e.value("degree", 4.0);
		if (n_dims != 3) throw std::runtime_error("Can only encode 3D directions in spherical harmonics.");  // spherical_harmonics.h:109-111
		if (sg.param == 0) throw std::runtime_error("Spherical harmonics must have positive degree.");
		if (sg.param > 8) throw std::runtime_error("Spherical harmonics are only implemented up to degree 8.");
		sg.n_out = sg.param * sg.param;
	} else if (lower == "composite" || lower == "oneblobfrequency" || lower == "nrc" || lower == "empty") {
		throw std::runtime_error("Encoding '" + otype + "' nested inside a Composite is not built in tcnn_b200");
	} else {
		throw std::runtime_error("Encoding '" + otype + "' not found");
	}
	plan.segs.s[plan.segs.n++] = sg;
}

inline uint32_t required_alignment(const EncodingPlan& plan, uint32_t seg) { return plan.segs.s[seg].type == FEAT_GRID ? plan.segs.s[seg].param : 1u; }

}  // namespace plan_detail

// `alignment`: what the consumer needs the padded width to be a multiple of (16 in front of a network, 1 for cpp::create_encoding).
// `scales_scratch_dev`: >= 128 floats of device memory (per-level grid scales are evaluated on the device).
inline void build_encoding_plan(EncodingPlan& plan, uint32_t n_in, const json::Value& cfg, uint32_t alignment, float* scales_scratch_dev) {
	using namespace plan_detail;
	plan.n_in = n_in;
	const std::string lower = to_lower(cfg.value("otype", "OneBlob"));
	if (lower == "composite") {
		plan.composite = true;
		const json::Value* nested = cfg.find("nested");
		if (!nested || nested->type != json::Value::Array) throw std::runtime_error("Must provide an array of nested encodings to CompositeEncoding.");
		if (to_lower(cfg.value("reduction", "Concatenation")) != "concatenation") throw std::runtime_error("tcnn_b200: Composite encodings are built for the Concatenation reduction only");
		uint32_t total = 0;
		uint32_t n_unspecified = 0;
		for (auto& e : nested->arr) {
			if (e.contains("dims_to_encode_begin")) throw std::runtime_error("tcnn_b200: 'dims_to_encode_begin' of Composite encodings is not built");
			if (e.contains("n_dims_to_encode")) total += (uint32_t)e.value("n_dims_to_encode", 0.0);
			else ++n_unspecified;
		}
		if (total > n_in) throw std::runtime_error("CompositeEncoding: nested encodings must not encode more dims " + std::to_string(total) + " than composite " + std::to_string(n_in));
		if (n_unspecified > 1) throw std::runtime_error("CompositeEncoding: may only leave 'n_dims_to_encode' unspecified for a single nested encoding");
		uint32_t offset = 0;
		for (auto& e : nested->arr) {
			const uint32_t dims = e.contains("n_dims_to_encode") ? (uint32_t)e.value("n_dims_to_encode", 0.0) : n_in - total;
			if (dims > 0) add_segment(plan, offset, dims, e);
			offset += dims;
		}
		if (plan.segs.n == 0) throw std::runtime_error("tcnn_b200: Composite encoding without any nested encoding");
	} else {
		add_segment(plan, 0, n_in, cfg);
	}
	// output columns and padding
	uint32_t so_far = 0;
	for (uint32_t i = 0; i < plan.segs.n; ++i) {
		FeatureSegment& sg = plan.segs.s[i];
		sg.out_begin = so_far;
		uint32_t padded = sg.n_out;
		if (i + 1 < plan.segs.n) {
			const uint32_t a = required_alignment(plan, i + 1);
			padded = next_multiple(so_far + sg.n_out, a) - so_far;
		} else {
			plan.n_features = so_far + sg.n_out;
			uint32_t a = alignment ? alignment : 1u;
			// lcm(alignment, own requirement): both are powers of two here
			const uint32_t own = required_alignment(plan, i);
			while (a % own != 0) a *= 2;
			padded = next_multiple(so_far + sg.n_out, a) - so_far;
		}
		sg.n_pad = padded - sg.n_out;
		so_far += padded;
	}
	plan.width = so_far;
	// grids: parameters, device-evaluated level scales, level tables
	size_t param_offset = 0;
	for (auto& g : plan.grids) {
		g->param_offset = param_offset;
		param_offset += g->cfg.n_params;
		g->cfg.padded_width = g->cfg.n_levels * g->cfg.n_features_per_level;
		evaluate_level_scales(g->cfg, scales_scratch_dev);
		g->levels.resize(g->cfg.n_levels);
		for (uint32_t l = 0; l < g->cfg.n_levels; ++l) g->levels[l] = make_level_info(g->cfg, l);
		g->levels_dev.resize(g->levels.size());
		TCNNB_CUDA_CHECK(cudaMemcpy(g->levels_dev.ptr, g->levels.data(), sizeof(LevelInfo) * g->levels.size(), cudaMemcpyHostToDevice));
	}
	plan.n_params = param_offset;
}

// Kernel arguments of grid `g` of the plan for rows of `n` samples: positions rows [n][x_stride], feature rows [n][row_stride]
// (pointers already advanced to the grid's own input / output columns by the caller).
inline GridKernelArgs plan_grid_args(const EncodingPlan::Grid& g, uint32_t n, const float* x_cols, uint32_t x_stride, uint32_t row_stride, float max_level = 1.0f) {
	GridKernelArgs a{};
	a.n_pos_dims = g.cfg.n_pos_dims;
	a.n_features_per_level = g.cfg.n_features_per_level;
	a.n_levels = g.cfg.n_levels;
	a.interpolation = g.cfg.interpolation;
	a.max_level = max_level;
	a.levels_dev = g.levels_dev.ptr;
	a.n_elements = n;
	a.positions = x_cols;
	a.pos_stride = x_stride;
	a.row_stride = row_stride;
	return a;
}

}  // namespace tcnnb
