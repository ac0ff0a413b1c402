// ref_harness -- drives the UNMODIFIED reference (tiny-cuda-nn, compiled from /root/reference by
// oracle/Makefile into oracle/_ref/) through its own public C++ API.
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY. Nothing under tiny-cuda-nn_b200/ links or calls this.
// Two jobs:
//   dump   write golden vectors (inputs, initial params, encoded features, outputs, losses,
//          gradients, post-step params) for tests/golden/ -- see tests/golden/make_golden.sh
//   bench  time trainer->training_step / network->inference with CUDA events for the
//          `bench.py --impl reference` arm (jit on / jit off / CutlassMLP via the JSON config)
//   dumpbig  the same at the benchmarked size, sub-sampled so that the vectors are small enough to commit
//   mlpdump / mlpbench  the network on its own (create_network<T>, fp16 in / out): golden vectors and inference throughput
//   probe  print the per-level grid scale / resolution as evaluated on the device with the
//          reference's build flags next to the host evaluation (SURVEY.md §7 "hard parts")
//
// API used: tcnn::create_from_config (config.h:53), Trainer::training_step / loss (trainer.h:254,372),
// NetworkWithInputEncoding::inference / encoding() (object.h:214, network_with_input_encoding.h:169),
// generate_random_uniform (random.h:69).
#include <tiny-cuda-nn/common_device.h>
#include <tiny-cuda-nn/config.h>
#include <tiny-cuda-nn/network.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <sys/stat.h>
#include <vector>

using namespace tcnn;
using precision_t = network_precision_t;

static json load_json(const std::string& path) {
	std::ifstream f{path};
	if (!f) {
		throw std::runtime_error{"cannot open " + path};
	}
	return json::parse(f, nullptr, true, true);
}

template <typename T>
static void write_bin(const std::string& path, const T* dev, size_t n) {
	std::vector<T> h(n);
	CUDA_CHECK_THROW(cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
	std::ofstream f{path, std::ios::binary};
	f.write((const char*)h.data(), n * sizeof(T));
}

template <typename T>
static void write_host_bin(const std::string& path, const std::vector<T>& h) {
	std::ofstream f{path, std::ios::binary};
	f.write((const char*)h.data(), h.size() * sizeof(T));
}

// Closed-form smooth target field in [0,1]^n_out, evaluated on the host in fp32. The same expression
// is restated in tiny-cuda-nn_b200 (bench / tests) so that both arms train on identical data.
static void make_targets(const std::vector<float>& x, uint32_t n_in, uint32_t n_out, uint32_t B, std::vector<float>& y) {
	y.resize((size_t)n_out * B);
	for (uint32_t i = 0; i < B; ++i) {
		for (uint32_t c = 0; c < n_out; ++c) {
			float phase = 0.0f;
			for (uint32_t d = 0; d < n_in; ++d) {
				phase += x[(size_t)i * n_in + d] * (float)(c + 1 + d) / (float)(1u << d);
			}
			y[(size_t)i * n_out + c] = 0.5f + 0.5f * sinf(6.2831853f * phase);
		}
	}
}

__global__ void probe_scales(uint32_t n_levels, float log2_per_level_scale, uint32_t base_resolution, float* scales, uint32_t* resolutions) {
	const uint32_t l = threadIdx.x;
	if (l >= n_levels) return;
	const float s = grid_scale(l, log2_per_level_scale, base_resolution);
	scales[l] = s;
	resolutions[l] = grid_resolution(s);
}

static int cmd_probe(int argc, char** argv) {
	const float per_level_scale = argc > 2 ? (float)atof(argv[2]) : 1.5f;
	const uint32_t base = argc > 3 ? atoi(argv[3]) : 16;
	const uint32_t L = argc > 4 ? atoi(argv[4]) : 16;
	GPUMemory<float> s(L);
	GPUMemory<uint32_t> r(L);
	const float l2 = std::log2(per_level_scale);
	probe_scales<<<1, 128>>>(L, l2, base, s.data(), r.data());
	std::vector<float> hs(L);
	std::vector<uint32_t> hr(L);
	s.copy_to_host(hs);
	r.copy_to_host(hr);
	printf("{\"per_level_scale\": %.9g, \"base\": %u, \"levels\": [", per_level_scale, base);
	for (uint32_t l = 0; l < L; ++l) {
		const float host = grid_scale(l, l2, base);
		uint32_t db, hb;
		memcpy(&db, &hs[l], 4);
		memcpy(&hb, &host, 4);
		printf("%s{\"level\": %u, \"dev_scale\": %.9g, \"dev_bits\": %u, \"dev_res\": %u, \"host_scale\": %.9g, \"host_bits\": %u, \"host_res\": %u}", l ? ", " : "", l, hs[l], db, hr[l], host, hb, grid_resolution(host));
	}
	printf("]}\n");
	return 0;
}

struct Setup {
	json config;
	uint32_t n_in, n_out, B;
	TrainableModel model;
	GPUMatrix<float> x, y;
	std::vector<float> hx, hy;
};

static void make_setup(Setup& s, const std::string& config_path, uint32_t n_in, uint32_t n_out, uint32_t B, bool jit, uint32_t input_seed = 1337) {
	s.config = load_json(config_path);
	s.n_in = n_in;
	s.n_out = n_out;
	s.B = B;
	s.model = create_from_config(n_in, n_out, s.config);
	s.model.network->set_jit_fusion(jit && tcnn::supports_jit_fusion());
	s.x = GPUMatrix<float>(n_in, B);
	s.y = GPUMatrix<float>(n_out, B);
	default_rng_t rng{input_seed};
	generate_random_uniform<float>(nullptr, rng, (size_t)B * n_in, s.x.data());
	s.hx.resize((size_t)B * n_in);
	CUDA_CHECK_THROW(cudaMemcpy(s.hx.data(), s.x.data(), s.hx.size() * sizeof(float), cudaMemcpyDeviceToHost));
	make_targets(s.hx, n_in, n_out, B, s.hy);
	CUDA_CHECK_THROW(cudaMemcpy(s.y.data(), s.hy.data(), s.hy.size() * sizeof(float), cudaMemcpyHostToDevice));
}

// dump <config.json> <n_in> <n_out> <B> <n_steps> <outdir> <jit 0|1>
static int cmd_dump(int argc, char** argv) {
	if (argc < 9) {
		fprintf(stderr, "usage: dump config n_in n_out B n_steps outdir jit\n");
		return 2;
	}
	const std::string outdir = argv[7];
	mkdir(outdir.c_str(), 0755);
	Setup s;
	make_setup(s, argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[8]) != 0);
	const uint32_t n_steps = atoi(argv[6]);
	auto& trainer = s.model.trainer;
	auto& network = s.model.network;
	const size_t n_params = trainer->n_params();
	const uint32_t B = s.B;

	write_host_bin(outdir + "/x.f32", s.hx);
	write_host_bin(outdir + "/y.f32", s.hy);
	write_bin(outdir + "/params_init.f32", trainer->params_full_precision(), n_params);

	// Encoded features exactly as the MLP consumes them (padded width, preferred layout = SoA for grids).
	auto enc = network->encoding();
	GPUMatrixDynamic<precision_t> encoded{enc->padded_output_width(), B, nullptr, enc->preferred_output_layout()};
	enc->inference_mixed_precision(nullptr, s.x, encoded, true);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(outdir + "/encoded.f16", (const uint16_t*)encoded.data(), (size_t)enc->padded_output_width() * B);

	// fp32 inference output (n_out x B column-major == [B][n_out]).
	GPUMatrix<float> pred(s.n_out, B);
	network->inference(nullptr, s.x, pred);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(outdir + "/inference.f32", pred.data(), (size_t)s.n_out * B);

	// One fwd+bwd without optimizer: padded fp16 output, loss values, fp16 gradients.
	std::vector<float> losses;
	{
		auto ctx = trainer->training_step(nullptr, s.x, s.y, nullptr, /*run_optimizer=*/false);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		if (ctx->output.data()) {
			write_bin(outdir + "/output.f16", (const uint16_t*)ctx->output.data(), (size_t)network->padded_output_width() * B);
		}
		write_bin(outdir + "/loss_values.f32", ctx->L.data(), ctx->L.n_elements());
		write_bin(outdir + "/grads_step0.f16", (const uint16_t*)trainer->param_gradients(), n_params);
		losses.push_back(trainer->loss(nullptr, *ctx));
	}

	// n_steps full training steps on the same batch.
	for (uint32_t i = 0; i < n_steps; ++i) {
		auto ctx = trainer->training_step(nullptr, s.x, s.y);
		losses.push_back(trainer->loss(nullptr, *ctx));
		if (i == 0) {
			CUDA_CHECK_THROW(cudaDeviceSynchronize());
			write_bin(outdir + "/params_step1.f32", trainer->params_full_precision(), n_params);
		}
	}
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(outdir + "/params_final.f32", trainer->params_full_precision(), n_params);
	write_bin(outdir + "/params_final.f16", (const uint16_t*)trainer->params(), n_params);
	network->inference(nullptr, s.x, pred);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(outdir + "/inference_final.f32", pred.data(), (size_t)s.n_out * B);

	json meta;
	meta["config"] = s.config;
	meta["n_in"] = s.n_in;
	meta["n_out"] = s.n_out;
	meta["batch"] = B;
	meta["n_steps"] = n_steps;
	meta["n_params"] = n_params;
	meta["n_encoding_params"] = enc->n_params();
	meta["encoded_width"] = enc->padded_output_width();
	meta["encoded_layout"] = enc->preferred_output_layout() == SoA ? "SoA" : "AoS";
	meta["padded_output_width"] = network->padded_output_width();
	meta["jit_fusion"] = network->jit_fusion();
	meta["loss_n_elements"] = losses.empty() ? 0 : 1;
	meta["losses"] = losses;
	meta["hyperparams"] = network->hyperparams();
	std::ofstream f{outdir + "/meta.json"};
	f << meta.dump(1) << std::endl;
	printf("dumped %s: n_params=%zu losses[0]=%g losses[last]=%g jit=%d\n", outdir.c_str(), n_params, losses.front(), losses.back(), (int)network->jit_fusion());
	return 0;
}

// bench <config.json> <n_in> <n_out> <B> <steps> <warmup> <jit 0|1> [inference 0|1] [e2e 0|1]
// Trains on a pool of 4 batches drawn one after the other from pcg32{1337} (the same data sequence bench.py's own arm uses on
// rank 0), step i on batch i % 4. e2e = 1: every step copies the inputs and targets from PINNED HOST buffers to the device and
// reads the loss back -- the same end-to-end region bench.py times for the new library; wall-clock timed.
// "loss_after_steps" = the loss of training step number warmup + steps (0-based), i.e. of one more step after the timed loop.
static int cmd_bench(int argc, char** argv) {
	if (argc < 9) {
		fprintf(stderr, "usage: bench config n_in n_out B steps warmup jit [inference] [e2e]\n");
		return 2;
	}
	constexpr uint32_t POOL = 4;
	Setup s;
	make_setup(s, argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[8]) != 0);
	const uint32_t steps = atoi(argv[6]), warmup = atoi(argv[7]);
	const bool inference = argc > 9 && atoi(argv[9]) != 0;
	const bool e2e = argc > 10 && atoi(argv[10]) != 0;
	auto& trainer = s.model.trainer;
	auto& network = s.model.network;
	cudaStream_t stream;
	CUDA_CHECK_THROW(cudaStreamCreate(&stream));
	GPUMatrix<float> pred(s.n_out, s.B);

	// pool of device batches: one generator, consecutive draws (random.h:56-69 advances the caller's rng by n per call)
	std::vector<GPUMatrix<float>> xs, ys;
	std::vector<float*> hxs(POOL, nullptr), hys(POOL, nullptr);
	{
		default_rng_t rng{1337};
		std::vector<float> hx((size_t)s.B * s.n_in), hy;
		for (uint32_t i = 0; i < POOL; ++i) {
			xs.emplace_back(s.n_in, s.B);
			ys.emplace_back(s.n_out, s.B);
			generate_random_uniform<float>(nullptr, rng, (size_t)s.B * s.n_in, xs[i].data());
			CUDA_CHECK_THROW(cudaMemcpy(hx.data(), xs[i].data(), hx.size() * sizeof(float), cudaMemcpyDeviceToHost));
			make_targets(hx, s.n_in, s.n_out, s.B, hy);
			CUDA_CHECK_THROW(cudaMemcpy(ys[i].data(), hy.data(), hy.size() * sizeof(float), cudaMemcpyHostToDevice));
			if (e2e) {
				CUDA_CHECK_THROW(cudaMallocHost(&hxs[i], xs[i].n_bytes()));
				CUDA_CHECK_THROW(cudaMallocHost(&hys[i], ys[i].n_bytes()));
				memcpy(hxs[i], hx.data(), xs[i].n_bytes());
				memcpy(hys[i], hy.data(), ys[i].n_bytes());
			}
		}
	}
	float e2e_loss = 0;
	uint32_t it = 0;
	auto one = [&]() {
		const uint32_t b = it++ % POOL;
		if (e2e) {
			// staged into one device batch, as an application with host-side data does
			CUDA_CHECK_THROW(cudaMemcpyAsync(s.x.data(), hxs[b], s.x.n_bytes(), cudaMemcpyHostToDevice, stream));
			CUDA_CHECK_THROW(cudaMemcpyAsync(s.y.data(), hys[b], s.y.n_bytes(), cudaMemcpyHostToDevice, stream));
			auto ctx = trainer->training_step(stream, s.x, s.y);
			e2e_loss = trainer->loss(stream, *ctx);  // device -> host + stream synchronisation (trainer.h:372-378)
		} else if (inference) {
			network->inference(stream, xs[b], pred);
		} else {
			trainer->training_step(stream, xs[b], ys[b]);
		}
	};
	for (uint32_t i = 0; i < warmup; ++i) one();
	CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	auto w0 = std::chrono::steady_clock::now();
	cudaEventRecord(e0, stream);
	for (uint32_t i = 0; i < steps; ++i) one();
	cudaEventRecord(e1, stream);
	CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
	auto w1 = std::chrono::steady_clock::now();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double wall_ms = std::chrono::duration<double, std::milli>(w1 - w0).count();
	float final_loss = -1.0f;
	if (!inference) {
		const uint32_t b = it % POOL;
		auto ctx = trainer->training_step(stream, xs[b], ys[b]);
		final_loss = trainer->loss(stream, *ctx);
	}
	(void)e2e_loss;
	printf("{\"impl\": \"reference\", \"e2e\": %s, \"h2d_bytes_per_step\": %zu, \"mode\": \"%s\", \"jit_fusion\": %s, \"network\": \"%s\", \"batch\": %u, \"steps\": %u, \"warmup\": %u, \"ms_per_step\": %.6f, \"wall_ms_per_step\": %.6f, \"samples_per_s\": %.6e, \"n_params\": %zu, \"loss_after_steps\": %.6g, \"batch_pool\": %u}\n",
		e2e ? "true" : "false", e2e ? s.x.n_bytes() + s.y.n_bytes() : (size_t)0, inference ? "inference" : "training_step", network->jit_fusion() ? "true" : "false",
		s.config.value("network", json::object()).value("otype", "MLP").c_str(),
		s.B, steps, warmup, ms / steps, wall_ms / steps, (double)s.B * steps / (ms * 1e-3), trainer->n_params(), final_loss, POOL);
	return 0;
}

// Gather a strided subset of a device array: element k of the result = src[first + k * stride].
template <typename T>
static void write_strided(const std::string& path, const T* dev, size_t first, size_t n, size_t stride) {
	std::vector<T> h(n);
	CUDA_CHECK_THROW(cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
	std::vector<T> out;
	for (size_t i = first; i < n; i += stride) out.push_back(h[i]);
	write_host_bin(path, out);
}

// dumpbig <config.json> <n_in> <n_out> <B> <n_steps> <outdir> <jit 0|1> <stride> <n_head>
// Golden vectors at the BENCHMARKED size (T = 2^19, B = 2^16 .. 2^18) in a form small enough to commit: inputs are regenerated by
// the test from the seed (sums are recorded here to check that), per-sample outputs only for the first n_head samples,
// parameter-sized arrays as [all network weights | every stride-th grid parameter], counts of non-zero gradients / moved
// parameters for the touched-set comparison, and the whole loss trajectory.
static int cmd_dumpbig(int argc, char** argv) {
	if (argc < 11) {
		fprintf(stderr, "usage: dumpbig config n_in n_out B n_steps outdir jit stride n_head\n");
		return 2;
	}
	const std::string outdir = argv[7];
	mkdir(outdir.c_str(), 0755);
	Setup s;
	make_setup(s, argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[8]) != 0);
	const uint32_t n_steps = atoi(argv[6]);
	const size_t stride = atoi(argv[9]);
	const uint32_t n_head = atoi(argv[10]);
	auto& trainer = s.model.trainer;
	auto& network = s.model.network;
	const size_t n_params = trainer->n_params();
	const uint32_t B = s.B;
	auto enc = network->encoding();
	const size_t n_net = n_params - enc->n_params();  // network weights come first (network_with_input_encoding.h:115-130)

	auto write_param_sample = [&](const std::string& stem, auto* dev) {
		using T = std::remove_cv_t<std::remove_pointer_t<decltype(dev)>>;
		std::vector<T> h(n_params);
		CUDA_CHECK_THROW(cudaMemcpy(h.data(), dev, n_params * sizeof(T), cudaMemcpyDeviceToHost));
		std::vector<T> out(h.begin(), h.begin() + n_net);
		for (size_t i = n_net; i < n_params; i += stride) out.push_back(h[i]);
		write_host_bin(outdir + "/" + stem, out);
		return h;
	};

	double sum_x = 0, sum_y = 0;
	for (float v : s.hx) sum_x += v;
	for (float v : s.hy) sum_y += v;
	auto p_init = write_param_sample("params_init.f32", (const float*)trainer->params_full_precision());

	// encoded features of the first n_head samples (SoA [width][B] -> [width][n_head])
	GPUMatrixDynamic<precision_t> encoded{enc->padded_output_width(), B, nullptr, enc->preferred_output_layout()};
	enc->inference_mixed_precision(nullptr, s.x, encoded, true);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	{
		std::vector<uint16_t> h((size_t)enc->padded_output_width() * B), out;
		CUDA_CHECK_THROW(cudaMemcpy(h.data(), encoded.data(), h.size() * 2, cudaMemcpyDeviceToHost));
		const bool soa = enc->preferred_output_layout() == SoA;
		for (uint32_t f = 0; f < enc->padded_output_width(); ++f)
			for (uint32_t i = 0; i < n_head; ++i) out.push_back(soa ? h[(size_t)f * B + i] : h[(size_t)i * enc->padded_output_width() + f]);
		write_host_bin(outdir + "/encoded_head.f16", out);
	}
	GPUMatrix<float> pred(s.n_out, B);
	network->inference(nullptr, s.x, pred);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(outdir + "/inference_head.f32", pred.data(), (size_t)s.n_out * n_head);

	std::vector<float> losses;
	size_t n_grad_nonzero = 0;
	{
		auto ctx = trainer->training_step(nullptr, s.x, s.y, nullptr, /*run_optimizer=*/false);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		if (ctx->output.data()) write_bin(outdir + "/output_head.f16", (const uint16_t*)ctx->output.data(), (size_t)network->padded_output_width() * n_head);
		write_bin(outdir + "/loss_values_head.f32", ctx->L.data(), (size_t)network->padded_output_width() * n_head);
		auto g = write_param_sample("grads_step0.f16", (const uint16_t*)trainer->param_gradients());
		for (size_t i = n_net; i < n_params; ++i) n_grad_nonzero += (g[i] & 0x7FFFu) != 0;
		losses.push_back(trainer->loss(nullptr, *ctx));
	}
	size_t n_moved = 0;
	for (uint32_t i = 0; i < n_steps; ++i) {
		auto ctx = trainer->training_step(nullptr, s.x, s.y);
		losses.push_back(trainer->loss(nullptr, *ctx));
		if (i == 0) {
			CUDA_CHECK_THROW(cudaDeviceSynchronize());
			auto p1 = write_param_sample("params_step1.f32", (const float*)trainer->params_full_precision());
			for (size_t k = n_net; k < n_params; ++k) n_moved += p1[k] != p_init[k];
		}
	}
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_param_sample("params_final.f16", (const uint16_t*)trainer->params());
	network->inference(nullptr, s.x, pred);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(outdir + "/inference_final_head.f32", pred.data(), (size_t)s.n_out * n_head);

	json meta;
	meta["config"] = s.config;
	meta["n_in"] = s.n_in;
	meta["n_out"] = s.n_out;
	meta["batch"] = B;
	meta["n_steps"] = n_steps;
	meta["n_params"] = n_params;
	meta["n_network_params"] = n_net;
	meta["stride"] = stride;
	meta["n_head"] = n_head;
	meta["input_seed"] = 1337;
	meta["sum_x"] = sum_x;
	meta["sum_y"] = sum_y;
	meta["n_grid_grad_nonzero"] = n_grad_nonzero;
	meta["n_grid_params_moved_step1"] = n_moved;
	meta["encoded_width"] = enc->padded_output_width();
	meta["padded_output_width"] = network->padded_output_width();
	meta["jit_fusion"] = network->jit_fusion();
	meta["losses"] = losses;
	std::ofstream f{outdir + "/meta.json"};
	f << meta.dump(1) << std::endl;
	printf("dumpbig %s: n_params=%zu B=%u losses[0]=%g losses[last]=%g nonzero grid grads=%zu\n", outdir.c_str(), n_params, B, losses.front(), losses.back(), n_grad_nonzero);
	return 0;
}

// ---- the network on its own (benchmarks/mlp/bench_mlp_ours.cu): tcnn::create_network<T>(json), fp16 inputs and outputs ----------
struct NetSetup {
	std::shared_ptr<Network<precision_t>> network;
	std::shared_ptr<Trainer<precision_t, precision_t, precision_t>> trainer;
	uint32_t n_in, n_out;
};

static void make_net(NetSetup& s, const char* otype, uint32_t width, uint32_t hidden, uint32_t n_in, uint32_t n_out, bool jit, const char* act = "ReLU", const char* out_act = "None") {
	json opts = {{"otype", otype}, {"n_input_dims", n_in}, {"n_output_dims", n_out}, {"n_neurons", width}, {"n_hidden_layers", hidden}, {"activation", act}, {"output_activation", out_act}};
	std::shared_ptr<Loss<precision_t>> loss{create_loss<precision_t>(json::object())};
	std::shared_ptr<Optimizer<precision_t>> optimizer{create_optimizer<precision_t>(json::object())};
	s.network.reset(create_network<precision_t>(opts));
	s.network->set_jit_fusion(jit && tcnn::supports_jit_fusion());
	s.trainer = std::make_shared<Trainer<precision_t, precision_t, precision_t>>(s.network, optimizer, loss);  // owns + initialises the parameters
	s.n_in = n_in;
	s.n_out = n_out;
}

// mlpdump <otype> <width> <hidden> <n_in> <n_out> <B> <outdir> <jit> [activation] [output_activation]
static int cmd_mlpdump(int argc, char** argv) {
	if (argc < 10) {
		fprintf(stderr, "usage: mlpdump otype width hidden n_in n_out B outdir jit [activation] [output_activation]\n");
		return 2;
	}
	NetSetup s;
	const uint32_t B = atoi(argv[7]);
	make_net(s, argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[9]) != 0, argc > 10 ? argv[10] : "ReLU", argc > 11 ? argv[11] : "None");
	const std::string outdir = argv[8];
	mkdir(outdir.c_str(), 0755);
	default_rng_t rng{1337};
	GPUMatrixDynamic<precision_t> x(s.n_in, B, CM);  // column-major n_in x B == [B][n_in]
	generate_random_uniform<precision_t>(nullptr, rng, (size_t)B * s.n_in, x.data());
	const uint32_t out_w = s.network->padded_output_width();
	GPUMatrixDynamic<precision_t> y(out_w, B, CM);
	s.network->inference_mixed_precision(nullptr, x, y);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(outdir + "/x.f16", (const uint16_t*)x.data(), (size_t)B * s.n_in);
	write_bin(outdir + "/output.f16", (const uint16_t*)y.data(), (size_t)B * out_w);
	write_bin(outdir + "/params.f16", (const uint16_t*)s.trainer->params(), s.trainer->n_params());
	json meta;
	meta["otype"] = argv[2];
	meta["width"] = atoi(argv[3]);
	meta["n_hidden_layers"] = atoi(argv[4]);
	meta["n_in"] = s.n_in;
	meta["n_out"] = s.n_out;
	meta["padded_output_width"] = out_w;
	meta["batch"] = B;
	meta["n_params"] = s.trainer->n_params();
	meta["jit_fusion"] = s.network->jit_fusion();
	meta["activation"] = argc > 10 ? argv[10] : "ReLU";
	meta["output_activation"] = argc > 11 ? argv[11] : "None";
	std::ofstream f{outdir + "/meta.json"};
	f << meta.dump(1) << std::endl;
	printf("mlpdump %s: n_params=%zu out_w=%u\n", outdir.c_str(), s.trainer->n_params(), out_w);
	return 0;
}

// mlpbench <otype> <width> <hidden> <n_in> <n_out> <B> <iters> <warmup> <jit>: inference_mixed_precision throughput (CUDA events)
static int cmd_mlpbench(int argc, char** argv) {
	if (argc < 11) {
		fprintf(stderr, "usage: mlpbench otype width hidden n_in n_out B iters warmup jit\n");
		return 2;
	}
	NetSetup s;
	const uint32_t B = atoi(argv[7]), iters = atoi(argv[8]), warmup = atoi(argv[9]);
	const bool jit = atoi(argv[10]) != 0;
	make_net(s, argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), jit);
	cudaStream_t stream;
	CUDA_CHECK_THROW(cudaStreamCreate(&stream));
	default_rng_t rng{1337};
	// "Most efficient in RM layout when used with JIT, CM layout otherwise" (bench_mlp_ours.cu:81)
	GPUMatrixDynamic<precision_t> x(s.n_in, B, jit ? RM : CM);
	GPUMatrix<precision_t, RM> y(s.network->padded_output_width(), B);
	generate_random_uniform<precision_t>(stream, rng, (size_t)B * s.n_in, x.data());
	for (uint32_t i = 0; i < warmup; ++i) s.network->inference_mixed_precision(stream, x, y);
	CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0, stream);
	for (uint32_t i = 0; i < iters; ++i) s.network->inference_mixed_precision(stream, x, y);
	cudaEventRecord(e1, stream);
	CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	printf("{\"impl\": \"reference\", \"mode\": \"mlp_inference\", \"otype\": \"%s\", \"jit_fusion\": %s, \"width\": %d, \"n_hidden_layers\": %d, \"n_in\": %u, \"n_out\": %u, \"batch\": %u, \"iters\": %u, \"ms_per_batch\": %.6f, \"samples_per_s\": %.6e}\n",
		argv[2], s.network->jit_fusion() ? "true" : "false", atoi(argv[3]), atoi(argv[4]), s.n_in, s.n_out, B, iters, ms / iters, (double)B * iters / (ms * 1e-3));
	return 0;
}

// ---- snapshot interoperability (trainer.h:442-482): the reference's own serialize() / deserialize(), stored as msgpack like instant-ngp does
// snapshot save <config> <n_in> <n_out> <B> <n_steps> <outdir> <with_optimizer>: train n_steps, write snapshot.msgpack, x.f32, y.f32, inference.f32
// snapshot load <config> <n_in> <n_out> <B> <dir> <file.msgpack>: deserialize, inference on dir/x.f32 -> dir/inference_loaded.f32, one more
//                training step on (x, y) -> prints its loss (exercises the restored optimizer state)
static std::vector<float> read_f32(const std::string& path) {
	std::ifstream f{path, std::ios::binary | std::ios::ate};
	if (!f) throw std::runtime_error{"cannot open " + path};
	std::vector<float> v((size_t)f.tellg() / 4);
	f.seekg(0);
	f.read((char*)v.data(), v.size() * 4);
	return v;
}

static int cmd_snapshot(int argc, char** argv) {
	if (argc < 9) {
		fprintf(stderr, "usage: snapshot save config n_in n_out B n_steps outdir with_optimizer | snapshot load config n_in n_out B dir file\n");
		return 2;
	}
	const std::string mode = argv[2];
	Setup s;
	make_setup(s, argv[3], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), false);
	auto& trainer = s.model.trainer;
	auto& network = s.model.network;
	GPUMatrix<float> pred(s.n_out, s.B);
	if (mode == "save") {
		const std::string outdir = argv[8];
		mkdir(outdir.c_str(), 0755);
		float loss = 0;
		for (int i = 0; i < atoi(argv[7]); ++i) {
			auto ctx = trainer->training_step(nullptr, s.x, s.y);
			loss = trainer->loss(nullptr, *ctx);
		}
		const json snap = trainer->serialize(argc > 9 && atoi(argv[9]) != 0);
		const std::vector<uint8_t> bytes = json::to_msgpack(snap);
		std::ofstream f{outdir + "/snapshot.msgpack", std::ios::binary};
		f.write((const char*)bytes.data(), bytes.size());
		network->inference(nullptr, s.x, pred);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		write_host_bin(outdir + "/x.f32", s.hx);
		write_host_bin(outdir + "/y.f32", s.hy);
		write_bin(outdir + "/inference.f32", pred.data(), (size_t)s.n_out * s.B);
		printf("{\"saved\": \"%s/snapshot.msgpack\", \"bytes\": %zu, \"last_loss\": %.6g}\n", outdir.c_str(), bytes.size(), loss);
		return 0;
	}
	const std::string dir = argv[7];
	std::ifstream f{argv[8], std::ios::binary};
	if (!f) throw std::runtime_error{std::string{"cannot open "} + argv[8]};
	const std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	trainer->deserialize(json::from_msgpack(bytes));
	const std::vector<float> hx = read_f32(dir + "/x.f32"), hy = read_f32(dir + "/y.f32");
	CUDA_CHECK_THROW(cudaMemcpy(s.x.data(), hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
	CUDA_CHECK_THROW(cudaMemcpy(s.y.data(), hy.data(), hy.size() * 4, cudaMemcpyHostToDevice));
	network->inference(nullptr, s.x, pred);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_bin(dir + "/inference_loaded.f32", pred.data(), (size_t)s.n_out * s.B);
	auto ctx = trainer->training_step(nullptr, s.x, s.y);
	printf("{\"loaded\": \"%s\", \"next_step_loss\": %.6g}\n", argv[8], trainer->loss(nullptr, *ctx));
	return 0;
}

int main(int argc, char** argv) {
	try {
		if (argc < 2) {
			fprintf(stderr, "usage: %s dump|dumpbig|mlpdump|mlpbench|bench|probe ...\n", argv[0]);
			return 2;
		}
		const std::string cmd = argv[1];
		if (cmd == "dump") return cmd_dump(argc, argv);
		if (cmd == "bench") return cmd_bench(argc, argv);
		if (cmd == "dumpbig") return cmd_dumpbig(argc, argv);
		if (cmd == "mlpdump") return cmd_mlpdump(argc, argv);
		if (cmd == "snapshot") return cmd_snapshot(argc, argv);
		if (cmd == "mlpbench") return cmd_mlpbench(argc, argv);
		if (cmd == "probe") return cmd_probe(argc, argv);
		fprintf(stderr, "unknown command %s\n", cmd.c_str());
		return 2;
	} catch (const std::exception& e) {
		fprintf(stderr, "ref_harness: uncaught exception: %s\n", e.what());
		return 1;
	}
}
