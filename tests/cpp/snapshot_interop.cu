// snapshot_interop -- trainer->serialize() / deserialize() in the reference's JSON schema (include/tiny-cuda-nn/config.h over the C ABI),
// exchanged with the UNMODIFIED reference as msgpack files (tests/test_cpp_shim.py drives both sides):
//   snapshot_interop load <config.json> <n_in> <n_out> <B> <dir> <snapshot.msgpack>   deserialize a snapshot, inference on dir/x.f32 ->
//                                                                                     dir/inference_loaded.f32, one more training step
//   snapshot_interop save <config.json> <n_in> <n_out> <B> <dir> <n_steps> <with_opt> train on dir/x.f32, dir/y.f32, write
//                                                                                     dir/snapshot_b200.msgpack + dir/inference_b200.f32
#include <tiny-cuda-nn/config.h>

#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

using namespace tcnn;

static std::vector<float> read_f32(const std::string& path) {
	std::ifstream f{path, std::ios::binary | std::ios::ate};
	if (!f) throw std::runtime_error{"cannot open " + path};
	std::vector<float> v((size_t)f.tellg() / 4);
	f.seekg(0);
	f.read((char*)v.data(), v.size() * 4);
	return v;
}

static void write_f32(const std::string& path, const float* dev, size_t n) {
	std::vector<float> h(n);
	CUDA_CHECK_THROW(cudaMemcpy(h.data(), dev, n * 4, cudaMemcpyDeviceToHost));
	std::ofstream f{path, std::ios::binary};
	f.write((const char*)h.data(), n * 4);
}

int main(int argc, char** argv) {
	try {
		if (argc < 8) {
			fprintf(stderr, "usage: snapshot_interop load|save config n_in n_out B dir ...\n");
			return 2;
		}
		const std::string mode = argv[1], dir = argv[6];
		std::ifstream cf{argv[2]};
		const json config = json::parse(cf, nullptr, true, true);
		const uint32_t n_in = atoi(argv[3]), n_out = atoi(argv[4]), B = atoi(argv[5]);
		auto model = create_from_config(n_in, n_out, config);
		GPUMatrix<float> x(n_in, B), y(n_out, B), pred(n_out, B);
		const std::vector<float> hx = read_f32(dir + "/x.f32"), hy = read_f32(dir + "/y.f32");
		CUDA_CHECK_THROW(cudaMemcpy(x.data(), hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
		CUDA_CHECK_THROW(cudaMemcpy(y.data(), hy.data(), hy.size() * 4, cudaMemcpyHostToDevice));
		if (mode == "load") {
			std::ifstream f{argv[7], std::ios::binary};
			if (!f) throw std::runtime_error{std::string{"cannot open "} + argv[7]};
			const std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
			const json snap = json::from_msgpack(bytes);
			model.trainer->deserialize(snap);
			model.network->inference(nullptr, x, pred);
			CUDA_CHECK_THROW(cudaDeviceSynchronize());
			write_f32(dir + "/inference_loaded.f32", pred.data(), (size_t)n_out * B);
			auto ctx = model.trainer->training_step(nullptr, x, y);
			printf("{\"loaded\": \"%s\", \"has_optimizer\": %s, \"next_step_loss\": %.6g}\n", argv[7], snap.contains("optimizer") ? "true" : "false", model.trainer->loss(nullptr, *ctx));
		} else {
			float loss = 0;
			for (int i = 0; i < atoi(argv[7]); ++i) {
				auto ctx = model.trainer->training_step(nullptr, x, y);
				loss = model.trainer->loss(nullptr, *ctx);
			}
			const json snap = model.trainer->serialize(argc > 8 && atoi(argv[8]) != 0);
			const std::vector<uint8_t> bytes = json::to_msgpack(snap);
			std::ofstream f{dir + "/snapshot_b200.msgpack", std::ios::binary};
			f.write((const char*)bytes.data(), bytes.size());
			model.network->inference(nullptr, x, pred);
			CUDA_CHECK_THROW(cudaDeviceSynchronize());
			write_f32(dir + "/inference_b200.f32", pred.data(), (size_t)n_out * B);
			printf("{\"saved\": \"%s/snapshot_b200.msgpack\", \"bytes\": %zu, \"last_loss\": %.6g}\n", dir.c_str(), bytes.size(), loss);
		}
		return 0;
	} catch (const std::exception& e) {
		fprintf(stderr, "snapshot_interop: %s\n", e.what());
		return 1;
	}
}
