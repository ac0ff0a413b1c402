// Application-style use of the tiny-cuda-nn C++ names over libtcnn_b200 (include/tiny-cuda-nn/*.h), following the flow of the
// reference's samples/mlp_learning_an_image.cu:213-300 with a closed-form "image" instead of a JPEG: separate loss / optimizer /
// network objects tied together by a Trainer, training batches drawn with generate_random_uniform, periodic loss read-back,
// full-frame inference. Prints one JSON line that tests/test_cpp_shim.py checks.
#include <tiny-cuda-nn/config.h>

#include <cstdio>
#include <vector>

using namespace tcnn;
using precision_t = network_precision_t;

template <uint32_t stride>
__global__ void eval_field(uint32_t n_elements, const float* __restrict__ xs_and_ys, float* __restrict__ result) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const float x = xs_and_ys[i * 2 + 0], y = xs_and_ys[i * 2 + 1];
	const float v[3] = {0.5f + 0.5f * __sinf(18.0f * x) * __cosf(11.0f * y), x * y, 0.5f + 0.5f * __cosf(25.0f * (x - y))};
	for (uint32_t c = 0; c < stride; ++c) result[i * stride + c] = v[c % 3];
}

int main(int argc, char* argv[]) {
	try {
		const uint32_t n_training_steps = argc >= 2 ? atoi(argv[1]) : 300;
		json config = {
			{"loss", {{"otype", "RelativeL2"}}},
			{"optimizer", {{"otype", "Adam"}, {"learning_rate", 1e-2}, {"beta1", 0.9}, {"beta2", 0.99}, {"epsilon", 1e-15}, {"l2_reg", 1e-6}}},
			{"encoding", {{"otype", "HashGrid"}, {"n_levels", 16}, {"n_features_per_level", 2}, {"log2_hashmap_size", 15}, {"base_resolution", 16}, {"per_level_scale", 1.5}}},
			{"network", {{"otype", "FullyFusedMLP"}, {"activation", "ReLU"}, {"output_activation", "None"}, {"n_neurons", 64}, {"n_hidden_layers", 2}}},
		};

		const uint32_t batch_size = 1 << 16, n_input_dims = 2, n_output_dims = 3;
		const uint32_t width = 200, height = 120, n_coords = width * height, n_coords_padded = next_multiple(n_coords, BATCH_SIZE_GRANULARITY);

		cudaStream_t stream;
		CUDA_CHECK_THROW(cudaStreamCreate(&stream));
		default_rng_t rng{1337};

		GPUMatrix<float> training_target(n_output_dims, batch_size);
		GPUMatrix<float> training_batch(n_input_dims, batch_size);
		GPUMemory<float> xs_and_ys(n_coords_padded * 2);
		std::vector<float> host_xy(n_coords_padded * 2, 0.5f);
		for (uint32_t y = 0; y < height; ++y)
			for (uint32_t x = 0; x < width; ++x) {
				host_xy[(y * width + x) * 2 + 0] = (x + 0.5f) / width;
				host_xy[(y * width + x) * 2 + 1] = (y + 0.5f) / height;
			}
		xs_and_ys.copy_from_host(host_xy.data());
		GPUMatrix<float> prediction(n_output_dims, n_coords_padded);
		GPUMatrix<float> inference_batch(xs_and_ys.data(), n_input_dims, n_coords_padded);
		GPUMemory<float> truth(n_coords_padded * 3);
		linear_kernel(eval_field<3>, 0, stream, n_coords_padded, xs_and_ys.data(), truth.data());

		json encoding_opts = config.value("encoding", json::object()), loss_opts = config.value("loss", json::object());
		json optimizer_opts = config.value("optimizer", json::object()), network_opts = config.value("network", json::object());
		std::shared_ptr<Loss<precision_t>> loss{create_loss<precision_t>(loss_opts)};
		std::shared_ptr<Optimizer<precision_t>> optimizer{create_optimizer<precision_t>(optimizer_opts)};
		auto network = std::make_shared<NetworkWithInputEncoding<precision_t>>(n_input_dims, n_output_dims, encoding_opts, network_opts);
		network->set_jit_fusion(tcnn::supports_jit_fusion());
		auto trainer = std::make_shared<Trainer<float, precision_t, precision_t>>(network, optimizer, loss);

		// the first numbers of the first training batch, for comparison with the reference's generator
		default_rng_t probe{1337};
		GPUMemory<float> first(8);
		generate_random_uniform<float>(stream, probe, 8, first.data());
		std::vector<float> first_host(8);
		CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
		first.copy_to_host(first_host);

		float first_loss = 0, last_loss = 0;
		for (uint32_t i = 0; i < n_training_steps; ++i) {
			generate_random_uniform<float>(stream, rng, batch_size * n_input_dims, training_batch.data());
			linear_kernel(eval_field<n_output_dims>, 0, stream, batch_size, training_batch.data(), training_target.data());
			auto ctx = trainer->training_step(stream, training_batch, training_target);
			if (i == 0) first_loss = trainer->loss(stream, *ctx);
			if (i + 1 == n_training_steps) last_loss = trainer->loss(stream, *ctx);
		}
		network->inference(stream, inference_batch, prediction);
		CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
		std::vector<float> pred(n_coords_padded * 3), ref(n_coords_padded * 3);
		CUDA_CHECK_THROW(cudaMemcpy(pred.data(), prediction.data(), pred.size() * sizeof(float), cudaMemcpyDeviceToHost));
		truth.copy_to_host(ref);
		double mse = 0;
		for (size_t i = 0; i < (size_t)n_coords * 3; ++i) mse += (pred[i] - ref[i]) * (double)(pred[i] - ref[i]);
		mse /= (double)n_coords * 3;

		// second flavour: create_from_config (config.h:53-63) and one step on the same data
		TrainableModel model = create_from_config(n_input_dims, n_output_dims, config);
		auto ctx2 = model.trainer->training_step(stream, training_batch, training_target);
		const float loss2 = model.trainer->loss(stream, *ctx2);

		printf("{\"n_params\": %zu, \"padded_output_width\": %u, \"first_loss\": %g, \"last_loss\": %g, \"inference_mse\": %g, \"create_from_config_loss\": %g, \"first_uniform\": [", trainer->n_params(),
		       network->padded_output_width(), first_loss, last_loss, mse, loss2);
		for (int i = 0; i < 8; ++i) printf("%s%.9g", i ? ", " : "", first_host[i]);
		printf("]}\n");
		return 0;
	} catch (const std::exception& e) {
		fprintf(stderr, "shim_sample failed: %s\n", e.what());
		return 1;
	}
}
