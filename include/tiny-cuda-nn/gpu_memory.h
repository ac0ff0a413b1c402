/* tiny-cuda-nn/gpu_memory.h -- GPUMemory<T> with the members applications use (reference: gpu_memory.h:60-330). */
#pragma once
#include <vector>

#include "common.h"

namespace tcnn {

template <typename T>
class GPUMemory {
public:
	GPUMemory() = default;
	explicit GPUMemory(size_t size) { resize(size); }
	GPUMemory(const GPUMemory&) = delete;
	GPUMemory& operator=(const GPUMemory&) = delete;
	GPUMemory(GPUMemory&& o) noexcept : m_data{o.m_data}, m_size{o.m_size} { o.m_data = nullptr; o.m_size = 0; }
	GPUMemory& operator=(GPUMemory&& o) noexcept {
		if (this != &o) { free_memory(); m_data = o.m_data; m_size = o.m_size; o.m_data = nullptr; o.m_size = 0; }
		return *this;
	}
	~GPUMemory() { try { free_memory(); } catch (...) {} }

	void resize(size_t size) {
		if (size == m_size) return;
		free_memory();
		if (size) CUDA_CHECK_THROW(cudaMalloc((void**)&m_data, size * sizeof(T)));
		m_size = size;
	}
	void enlarge(size_t size) { if (size > m_size) resize(size); }
	void memset(int value) { if (m_data) CUDA_CHECK_THROW(cudaMemset(m_data, value, m_size * sizeof(T))); }
	void free_memory() {
		if (m_data) CUDA_CHECK_THROW(cudaFree(m_data));
		m_data = nullptr;
		m_size = 0;
	}
	void copy_from_host(const T* host_data, size_t num_elements) { CUDA_CHECK_THROW(cudaMemcpy(m_data, host_data, num_elements * sizeof(T), cudaMemcpyHostToDevice)); }
	void copy_from_host(const T* host_data) { copy_from_host(host_data, m_size); }
	void copy_from_host(const std::vector<T>& data) { copy_from_host(data.data(), data.size()); }
	void resize_and_copy_from_host(const std::vector<T>& data) { resize(data.size()); copy_from_host(data); }
	void copy_to_host(T* host_data, size_t num_elements) const { CUDA_CHECK_THROW(cudaMemcpy(host_data, m_data, num_elements * sizeof(T), cudaMemcpyDeviceToHost)); }
	void copy_to_host(T* host_data) const { copy_to_host(host_data, m_size); }
	void copy_to_host(std::vector<T>& data) const { data.resize(m_size); copy_to_host(data.data(), m_size); }

	T* data() const { return m_data; }
	size_t size() const { return m_size; }
	size_t get_num_elements() const { return m_size; }
	size_t bytes() const { return m_size * sizeof(T); }

private:
	T* m_data = nullptr;
	size_t m_size = 0;
};

}  // namespace tcnn
