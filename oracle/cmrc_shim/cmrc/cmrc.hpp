// Minimal stand-in for the CMakeRC ("cmrc") embedded-resource API, written for this repo.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md). The upstream build embeds its public headers into the
// binary through the CMakeRC CMake module so that NVRTC can resolve `#include <tiny-cuda-nn/...>` at run
// time (reference CMakeLists.txt:347-381, consumer src/rtc_kernel.cu:98-113,211-213). This repo may not run
// the reference's CMake build, so oracle/Makefile generates the same embedding with oracle/gen_resources.py
// and this header supplies the three calls the consumer makes: iterate_directory / open / get_filesystem.
// The listing is flat: iterate_directory("") returns every file with its full relative path.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace cmrc {

struct file {
	const char* b = nullptr;
	const char* e = nullptr;
	const char* begin() const { return b; }
	const char* end() const { return e; }
	size_t size() const { return (size_t)(e - b); }
};

struct directory_entry {
	std::string name;
	const std::string& filename() const { return name; }
	bool is_file() const { return true; }
	bool is_directory() const { return false; }
};

class embedded_filesystem {
public:
	struct item { const char* path; const char* data; size_t size; };
	embedded_filesystem(const item* items, size_t n) : m_items{items}, m_n{n} {}

	std::vector<directory_entry> iterate_directory(const std::string& dir) const {
		std::vector<directory_entry> out;
		if (!dir.empty()) return out;
		for (size_t i = 0; i < m_n; ++i) out.push_back({m_items[i].path});
		return out;
	}

	file open(const std::string& path) const {
		for (size_t i = 0; i < m_n; ++i) {
			if (path == m_items[i].path) return {m_items[i].data, m_items[i].data + m_items[i].size};
		}
		return {};
	}

private:
	const item* m_items;
	size_t m_n;
};

}

#define CMRC_DECLARE(ns) namespace cmrc { namespace ns { cmrc::embedded_filesystem get_filesystem(); } }
