// binning.h -- launch interface of binning.cu (per-step counting sort of the batch by (y, z) column).
#pragma once
#include "common.cuh"

namespace tcnnb {

// log2 of the per-axis bin resolution R chosen for `n_samples` samples (about 16 samples per bin).
uint32_t binning_log2_resolution(uint32_t n_samples, uint32_t n_pos_dims);
uint32_t binning_n_bins(uint32_t log2_r, uint32_t n_pos_dims);

// pos [n][D] -> perm[binned index] = original sample index.
// keys: scratch [2 * n] (bin, rank-in-bin per sample); hist: scratch [2 * n_bins], whose first n_bins words must be ZERO on
// entry (they are left zero on exit, so one memset at allocation time suffices). 3 launches on `stream`.
// The first launch also zero-fills `zero_bytes` at `zero_ptr` (both multiples of 16; the step's gradient table) and `*zero_scalar`.
cudaError_t launch_binning(cudaStream_t stream, uint32_t n_pos_dims, uint32_t n, const float* pos, uint32_t log2_r, uint32_t* keys, uint32_t* hist, uint32_t* perm, void* zero_ptr = nullptr, size_t zero_bytes = 0, float* zero_scalar = nullptr);

}  // namespace tcnnb
