// CPU dry-run stand-in for cub::DeviceSelect (tests/emu).
#pragma once
#include <cuda_runtime.h>

namespace cub {
struct DeviceSelect {
  template <class In, class Flag, class Out, class Num>
  static cudaError_t Flagged(void* d_temp, size_t& temp_bytes, In in, Flag flags, Out out, Num num_out, int64_t n,
                             cudaStream_t = nullptr) {
    if (!d_temp) { temp_bytes = 16; return cudaSuccess; }
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i)
      if (flags[i]) out[k++] = in[i];
    *num_out = k;
    return cudaSuccess;
  }
};
}  // namespace cub
