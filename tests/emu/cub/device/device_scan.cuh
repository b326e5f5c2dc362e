// CPU dry-run stand-in for cub::DeviceScan (tests/emu).
#pragma once
#include <cuda_runtime.h>

namespace cub {
struct DeviceScan {
  template <class In, class Out, class Op>
  static cudaError_t InclusiveScan(void* d_temp, size_t& temp_bytes, In in, Out out, Op op, int64_t n, cudaStream_t = nullptr) {
    if (!d_temp) { temp_bytes = 16; return cudaSuccess; }
    for (int64_t i = 0; i < n; ++i) out[i] = i ? op(out[i - 1], in[i]) : in[i];
    return cudaSuccess;
  }
};
}  // namespace cub
