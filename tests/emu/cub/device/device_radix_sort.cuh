// CPU dry-run stand-in for cub::DeviceRadixSort (tests/emu): a stable sort on the key bits.
#pragma once
#include <cuda_runtime.h>

#include <numeric>

namespace cub {
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* d_temp, size_t& temp_bytes, const K* keys_in, K* keys_out, const V* vals_in,
                               V* vals_out, uint64_t n, int begin_bit, int end_bit, cudaStream_t = nullptr) {
    if (!d_temp) { temp_bytes = 16; return cudaSuccess; }
    const K mask = end_bit >= int(sizeof(K) * 8) ? ~K(0) : ((K(1) << end_bit) - 1);
    std::vector<uint64_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
      return ((keys_in[a] & mask) >> begin_bit) < ((keys_in[b] & mask) >> begin_bit);
    });
    for (uint64_t i = 0; i < n; ++i) { keys_out[i] = keys_in[idx[i]]; vals_out[i] = vals_in[idx[i]]; }
    return cudaSuccess;
  }
};
}  // namespace cub
