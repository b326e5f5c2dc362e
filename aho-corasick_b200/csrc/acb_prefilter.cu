// acb_prefilter.cu -- K3/K3b: position-parallel k-gram prefilter fused with the
// anchored DFA verify, and the non-overlapping chain resolution.
//
// Why this shape on B200: the reference's per-byte loop (src/automaton.rs:1310-1418,
// 1491-1534) is one dependent table load per haystack byte -- latency bound, and
// on a GPU it forces one lane per shard with strided haystack reads.  Testing
// every *start position* independently instead reads the haystack exactly once,
// fully coalesced (16 B per lane), costs one shared-memory probe per position and
// leaves the dependent DFA walk to the few positions that survive (the role the
// packed/Teddy prefilter plays in the reference, src/packed/teddy/README.md).
// Exactness comes from the verifier, which walks the shipped DFA from the
// candidate offset while the state stays on the trie path that starts there
// (depth(state) == bytes consumed) and reports the node's own patterns -- the set
// of patterns that are a prefix of the haystack at that offset.
#include "acb_device.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

namespace acb {
namespace {

constexpr int kPfThreads = 512;
constexpr int kPfQcap = 4096;   // candidate queue entries per CTA
constexpr int kPfMcap = 512;    // staged match tuples per CTA

__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

struct PfShared {
  unsigned long long m_base;
  unsigned long long q_base;  // absolute offset that queue entries are relative to
  unsigned int q_count;
  unsigned int q_tile_mark;   // q_count at the start of the current tile
  unsigned int m_count;
  unsigned int overflow;
  unsigned long long cand_total;
  uint8_t cls[256];
};

struct Emitter {
  uint64_t* g_keys;
  uint32_t* g_pids;
  unsigned long long* g_counter;
  uint64_t cap;
  uint64_t* s_keys;
  uint32_t* s_pids;
  unsigned int* s_count;
  __device__ __forceinline__ void emit(uint64_t key, uint32_t pid) {
    const unsigned int slot = atomicAdd(s_count, 1u);
    if (slot < (unsigned)kPfMcap) {
      s_keys[slot] = key;
      s_pids[slot] = pid;
    } else {  // staging full: append directly
      const unsigned long long g = atomicAdd(g_counter, 1ull);
      if (g < cap) { g_keys[g] = key; g_pids[g] = pid; }
    }
  }
};

// Verify one candidate start offset `s` (K3b).
template <int MODE>
__device__ __forceinline__ void verify_at(const DfaDev& d, const PrefilterLaunch& p, const uint8_t* s_cls,
                                          uint64_t s, Emitter& em) {
  const uint8_t* __restrict__ hay = p.hay;
  const uint32_t* __restrict__ trans = d.trans;
  uint32_t sid = d.start_unanchored_id;
  uint64_t pos = s;
  uint32_t j = 0;
  uint32_t best_pid = 0, best_len = 0;
  while (pos < p.span_end) {
    const uint32_t b = __ldg(hay + pos);
    sid = __ldg(trans + sid + s_cls[b]);
    ++j;
    ++pos;
    if (sid == 0) break;  // DEAD (leftmost automata): nothing longer can start at s
    const uint32_t row = sid >> d.stride2;
    if (__ldg(d.depth16 + row) != j) break;  // left the trie path anchored at s
    if (sid <= d.max_match_id) {
      const uint32_t lo = __ldg(d.match_offsets + row - 2), hi = __ldg(d.match_offsets + row - 1);
      if (MODE == 0) {
        // the node's own patterns come first in its list and all have length j
        for (uint32_t i = lo; i < hi; ++i) {
          const uint32_t pid = __ldg(d.match_pids + i);
          if (__ldg(d.pattern_lens + pid) != j) break;
          const uint64_t tie = ((uint64_t)(d.max_pattern_len - j) << p.dup_shift) | (uint64_t)(i - lo);
          em.emit(((pos - p.span_start) << kTieBits) | tie, pid);
        }
      } else {
        const uint32_t pid = __ldg(d.match_pids + lo);
        if (__ldg(d.pattern_lens + pid) == j) { best_pid = pid; best_len = j; }
      }
    }
  }
  if (MODE == 1 && best_len) em.emit(((s - p.span_start) << kTieBits) | best_len, best_pid);
}

template <int MODE>
__global__ void __launch_bounds__(kPfThreads, 2)
prefilter_kernel(DfaDev d, PrefilterLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_mkeys = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* s_mpids = reinterpret_cast<uint32_t*>(s_mkeys + kPfMcap);
  uint32_t* s_queue = s_mpids + kPfMcap;
  uint32_t* s_bitmap = s_queue + kPfQcap;
  __shared__ PfShared sh;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const uint32_t bitmap_words = p.brute ? 0u : (1u << (p.log_bits - 5));
  for (uint32_t i = tid; i < bitmap_words; i += kPfThreads) s_bitmap[i] = p.bitmap[i];
  if (tid < 256) sh.cls[tid] = d.classes[tid];
  if (tid == 0) {
    sh.q_count = 0;
    sh.m_count = 0;
    sh.overflow = 0;
    sh.q_base = p.region_lo;
    sh.cand_total = 0;
  }
  __syncthreads();

  Emitter em{p.keys, p.pids, p.counter, p.cap, s_mkeys, s_mpids, &sh.m_count};

  auto flush_matches = [&]() {
    // called by all threads, after a __syncthreads
    const unsigned int n = min(sh.m_count, (unsigned)kPfMcap);
    if (n) {
      if (tid == 0) sh.m_base = atomicAdd(p.counter, (unsigned long long)n);
      __syncthreads();
      const unsigned long long base = sh.m_base;
      for (unsigned int i = tid; i < n; i += kPfThreads)
        if (base + i < p.cap) { p.keys[base + i] = s_mkeys[i]; p.pids[base + i] = s_mpids[i]; }
    }
    __syncthreads();
    if (tid == 0) sh.m_count = 0;
    __syncthreads();
  };

  auto drain_queue = [&]() {
    // all threads; queue is stable
    const unsigned int qn = sh.q_count;
    const unsigned long long qb = sh.q_base;
    for (unsigned int i = tid; i < qn; i += kPfThreads) verify_at<MODE>(d, p, sh.cls, qb + s_queue[i], em);
    __syncthreads();
    if (tid == 0) { sh.cand_total += qn; sh.q_count = 0; }
    __syncthreads();
    flush_matches();
  };

  // head / tail positions outside the aligned filter region are unconditional candidates
  if (blockIdx.x == 0) {
    const uint64_t head_n = p.region_lo - p.span_start;
    const uint64_t tail_n = p.span_end >= p.region_hi ? p.span_end - p.region_hi + 1 : 0;  // incl. s == span_end (no-op)
    for (uint64_t i = tid; i < head_n + tail_n; i += kPfThreads) {
      const uint64_t s = i < head_n ? p.span_start + i : p.region_hi + (i - head_n);
      if (s < p.span_end) verify_at<MODE>(d, p, sh.cls, s, em);
    }
    __syncthreads();
    flush_matches();
  }

  const uint32_t kmask = p.kmask, fold = p.fold, mult = p.mult, shift = p.shift;

  for (uint64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const uint64_t t0 = p.region_lo + tile * p.tile_bytes;
    const uint64_t t1 = min(t0 + p.tile_bytes, p.region_hi);
    if (tid == 0) {
      sh.q_tile_mark = sh.q_count;
      if (sh.q_count == 0) sh.q_base = t0;
    }
    __syncthreads();
    const uint64_t qb = sh.q_base;

    if (!p.brute) {
      // ---- K3: fingerprint test, 16 positions per lane per step ----
      for (uint64_t blk = t0 + (uint64_t)tid * 16; blk < t1 + (uint64_t)lane * 16; blk += (uint64_t)kPfThreads * 16) {
        // the loop bound keeps whole warps together (lane 0's block decides), so the shuffle is safe
        const bool active = blk < t1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (active) v = ld_stream_u4(p.hay + blk);
        uint32_t nx = __shfl_down_sync(0xffffffffu, v.x, 1);
        if (active && (lane == 31 || blk + 16 >= t1)) nx = __ldg(reinterpret_cast<const uint32_t*>(p.hay + blk + 16));
        const uint32_t w[5] = {v.x, v.y, v.z, v.w, nx};
        uint32_t mask = 0;
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const uint32_t win = (o & 3) ? __funnelshift_r(w[o >> 2], w[(o >> 2) + 1], (o & 3) * 8) : w[o >> 2];
          const uint32_t idx = (((win | fold) & kmask) * mult) >> shift;
          const uint32_t word = s_bitmap[idx >> 5];
          mask |= ((word >> (idx & 31)) & 1u) << o;
        }
        if (active && mask) {
          const unsigned int cnt = __popc(mask);
          const unsigned int slot = atomicAdd(&sh.q_count, cnt);
          if (slot + cnt <= (unsigned)kPfQcap) {
            unsigned int k = slot;
            const uint32_t rel = (uint32_t)(blk - qb);
            while (mask) {
              const int o = __ffs(mask) - 1;
              mask &= mask - 1;
              s_queue[k++] = rel + o;
            }
          } else {
            sh.overflow = 1;
          }
        }
      }
      __syncthreads();
    }

    if (p.brute || sh.overflow) {
      // fingerprints not selective here: drop this tile's queue entries and verify every position
      __syncthreads();
      if (tid == 0) { sh.q_count = sh.q_tile_mark; sh.overflow = 0; }
      __syncthreads();
      for (uint64_t s = t0 + tid; s < t1; s += kPfThreads) verify_at<MODE>(d, p, sh.cls, s, em);
      __syncthreads();
      if (tid == 0) sh.cand_total += (t1 - t0);
      __syncthreads();
      flush_matches();
    }

    const bool last = tile + gridDim.x >= p.n_tiles;
    if (sh.q_count > (unsigned)(kPfQcap / 2) || last || (t1 - qb) > 0xF0000000ull) drain_queue();
  }
  if (tid == 0 && sh.cand_total) atomicAdd(p.counter + 1, sh.cand_total);
}

// ---- chain resolution ---------------------------------------------------------

__device__ __forceinline__ void tuple_span(const ChainLaunch& c, uint64_t i, uint64_t* s, uint64_t* e) {
  const uint64_t key = c.keys[i];
  if (c.mode == 1) {
    *s = key >> kTieBits;
    *e = *s + (key & kTieMask);
  } else {
    *e = key >> kTieBits;
    *s = *e - c.pattern_lens[c.pids[i]];
  }
}

__global__ void chain_ends_kernel(ChainLaunch c) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  uint64_t s, e;
  tuple_span(c, i, &s, &e);
  c.scratch_end[i] = e;
}

// Entry j is an "anchor" when every earlier tuple ends at or before its start:
// the iterator's cursor is then <= start(j) whatever it did before, so j is
// yielded and the search restarts at end(j) (src/automaton.rs:927-935).  Each
// anchor's thread resolves the short run of mutually overlapping tuples after it.
__global__ void chain_select_kernel(ChainLaunch c) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  uint64_t s, e;
  tuple_span(c, i, &s, &e);
  if (i > 0 && c.scratch_end[i - 1] > s) return;  // not an anchor: an earlier anchor's thread decides
  c.flags[i] = 1;
  uint64_t cur_end = e;
  for (uint64_t j = i + 1; j < c.n; ++j) {
    uint64_t sj, ej;
    tuple_span(c, j, &sj, &ej);
    if (c.scratch_end[j - 1] <= sj) break;  // next anchor
    if (sj >= cur_end) { c.flags[j] = 1; cur_end = ej; } else c.flags[j] = 0;
  }
}

struct MaxOp {
  __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; }
};

}  // namespace

cudaError_t launch_prefilter(const DfaDev& dfa, const PrefilterLaunch& p, int sm_count, cudaStream_t s) {
  const size_t bitmap_bytes = p.brute ? 0 : (size_t(1) << (p.log_bits - 3));
  const size_t smem = size_t(kPfMcap) * 12 + size_t(kPfQcap) * 4 + bitmap_bytes;
  auto kern = p.mode == 0 ? prefilter_kernel<0> : prefilter_kernel<1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPfThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = (uint64_t)sm_count * per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles ? p.n_tiles : 1;
  kern<<<(unsigned)grid, kPfThreads, smem, s>>>(dfa, p);
  return cudaGetLastError();
}

cudaError_t launch_chain_ends(const ChainLaunch& c, cudaStream_t s) {
  const unsigned blocks = (unsigned)((c.n + 255) / 256);
  chain_ends_kernel<<<blocks, 256, 0, s>>>(c);
  return cudaGetLastError();
}
cudaError_t launch_chain_select(const ChainLaunch& c, cudaStream_t s) {
  const unsigned blocks = (unsigned)((c.n + 255) / 256);
  chain_select_kernel<<<blocks, 256, 0, s>>>(c);
  return cudaGetLastError();
}
cudaError_t scan_max_u64(void* d_temp, size_t& temp_bytes, uint64_t* data, uint64_t n, cudaStream_t s) {
  return cub::DeviceScan::InclusiveScan(d_temp, temp_bytes, data, data, MaxOp(), (int64_t)n, s);
}
cudaError_t select_flagged(void* d_temp, size_t& temp_bytes, const uint64_t* keys_in, const uint32_t* pids_in,
                           const uint8_t* flags, uint64_t* keys_out, uint32_t* pids_out,
                           unsigned long long* d_num_out, uint64_t n, cudaStream_t s) {
  size_t a = 0, b = 0;
  if (d_temp == nullptr) {
    cudaError_t e = cub::DeviceSelect::Flagged(nullptr, a, keys_in, flags, keys_out, d_num_out, (int64_t)n, s);
    if (e != cudaSuccess) return e;
    e = cub::DeviceSelect::Flagged(nullptr, b, pids_in, flags, pids_out, d_num_out, (int64_t)n, s);
    temp_bytes = a > b ? a : b;
    return e;
  }
  cudaError_t e = cub::DeviceSelect::Flagged(d_temp, temp_bytes, keys_in, flags, keys_out, d_num_out, (int64_t)n, s);
  if (e != cudaSuccess) return e;
  return cub::DeviceSelect::Flagged(d_temp, temp_bytes, pids_in, flags, pids_out, d_num_out, (int64_t)n, s);
}

}  // namespace acb
