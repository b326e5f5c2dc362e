// acb_kernels.cu -- sm_100a kernels of the Aho-Corasick search path.
//
//   K1  walk_overlapping_kernel   sharded DFA state-transition scan; replaces the
//                                 loop of try_find_overlapping_fwd_imp
//                                 (src/automaton.rs:1491-1534) + DFA::next_state
//                                 (src/dfa.rs:218-226) + match expansion (:275-286)
//   Kseq seq_find_kernel          single-lane FindIter/try_find (anchored inputs,
//                                 empty-pattern automata)
//   K4  sort_pairs                ordering of the appended tuples (CUB radix sort)
#include "acb_device.cuh"
#ifndef ACB_PTX_HEADER
#define ACB_PTX_HEADER "acb_ptx.cuh"
#endif
#include ACB_PTX_HEADER

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

namespace acb {

namespace {

// Append the whole pattern list of match state `sid` (entered after consuming the byte at
// absolute offset end-1) to the tuple buffer.  Lanes of a warp that report in the same step share
// one atomic (warp-ballot + warp-aggregated append), list entry by list entry.  Out of line: matches
// are rare (one per 4 KiB in the BASELINE workloads) and the walk loop must stay small.
__device__ __noinline__ void emit_state_matches(const uint32_t* match_offsets, const uint32_t* match_pids,
                                                uint32_t row, uint64_t end_rel, uint64_t* keys, uint32_t* pids,
                                                unsigned long long* counter, uint64_t cap) {
  const uint32_t lo = match_offsets[row], hi = match_offsets[row + 1];
  const int lane = threadIdx.x & 31;
  for (uint32_t i = lo; i < hi; ++i) {
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    const unsigned long long g = base + __popc(m & ((1u << lane) - 1));
    if (g < cap) {  // overflow: the host sees counter > cap and retries
      keys[g] = (end_rel << kTieBits) | (uint64_t)(i - lo);
      pids[g] = match_pids[i];
    }
  }
}

constexpr int kWalkThreads = 256;
constexpr int kWalkChains = 4;  // independent shards walked by one lane (must match acb_api.cu)

// K1: the state-transition loop of src/automaton.rs:1491-1534 over DFA::next_state
// (src/dfa.rs:218-226), sharded.  A shard starts cold (start state) a little more than
// max_pattern_len-1 bytes before its first owned byte -- the Aho-Corasick state depends on at most
// that many trailing bytes -- and only reports matches whose end lies inside the shard, so every
// end offset is owned by exactly one shard.
// The loop is one dependent table load per byte (an L2 hit for everything below the first trie
// levels: ~250 cycles), so a lane that walks a single shard keeps one load in flight and the SM
// idles (r01 ncu: 34 long-scoreboard stall cycles per issued instruction, issue slots 23 % busy).
// Here every lane walks kWalkChains consecutive shards at once, byte k of each in turn: four
// independent chains per lane, 16-byte vector loads of each shard, byte classes in shared memory.
// (A variant that staged the start / depth-1 rows in shared memory behind a flagged table copy was
// measured in r02 and lost: 4.68 ms against 2.75 ms per GiB on cfg 2, profiles/r02a_ab_walk.jsonl.)
__global__ void __launch_bounds__(kWalkThreads, 4)
walk_overlapping_kernel(DfaDev d, WalkLaunch p) {
  __shared__ uint8_t s_cls[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cls[i] = d.classes[i];
  __syncthreads();

  const uint32_t* __restrict__ trans = d.trans;
  const uint8_t* __restrict__ hay = p.hay;
  const uint32_t max_match = d.max_match_id;
  const int64_t span_start = (int64_t)p.span_start, span_end = (int64_t)p.span_end;
  // shard grid: origin at the 16-byte boundary at or before the span start, so that every block of
  // every shard is a 16-byte aligned address; shard k = [origin + k * seg_len, + seg_len) cut to the span
  const int64_t origin = span_start - (int64_t)(reinterpret_cast<uintptr_t>(hay + p.span_start) & 15);
  const int64_t seg_len = (int64_t)p.seg_len;
  const int64_t back16 = ((d.max_pattern_len > 0 ? (int64_t)d.max_pattern_len - 1 : 0) + 15) & ~(int64_t)15;
  const uint64_t first_seg = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kWalkChains;
  if (first_seg >= p.n_segs) return;
  auto report = [&](uint32_t sid, int64_t pos) {
    emit_state_matches(d.match_offsets, d.match_pids, (sid >> d.stride2) - 2, (uint64_t)(pos + 1 - span_start),
                       p.keys, p.pids, p.counter, p.cap);
  };
  // matches of the start state itself (empty patterns) at the very beginning of the span are
  // reported before the first byte (src/automaton.rs:1456-1464)
  if (first_seg == 0 && d.start_unanchored_id != 0 && d.start_unanchored_id <= max_match)
    report(d.start_unanchored_id, span_start - 1);

  uint32_t sid[kWalkChains];
#pragma unroll
  for (int c = 0; c < kWalkChains; ++c) sid[c] = d.start_unanchored_id;
  const int64_t g00 = origin + (int64_t)first_seg * seg_len;  // first owned byte of chain 0; chain c: + c * seg_len
  const int64_t n_blocks = (back16 + seg_len) >> 4;
  for (int64_t j = 0; j < n_blocks; ++j) {
    const int64_t blk0 = g00 - back16 + (j << 4);  // this block of chain 0
    const bool warm = (j << 4) < back16;           // cold-start run-in: nothing is reported
    // every chain has a whole block inside the span (then also inside its shard): 16 x 4 transitions
    // with no per-byte checks, tracking only the smallest state id each chain saw -- match states are
    // the smallest ids (src/dfa.rs:229-247), so one comparison per chain and block tells whether
    // anything is to be reported; the rare chain that has something walks its 16 bytes again, carefully
    if (blk0 >= span_start && blk0 + (kWalkChains - 1) * seg_len + 16 <= span_end) {
      uint32_t at[kWalkChains], lowest[kWalkChains];
#pragma unroll
      for (int c = 0; c < kWalkChains; ++c) { at[c] = sid[c]; lowest[c] = 0xFFFFFFFFu; }
      const uint8_t* src = hay + blk0;
#pragma unroll 1
      for (int wi = 0; wi < 4; ++wi, src += 4) {  // rolled: keeps the live set to one word per chain
        uint32_t w[kWalkChains];
#pragma unroll
        for (int c = 0; c < kWalkChains; ++c) w[c] = __ldg(reinterpret_cast<const uint32_t*>(src + c * seg_len));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int c = 0; c < kWalkChains; ++c) {
            const uint32_t b = __byte_perm(w[c], 0, 0x4440 + k);
            sid[c] = __ldg(trans + sid[c] + s_cls[b]);
            lowest[c] = min(lowest[c], sid[c]);
          }
        }
      }
      if (!warm) {
#pragma unroll
        for (int c = 0; c < kWalkChains; ++c) {
          if (lowest[c] > max_match) continue;
          uint32_t s = at[c];
          const int64_t pos0 = blk0 + c * seg_len;
#pragma unroll 1
          for (int k = 0; k < 16; ++k) {
            s = __ldg(trans + s + s_cls[hay[pos0 + k]]);
            if (s <= max_match && s != 0) report(s, pos0 + k);
          }
        }
      }
    } else {
      // a block that crosses the span's edges, or shards that do not exist: byte by byte, nothing
      // outside [cold start, shard end) is read
#pragma unroll 1
      for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int c = 0; c < kWalkChains; ++c) {
          const int64_t g0 = g00 + c * seg_len;
          const int64_t pos = blk0 + c * seg_len + k;
          if (g0 >= span_end || pos < span_start || pos >= min(g0 + seg_len, span_end)) continue;
          const uint32_t s = __ldg(trans + sid[c] + s_cls[hay[pos]]);
          sid[c] = s;
          if (s <= max_match && s != 0 && pos >= g0) report(s, pos);
        }
      }
    }
  }
}

// ---- dense-table construction (one BFS level per launch) -------------------------
constexpr int kFillThreads = 256;

__global__ void __launch_bounds__(kFillThreads) dfa_fill_level_kernel(FillLaunch f) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t w = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (w >= f.n) return;
  uint32_t* dst = f.trans + ((size_t)f.row[w] << f.stride2);
  const uint32_t inherit = f.inherit_row[w];
  if (inherit != UINT32_MAX) {
    const uint32_t* src = f.trans + ((size_t)inherit << f.stride2);  // written by an earlier launch
    for (uint32_t c = lane; c < f.alphabet_len; c += 32) dst[c] = src[c];
  } else {
    const uint32_t v = f.fill_id[w];
    for (uint32_t c = lane; c < f.alphabet_len; c += 32) dst[c] = v;
  }
  __syncwarp();
  for (uint32_t e = f.edge_off[w] + lane; e < f.edge_off[w + 1]; e += 32) dst[f.edge_class[e]] = f.edge_to[e];
}

// ---- sequential engine -------------------------------------------------------

struct SeqMatch {
  uint32_t pid;
  uint64_t start, end;
};

__device__ __forceinline__ bool seq_is_match(const DfaDev& d, uint32_t sid) {
  return sid != 0 && sid <= d.max_match_id;
}
__device__ __forceinline__ SeqMatch seq_get_match(const DfaDev& d, uint32_t sid, uint32_t idx, uint64_t at) {
  const uint32_t row = (sid >> d.stride2) - 2;
  const uint32_t pid = d.match_pids[d.match_offsets[row] + idx];
  SeqMatch m;
  m.pid = pid;
  m.start = at - d.pattern_lens[pid];
  m.end = at;
  return m;
}

// try_find_fwd_imp, src/automaton.rs:1285-1420 (prefilter-free instance)
__device__ bool seq_try_find(const DfaDev& d, const uint8_t* hay, uint64_t start, uint64_t end,
                             bool anchored, bool earliest, SeqMatch* out) {
  if (start > end) return false;
  uint32_t sid = anchored ? d.start_anchored_id : d.start_unanchored_id;
  uint64_t at = start;
  bool have = false;
  SeqMatch mat;
  if (seq_is_match(d, sid)) {
    mat = seq_get_match(d, sid, 0, at);
    have = true;
    if (earliest) { *out = mat; return true; }
  }
  while (at < end) {
    sid = d.trans[sid + d.classes[hay[at]]];
    if (sid <= d.max_match_id) {
      if (sid == 0) break;
      SeqMatch m = seq_get_match(d, sid, 0, at + 1);
      if (!(anchored && m.start > start)) {
        mat = m;
        have = true;
        if (earliest) break;
      }
    }
    ++at;
  }
  if (have) *out = mat;
  return have;
}

__global__ void seq_find_kernel(DfaDev d, SeqLaunch p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool anchored = p.anchored != 0;
  const bool earliest = p.match_kind == 0 || p.earliest != 0;
  uint64_t start = p.span_start;
  unsigned long long n = 0;
  bool have_last = false;
  uint64_t last_end = 0;
  for (;;) {
    SeqMatch m;
    if (!seq_try_find(d, p.hay, start, p.span_end, anchored, earliest, &m)) break;
    if (!p.single && m.start == m.end && have_last && m.end == last_end) {
      // FindIter::handle_overlapping_empty_match, src/automaton.rs:910-920
      start += 1;
      if (!seq_try_find(d, p.hay, start, p.span_end, anchored, earliest, &m)) break;
    }
    if (n < p.cap) {
      p.out[n * 3 + 0] = m.pid;
      p.out[n * 3 + 1] = m.start;
      p.out[n * 3 + 2] = m.end;
    }
    ++n;
    if (p.single) break;
    start = m.end;
    last_end = m.end;
    have_last = true;
  }
  *p.counter = n;
}

// 256 records per CTA, built in shared memory and stored as 16-byte vectors: the target may be
// another GPU's HBM (peer mapping of rank 0's receive buffer, acb_comm.hpp), where full 128-byte
// lines per warp store matter more than at home.  Record = acg_match { u32 pid; u32 pad; u64 start; u64 end }.
__global__ void __launch_bounds__(256) expand_kernel(ExpandLaunch e) {
  __shared__ uint64_t s_rec[256 * 3];
  const uint64_t m = e.n - e.first;
  const uint64_t base = (uint64_t)blockIdx.x * 256;
  const uint64_t i = base + threadIdx.x;
  if (i < m) {
    const uint64_t key = e.keys[e.first + i];
    const uint32_t pid = e.pids[e.first + i];
    const uint64_t end = e.span_start + (key >> kTieBits) + e.offset_add;
    s_rec[threadIdx.x * 3 + 0] = (uint64_t)pid;
    s_rec[threadIdx.x * 3 + 1] = end - e.pattern_lens[pid];
    s_rec[threadIdx.x * 3 + 2] = end;
  }
  __syncthreads();
  const uint64_t cnt = m - base < 256 ? m - base : 256;  // records of this CTA
  uint64_t* dst = e.out + base * 3;
  const uint32_t words = (uint32_t)cnt * 3;  // 8-byte words; base * 24 is a multiple of 16
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (uint32_t w = threadIdx.x * 2; w + 1 < words; w += 512)
      *reinterpret_cast<ulonglong2*>(dst + w) = make_ulonglong2(s_rec[w], s_rec[w + 1]);
    if ((words & 1) && threadIdx.x == 0) dst[words - 1] = s_rec[words - 1];
  } else {
    for (uint32_t w = threadIdx.x; w < words; w += 256) dst[w] = s_rec[w];
  }
}

// keys are sorted: count the tuples with key < bound_key (single block, strided binary chunks)
__global__ void lower_bound_kernel(const uint64_t* keys, uint64_t n, uint64_t bound_key,
                                   unsigned long long* result) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (keys[mid] < bound_key) lo = mid + 1; else hi = mid;
  }
  *result = lo;
}

}  // namespace

cudaError_t launch_expand(const ExpandLaunch& e, cudaStream_t s) {
  const uint64_t m = e.n - e.first;
  if (m == 0) return cudaSuccess;
  ACB_LAUNCH(expand_kernel, (unsigned)((m + 255) / 256), 256, 0, s, e);
  return cudaGetLastError();
}
cudaError_t launch_lower_bound(const uint64_t* keys, uint64_t n, uint64_t bound_key,
                               unsigned long long* d_result, cudaStream_t s) {
  ACB_LAUNCH(lower_bound_kernel, 1, 32, 0, s, keys, n, bound_key, d_result);
  return cudaGetLastError();
}

cudaError_t launch_walk_overlapping(const DfaDev& dfa, const WalkLaunch& p, cudaStream_t s) {
  const uint64_t lanes = (p.n_segs + kWalkChains - 1) / kWalkChains;
  const uint64_t blocks = (lanes + kWalkThreads - 1) / kWalkThreads;
  ACB_LAUNCH(walk_overlapping_kernel, (unsigned)blocks, kWalkThreads, 0, s, dfa, p);
  return cudaGetLastError();
}

cudaError_t launch_dfa_fill_level(const FillLaunch& f, cudaStream_t s) {
  if (f.n == 0) return cudaSuccess;
#ifdef ACB_EMULATE
  if (getenv("ACB_EMU_TRACE")) fprintf(stderr, "launch_dfa_fill_level rows %u\n", f.n);
#endif
  const uint64_t blocks = ((uint64_t)f.n * 32 + kFillThreads - 1) / kFillThreads;
  ACB_LAUNCH(dfa_fill_level_kernel, (unsigned)blocks, kFillThreads, 0, s, f);
  return cudaGetLastError();
}


cudaError_t launch_seq_find(const DfaDev& dfa, const SeqLaunch& p, cudaStream_t s) {
  ACB_LAUNCH(seq_find_kernel, 1, 32, 0, s, dfa, p);
  return cudaGetLastError();
}

cudaError_t sort_pairs(void* d_temp, size_t& temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                       const uint32_t* vals_in, uint32_t* vals_out, uint64_t n, int end_bit,
                       cudaStream_t s) {
  return cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n,
                                         0, end_bit, s);
}

}  // namespace acb
