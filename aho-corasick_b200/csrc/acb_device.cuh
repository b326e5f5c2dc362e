// acb_device.cuh -- device-side views shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acb {

// Packed 64-bit ordering key of one emitted match: (end - span_start) in the
// high 40 bits, a 24-bit tie-break in the low bits.  Sorting the keys ascending
// reproduces the reference's emission order:
//   * walk engine:      tie-break = index into the match state's pattern list
//                        (src/automaton.rs:1475-1481 reports list entries in order)
//   * prefilter engine: tie-break = (max_len - len) * dup_cap + index among the
//                        node's own (equal-length) patterns, which orders equal
//                        ends by (length desc, list order) -- the order in which
//                        src/nfa/noncontiguous.rs:490-523 concatenates lists.
constexpr int kTieBits = 24;
constexpr uint64_t kTieMask = (1ull << kTieBits) - 1;
constexpr uint64_t kInvalidKey = ~0ull;

struct DfaDev {
  const uint32_t* trans;          // premultiplied ids, as shipped (src/dfa.rs:92-95)
  const uint8_t* classes;         // [256] byte -> class (src/util/alphabet.rs)
  const uint32_t* match_offsets;  // CSR over match-state rows 2..
  const uint32_t* match_pids;
  const uint32_t* pattern_lens;
  const uint8_t* depth8;          // trie depth per row (min(depth,255)); prefilter engine
  uint32_t stride2;
  uint32_t max_match_id;
  uint32_t start_unanchored_id;
  uint32_t start_anchored_id;
  uint32_t max_pattern_len;
  uint32_t min_pattern_len;
};

// ---- launch wrappers (acb_kernels.cu) --------------------------------------

struct WalkLaunch {
  const uint8_t* hay;   // device pointer to haystack byte 0
  uint64_t span_start, span_end;
  uint64_t seg_len;     // bytes owned per lane
  uint64_t n_segs;
  uint64_t* keys;       // [cap]
  uint32_t* pids;       // [cap]
  unsigned long long* counter;  // total tuples wanted (may exceed cap => overflow)
  uint64_t cap;
};
cudaError_t launch_walk_overlapping(const DfaDev& dfa, const WalkLaunch& p, cudaStream_t s);

// Single-lane restatement of FindIter (src/automaton.rs:857-936) over
// try_find_fwd (:1259-1420): anchored inputs, automata containing the empty
// pattern, and tiny spans.  Writes (pid,start,end) triples as 3 x u64.
struct SeqLaunch {
  const uint8_t* hay;
  uint64_t span_start, span_end;
  int anchored;
  int match_kind;
  int earliest;         // for single find
  int single;           // 1: stop after the first match (AhoCorasick::try_find)
  uint64_t* out;        // [cap * 3]
  unsigned long long* counter;
  uint64_t cap;
};
cudaError_t launch_seq_find(const DfaDev& dfa, const SeqLaunch& p, cudaStream_t s);

// key/pid pair sort (K4). temp storage is queried with d_temp == nullptr.
cudaError_t sort_pairs(void* d_temp, size_t& temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                       const uint32_t* vals_in, uint32_t* vals_out, uint64_t n, int end_bit,
                       cudaStream_t s);

}  // namespace acb
