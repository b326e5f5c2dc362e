// acb_build.cpp -- host-side automaton construction (see acb_build.hpp).
//
// Behaviour follows the reference (BurntSushi/aho-corasick 1.1.3) so that the
// tables are bit-identical; citations give the file:line whose *result* each
// step reproduces.  The data structures and the DFA fill are this project's own.
#include "acb_build.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>

#include "../../include/acb200.h"

namespace acb {
namespace {

constexpr uint32_t kDead = 0, kFailId = 1, kRoot = 2, kAnchoredRoot = 3;
constexpr uint64_t kMaxIndex = 0x7FFFFFFEull;  // SmallIndex::MAX, src/util/primitives.rs:96-111

inline uint8_t flip_ascii_case(uint8_t b) {  // src/util/prefilter.rs:906-914
  if (b >= 'A' && b <= 'Z') return uint8_t(b + 32);
  if (b >= 'a' && b <= 'z') return uint8_t(b - 32);
  return b;
}

struct Edge {
  uint8_t byte;
  uint32_t to;
};

// Explicit trie. Node ids are allocated exactly like the reference allocates
// NFA states (src/nfa/noncontiguous.rs:977-985, 1132-1143): 0 DEAD, 1 FAIL,
// 2 unanchored start, 3 anchored start, then one node per new trie edge in
// pattern order.
struct Trie {
  std::vector<std::vector<Edge>> edges;     // sorted by byte
  std::vector<std::vector<uint32_t>> hits;  // match list, in the reference's emission order
  std::vector<uint32_t> fail;
  bool root_loop_closed = false;  // src/nfa/noncontiguous.rs:1620-1638

  uint32_t add_node() {
    edges.emplace_back();
    hits.emplace_back();
    fail.push_back(kRoot);
    return uint32_t(edges.size() - 1);
  }
  int find(uint32_t s, uint8_t b) const {
    const auto& e = edges[s];
    auto it = std::lower_bound(e.begin(), e.end(), b, [](const Edge& x, uint8_t v) { return x.byte < v; });
    return (it != e.end() && it->byte == b) ? int(it - e.begin()) : -1;
  }
  void link(uint32_t s, uint8_t b, uint32_t to) {
    auto& e = edges[s];
    auto it = std::lower_bound(e.begin(), e.end(), b, [](const Edge& x, uint8_t v) { return x.byte < v; });
    if (it != e.end() && it->byte == b) it->to = to; else e.insert(it, Edge{b, to});
  }
  // goto function used while computing failure links: the root loops on
  // undefined bytes (src/nfa/noncontiguous.rs:1597-1606), DEAD is absorbing
  // (:1643-1646), everything else reports "undefined" as kFailId.
  uint32_t step(uint32_t s, uint8_t b) const {
    if (s == kDead) return kDead;
    int i = find(s, b);
    if (i >= 0) return edges[s][size_t(i)].to;
    return s == kRoot ? kRoot : kFailId;
  }
};

// ---- prefilter selection (decision only) -----------------------------------
// Data table: heuristic byte-frequency ranks of src/util/byte_frequencies.rs,
// hex encoded, index = byte value.
const char kRankHex[] =
    "3734333231302f2e2d67f24243e52c2b2a29282726252423222138201f1e1d1cff94a49588a09badddde867ae8cad7e0"
    "d0dcccbbb7b3b1a8b2c8e2c39ab8ae7e78bf9dc2aabda2a196c18e89abb0b9a7ba70afc0bc9c8c8f7b8580938a9272df"
    "97f9d8eeecfde3dae6f787b4f1e9f6f4e78bf5f3fbebc9c4f0d698b6cdb57f1bd4d3d2d5e4c5a99f83ac695062606151"
    "cf917473908299796b846d6e7c6f526c768d7181777da5755c6a5348635d414fa6eda3c7bee1d1cbc6d9dbceeaf89eef"
    "ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff"
    "ffffffffffffffffffffffffffffffff";
uint8_t rank_of(uint8_t b) {
  auto hv = [](char c) { return c <= '9' ? c - '0' : c - 'a' + 10; };
  return uint8_t(hv(kRankHex[2 * b]) * 16 + hv(kRankHex[2 * b + 1]));
}

// Mirrors the *outcome* of prefilter::Builder (src/util/prefilter.rs:91-326):
// which prefilter kind the reference ends up with for a pattern set.
class PrefilterChooser {
 public:
  PrefilterChooser(int match_kind, bool ci) : ci_(ci), wants_packed_(match_kind != kStandard) {}

  void add(const uint8_t* p, uint64_t n) {  // Builder::add :308-323
    if (n == 0) enabled_ = false;
    if (!enabled_) return;
    ++count_;
    // StartBytesBuilder::add :826-852
    if (start_count_ <= 3) {
      mark_start(p[0]);
      if (ci_) mark_start(flip_ascii_case(p[0]));
    }
    add_rare(p, n);
    // packed::Builder::add, src/packed/api.rs:294-322
    if (wants_packed_ && !packed_inert_) {
      if (packed_lens_.size() >= 128) { packed_inert_ = true; packed_lens_.clear(); }
      else packed_lens_.push_back(n);
    }
  }

  // The bytes (ascending) and offsets of the prefilter `kind` that choose() returned.
  void byte_set(int kind, uint32_t* n, uint8_t bytes[3], uint8_t back[3]) const {
    *n = 0;
    const bool* set = kind == kPreStartBytes ? start_set_ : (kind == kPreRareBytes ? rare_set_ : nullptr);
    if (!set) return;
    for (int b = 0; b < 256 && *n < 3; ++b)
      if (set[b]) { bytes[*n] = uint8_t(b); back[*n] = kind == kPreRareBytes ? max_off_[b] : 0; ++*n; }
  }

  int choose(PackedPlan* plan) const {  // Builder::build :163-305
    *plan = PackedPlan{};
    if (!enabled_) return kPreNone;
    if (!ci_ && count_ == 1) return kPreMemmem;
    uint64_t patlen = UINT64_MAX, minlen = 0;
    PackedPlan pk;
    if (!ci_ && wants_packed_) {
      patlen = packed_lens_.size();
      minlen = UINT64_MAX;
      for (uint64_t l : packed_lens_) minlen = std::min(minlen, l);
      pk = teddy_plan(patlen, minlen);
    }
    const bool has_start = start_available(), has_rare = rare_available();
    auto packed_or_none = [&]() { if (pk.active) { *plan = pk; return int(kPrePacked); } return int(kPreNone); };
    if (has_start && has_rare) {
      if (patlen <= 16 && minlen >= 2 && start_count_ >= 3 && rare_count_ >= 3) return packed_or_none();
      const bool fewer = start_count_ < rare_count_;
      const bool rarer = start_rank_sum_ <= uint32_t(rare_rank_sum_) + 50;
      return (fewer || rarer) ? kPreStartBytes : kPreRareBytes;
    }
    if (has_start) {
      if (patlen <= 16 && minlen >= 2 && start_count_ >= 3) return packed_or_none();
      return kPreStartBytes;
    }
    if (has_rare) {
      if (patlen <= 16 && minlen >= 2 && rare_count_ >= 3) return packed_or_none();
      return kPreRareBytes;
    }
    if (ci_) return kPreNone;
    return packed_or_none();
  }

 private:
  // teddy::Builder::build_imp on x86_64 with AVX2, src/packed/teddy/builder.rs:98-231
  static PackedPlan teddy_plan(uint64_t npat, uint64_t minlen) {
    PackedPlan p;
    if (npat == 0 || npat > 64) return p;
    const int mask_len = int(std::min<uint64_t>(4, minlen));
    if (mask_len == 1 && npat > 16) return p;
    p.active = true;
    p.fat = npat > 32;
    p.mask_len = mask_len;
    return p;
  }
  void mark_start(uint8_t b) {
    if (!start_set_[b]) { start_set_[b] = true; ++start_count_; start_rank_sum_ = uint16_t(start_rank_sum_ + rank_of(b)); }
  }
  bool start_available() const {  // StartBytesBuilder::build :784-824
    if (start_count_ > 3) return false;
    unsigned len = 0;
    for (int b = 0; b < 256; ++b) if (start_set_[b]) { if (b > 0x7F) return false; ++len; }
    return len != 0;
  }
  void mark_rare(uint8_t b) {
    if (!rare_set_[b]) { rare_set_[b] = true; ++rare_count_; rare_rank_sum_ = uint16_t(rare_rank_sum_ + rank_of(b)); }
  }
  void add_rare(const uint8_t* p, uint64_t n) {  // RareBytesBuilder::add :585-630
    if (!rare_ok_) return;
    if (rare_count_ > 3 || n >= 256) { rare_ok_ = false; return; }
    uint8_t best = p[0], best_rank = rank_of(p[0]);
    bool found = false;
    for (uint64_t i = 0; i < n; ++i) {
      set_offset(i, p[i]);  // RareByteOffsets::set keeps the maximum, for every byte of every pattern (:634-641)
      if (ci_) set_offset(i, flip_ascii_case(p[i]));
      if (found) continue;
      if (rare_set_[p[i]]) { found = true; continue; }  // an already chosen rare byte occurs in this pattern
      uint8_t r = rank_of(p[i]);
      if (r < best_rank) { best = p[i]; best_rank = r; }
    }
    if (found) return;
    mark_rare(best);
    if (ci_) mark_rare(flip_ascii_case(best));
  }
  void set_offset(uint64_t pos, uint8_t b) { if (pos > max_off_[b]) max_off_[b] = uint8_t(pos); }
  bool rare_available() const {  // RareBytesBuilder::build :535-575
    return rare_ok_ && rare_count_ <= 3 && rare_count_ != 0;
  }

  bool ci_, wants_packed_;
  bool enabled_ = true;
  uint64_t count_ = 0;
  bool start_set_[256] = {false};
  unsigned start_count_ = 0;
  uint16_t start_rank_sum_ = 0;
  bool rare_set_[256] = {false};
  uint8_t max_off_[256] = {0};
  bool rare_ok_ = true;
  unsigned rare_count_ = 0;
  uint16_t rare_rank_sum_ = 0;
  bool packed_inert_ = false;
  std::vector<uint64_t> packed_lens_;
};

}  // namespace

int build_dfa(const std::vector<PatternRef>& patterns, const BuildOptions& opts, HostDfa* out) {
  // ACB_BUILD_TRACE=1: phase times on stderr (where a 100 000-pattern build spends its second)
  static const bool trace = std::getenv("ACB_BUILD_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "acb200 build: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  const bool leftmost = opts.match_kind != kStandard;
  const bool ci = opts.ascii_case_insensitive;
  HostDfa& d = *out;
  d = HostDfa{};
  d.match_kind = opts.match_kind;
  d.start_kind = opts.start_kind;

  Trie t;
  for (int i = 0; i < 4; ++i) t.add_node();
  t.fail[kDead] = kDead;
  t.fail[kFailId] = kDead;       // allocated before the start id was known (:977-982)
  t.fail[kRoot] = kDead;
  t.fail[kAnchoredRoot] = kDead;  // :1584
  std::vector<uint32_t> depth(4, 0);

  bool boundary[256] = {false};  // ByteClassSet, src/util/alphabet.rs:207-230
  auto mark_byte = [&](uint8_t b) { if (b > 0) boundary[b - 1] = true; boundary[b] = true; };

  PrefilterChooser chooser(opts.match_kind, ci);

  // ---- trie (src/nfa/noncontiguous.rs:1057-1150) ----
  if (patterns.size() > kMaxIndex + 1) return ACG_E_PATTERN_ID_OVERFLOW;
  for (size_t pid = 0; pid < patterns.size(); ++pid) {
    const uint8_t* p = patterns[pid].p;
    const uint64_t n = patterns[pid].n;
    if (n > kMaxIndex) return ACG_E_PATTERN_TOO_LONG;
    d.min_pattern_len = std::min<uint64_t>(d.min_pattern_len, n);
    d.max_pattern_len = std::max<uint64_t>(d.max_pattern_len, n);
    d.pattern_lens.push_back(uint32_t(n));
    if (opts.prefilter) chooser.add(p, n);
    uint32_t cur = kRoot;
    bool passed_match = false, shadowed = false;
    for (uint64_t i = 0; i < n; ++i) {
      passed_match = passed_match || !t.hits[cur].empty();
      if (opts.match_kind == kLeftmostFirst && passed_match) { shadowed = true; break; }  // :1109-1114
      const uint8_t b = p[i];
      mark_byte(b);
      if (ci) mark_byte(flip_ascii_case(b));
      int e = t.find(cur, b);
      if (e >= 0) {
        cur = t.edges[cur][size_t(e)].to;
      } else {
        if (t.edges.size() > kMaxIndex) return ACG_E_STATE_ID_OVERFLOW;
        uint32_t nn = t.add_node();
        depth.push_back(uint32_t(i + 1));
        t.link(cur, b, nn);
        if (ci) t.link(cur, flip_ascii_case(b), nn);
        cur = nn;
      }
    }
    if (!shadowed) t.hits[cur].push_back(uint32_t(pid));
  }
  const size_t ns = t.edges.size();
  lap("trie");

  // byte classes (src/util/alphabet.rs:235-250); `byte_classes(false)` => singletons (src/dfa.rs:436-440)
  uint8_t nfa_classes[256];
  {
    unsigned c = 0;
    for (unsigned b = 0;; ++b) { nfa_classes[b] = uint8_t(c); if (b == 255) break; if (boundary[b]) ++c; }
  }
  if (opts.byte_classes) std::memcpy(d.classes, nfa_classes, 256);
  else for (int b = 0; b < 256; ++b) d.classes[b] = uint8_t(b);
  d.alphabet_len = uint32_t(d.classes[255]) + 1;
  d.stride2 = 0;
  while ((1u << d.stride2) < d.alphabet_len) ++d.stride2;
  const uint32_t s2 = d.stride2, stride = 1u << s2;

  // anchored start mirrors the root's edges and matches (:1561-1586)
  t.edges[kAnchoredRoot] = t.edges[kRoot];
  t.hits[kAnchoredRoot] = t.hits[kRoot];

  // ---- failure links + match propagation (:1275-1374), BFS in ascending byte order ----
  std::vector<uint32_t> bfs;  // visit order, reused to fill DFA rows
  bfs.reserve(ns);
  {
    std::vector<uint8_t> seen(ci ? ns : 0, 0);
    for (const Edge& e : t.edges[kRoot]) {
      if (e.to == kRoot || (ci && seen[e.to])) continue;
      bfs.push_back(e.to);
      if (ci) seen[e.to] = 1;
      if (leftmost && !t.hits[e.to].empty()) t.fail[e.to] = kDead;
    }
    for (size_t qi = 0; qi < bfs.size(); ++qi) {
      const uint32_t v = bfs[qi];
      for (const Edge& e : t.edges[v]) {
        if (ci && seen[e.to]) continue;
        bfs.push_back(e.to);
        if (ci) seen[e.to] = 1;
        if (leftmost && !t.hits[e.to].empty()) { t.fail[e.to] = kDead; continue; }
        uint32_t f = t.fail[v];
        while (t.step(f, e.byte) == kFailId) f = t.fail[f];
        f = t.step(f, e.byte);
        t.fail[e.to] = f;
        t.hits[e.to].insert(t.hits[e.to].end(), t.hits[f].begin(), t.hits[f].end());
        if (t.hits[e.to].size() > kMaxIndex) return ACG_E_STATE_ID_OVERFLOW;
      }
      if (!leftmost && !t.hits[kRoot].empty())
        t.hits[v].insert(t.hits[v].end(), t.hits[kRoot].begin(), t.hits[kRoot].end());  // :1368-1371
    }
  }
  t.root_loop_closed = leftmost && !t.hits[kRoot].empty();
  lap("failure links + match lists");

  // ---- state permutation: DEAD, FAIL, MATCH.., START_U, START_A, NON-MATCH.. (:1399-1481) ----
  // `slot[pos]` = trie node sitting at state index pos after the reference's swap sequence.
  std::vector<uint32_t> slot(ns), newid(ns);
  for (size_t i = 0; i < ns; ++i) slot[i] = uint32_t(i);
  uint32_t next_avail = 4;
  for (size_t i = 4; i < ns; ++i) {
    if (t.hits[slot[i]].empty()) continue;
    std::swap(slot[i], slot[next_avail]);
    ++next_avail;
  }
  const uint32_t n_start_a = next_avail - 1, n_start_u = next_avail - 2;
  std::swap(slot[3], slot[n_start_a]);
  std::swap(slot[2], slot[n_start_u]);
  uint32_t n_max_match = next_avail - 3;
  if (!t.hits[kAnchoredRoot].empty()) n_max_match = n_start_a;
  for (size_t i = 0; i < ns; ++i) newid[slot[i]] = uint32_t(i);

  lap("state permutation");
  d.prefilter_kind = chooser.choose(&d.packed);
  chooser.byte_set(d.prefilter_kind, &d.pre_n, d.pre_byte, d.pre_back);
  const uint32_t n_max_special = d.prefilter_kind != kPreNone ? n_start_a : n_max_match;  // :1036-1045

  // auto-selection as reported by AhoCorasick::kind() (src/ahocorasick.rs:2213-2261)
  d.reported_kind = opts.kind != 0 ? opts.kind
                    : (opts.start_kind != kStartBoth && patterns.size() <= 100) ? ACG_KIND_DFA
                                                                                 : ACG_KIND_CONTIGUOUS_NFA;

  // ---- dense table (src/dfa.rs:431-724) ----
  const bool both = opts.start_kind == kStartBoth;
  d.state_len = both ? ns * 2 - 4 : ns;
  const uint64_t trans_len = uint64_t(d.state_len) << s2;
  if (trans_len - stride > kMaxIndex) return ACG_E_STATE_ID_OVERFLOW;  // :462-478
  d.trans_len = trans_len;
  const size_t n_match_states = both ? size_t(n_max_match - 1) * 2 : size_t(n_max_match - 1);
  std::vector<std::vector<uint32_t>> mlists(n_match_states);

  auto rep_class = [&](uint8_t b) { return d.classes[b]; };

  if (!both && opts.defer_dense && opts.start_kind == kStartUnanchored) {
    // same rows as the branch below, described instead of written
    DenseFillPlan& f = d.fill;
    f.valid = true;
    const size_t n_rows = bfs.size() + 2;
    f.row.reserve(n_rows); f.inherit_row.reserve(n_rows); f.fill_id.reserve(n_rows); f.edge_off.reserve(n_rows + 1);
    auto add_row = [&](uint32_t node, uint32_t inherit, uint32_t fill) {
      f.row.push_back(newid[node]);
      f.inherit_row.push_back(inherit);
      f.fill_id.push_back(fill);
      f.edge_off.push_back(uint32_t(f.edge_to.size()));
      for (const Edge& e : t.edges[node]) {
        f.edge_class.push_back(rep_class(e.byte));
        f.edge_to.push_back(newid[e.to] << s2);
      }
    };
    f.level_off.push_back(0);
    add_row(kRoot, UINT32_MAX, t.root_loop_closed ? kDead : (newid[kRoot] << s2));
    add_row(kAnchoredRoot, UINT32_MAX, kDead);
    uint32_t cur_depth = 0;
    for (uint32_t v : bfs) {
      if (depth[v] != cur_depth) {  // BFS order: depths never decrease
        cur_depth = depth[v];
        f.level_off.push_back(uint32_t(f.row.size()));
      }
      add_row(v, t.fail[v] != kDead ? newid[t.fail[v]] : UINT32_MAX, kDead);
    }
    f.level_off.push_back(uint32_t(f.row.size()));
    f.edge_off.push_back(uint32_t(f.edge_to.size()));
    auto add_shallow = [&](uint32_t node) {
      for (const Edge& e : t.edges[node])
        f.shallow.push_back(DenseFillPlan::ShallowEdge{newid[node], e.byte, newid[e.to]});
    };
    add_shallow(kRoot);
    for (uint32_t v : bfs) {
      if (depth[v] >= 5) break;
      add_shallow(v);
    }
    for (size_t pos = 2; pos <= n_max_match && pos < ns; ++pos) {
      const auto& h = t.hits[slot[pos]];
      if (!h.empty()) mlists[pos - 2] = h;
    }
    d.row_depth.assign(ns, 0xFFFF);
    d.row_depth[newid[kRoot]] = 0;
    for (size_t v = 4; v < ns; ++v) d.row_depth[newid[v]] = uint16_t(std::min<uint32_t>(depth[v], 0xFFFE));
    d.max_special_id = n_max_special << s2;
    d.max_match_id = n_max_match << s2;
    d.start_unanchored_id = n_start_u << s2;
    d.start_anchored_id = kDead;
  } else if (!both) {
    const bool anchored = opts.start_kind == kStartAnchored;
    d.trans.assign(size_t(trans_len), kDead);
    uint32_t* T = d.trans.data();
    auto row_of = [&](uint32_t node) { return T + (size_t(newid[node]) << s2); };
    auto overlay_edges = [&](uint32_t node, uint32_t* row) {
      for (const Edge& e : t.edges[node]) row[rep_class(e.byte)] = newid[e.to] << s2;
    };
    // root: explicit loop (or DEAD once closed), then edges
    {
      uint32_t* row = row_of(kRoot);
      const uint32_t self = t.root_loop_closed ? kDead : (newid[kRoot] << s2);
      for (uint32_t c = 0; c < d.alphabet_len; ++c) row[c] = self;
      overlay_edges(kRoot, row);
    }
    overlay_edges(kAnchoredRoot, row_of(kAnchoredRoot));  // undefined -> DEAD (fail == DEAD)
    for (uint32_t v : bfs) {
      uint32_t* row = row_of(v);
      if (!anchored && t.fail[v] != kDead)
        std::memcpy(row, row_of(t.fail[v]), sizeof(uint32_t) * d.alphabet_len);  // inherit delta(fail, .)
      overlay_edges(v, row);
    }
    for (size_t pos = 2; pos <= n_max_match && pos < ns; ++pos) {
      const auto& h = t.hits[slot[pos]];
      if (!h.empty()) mlists[pos - 2] = h;
    }
    if (!anchored) {
      d.row_depth.assign(ns, 0xFFFF);
      d.row_depth[newid[kRoot]] = 0;
      for (size_t v = 4; v < ns; ++v) d.row_depth[newid[v]] = uint16_t(std::min<uint32_t>(depth[v], 0xFFFE));
    }
    d.max_special_id = n_max_special << s2;
    d.max_match_id = n_max_match << s2;
    d.start_unanchored_id = anchored ? kDead : (n_start_u << s2);
    d.start_anchored_id = anchored ? (n_start_a << s2) : kDead;
  } else {
    // unanchored delta in NFA-id space first (row inheritance), then interleave per :617-724
    const uint32_t alen = d.alphabet_len;
    std::vector<uint32_t> U(size_t(ns) * alen, kDead);
    auto urow = [&](uint32_t node) { return U.data() + size_t(newid[node]) * alen; };
    {
      uint32_t* row = urow(kRoot);
      const uint32_t self = t.root_loop_closed ? kDead : newid[kRoot];
      for (uint32_t c = 0; c < alen; ++c) row[c] = self;
      for (const Edge& e : t.edges[kRoot]) row[rep_class(e.byte)] = newid[e.to];
    }
    for (uint32_t v : bfs) {
      uint32_t* row = urow(v);
      if (t.fail[v] != kDead) std::memcpy(row, urow(t.fail[v]), sizeof(uint32_t) * alen);
      for (const Edge& e : t.edges[v]) row[rep_class(e.byte)] = newid[e.to];
    }
    std::vector<uint32_t> map_u(ns, kDead), map_a(ns, kDead);
    uint32_t next_sid = 0;
    for (size_t pos = 0; pos < ns; ++pos) {
      if (pos == kDead || pos == kFailId) { map_u[pos] = map_a[pos] = next_sid; next_sid += stride; }
      else if (pos == n_start_u) { map_u[pos] = next_sid; next_sid += stride; }
      else if (pos == n_start_a) { map_a[pos] = next_sid; next_sid += stride; }
      else { map_u[pos] = next_sid; next_sid += stride; map_a[pos] = next_sid; next_sid += stride; }
    }
    d.trans.assign(size_t(trans_len), kDead);
    uint32_t* T = d.trans.data();
    for (size_t pos = 2; pos < ns; ++pos) {
      const uint32_t node = slot[pos];
      if (pos != n_start_a) {
        uint32_t* row = T + map_u[pos];
        const uint32_t* src = U.data() + pos * alen;
        for (uint32_t c = 0; c < alen; ++c) row[c] = map_u[src[c]];
        if (!t.hits[node].empty()) mlists[(map_u[pos] >> s2) - 2] = t.hits[node];
      }
      if (pos != n_start_u) {
        uint32_t* row = T + map_a[pos];
        for (const Edge& e : t.edges[node]) row[rep_class(e.byte)] = map_a[newid[e.to]];
        if (!t.hits[node].empty()) mlists[(map_a[pos] >> s2) - 2] = t.hits[node];
      }
    }
    d.max_special_id = map_a[n_max_special];
    d.max_match_id = map_a[n_max_match];
    d.start_unanchored_id = map_u[n_start_u];
    d.start_anchored_id = map_a[n_start_a];
  }

  lap("dense table / fill plan");
  // CSR of `matches: Vec<Vec<PatternID>>` (src/dfa.rs:96-99)
  d.match_offsets.assign(n_match_states + 1, 0);
  size_t total = 0;
  for (size_t i = 0; i < n_match_states; ++i) { d.match_offsets[i] = uint32_t(total); total += mlists[i].size(); }
  d.match_offsets[n_match_states] = uint32_t(total);
  d.match_pids.reserve(total);
  for (auto& l : mlists) d.match_pids.insert(d.match_pids.end(), l.begin(), l.end());
  lap("match CSR");
  return 0;
}

}  // namespace acb
