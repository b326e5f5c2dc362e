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

constexpr uint32_t kNone = 0xFFFFFFFFu;

// Explicit trie on flat arrays (a 100 000-pattern set has 8 * 10^5 nodes: one heap block per node
// and list was most of the build).  Node ids are allocated exactly like the reference allocates
// NFA states (src/nfa/noncontiguous.rs:977-985, 1132-1143): 0 DEAD, 1 FAIL, 2 unanchored start,
// 3 anchored start, then one node per new trie edge in pattern order.
//
// Two phases.  While patterns are inserted, a node's edges are a list sorted by byte threaded
// through one pool (nodes with many children additionally get a 256-entry direct table, so that
// the hot shallow nodes are one load per step).  freeze() then lays the edges out contiguously per
// node (CSR); from there on edges_of() / find() / step() read that form.
//
// Match lists: the patterns that end at a node ("own", insertion order) are threaded through
// pid_next; the lists the failure-link pass composes (own ++ list(fail), src/nfa/noncontiguous.rs:
// 490-523) live in one pool as immutable segments (hoff, hlen) -- a node without own patterns
// shares its failure state's segment.
struct Trie {
  struct PoolEdge { uint32_t to, next; uint8_t byte; };
  struct BuildNode {             // insertion phase
    uint32_t first_edge = kNone;  // head of the node's sorted edge list
    uint32_t dense_of = kNone;    // index of its direct table
    uint32_t own_head = kNone, own_tail = kNone;  // first / last pattern ending here
    uint32_t n_edges = 0;
    uint32_t depth = 0;
  };
  struct Node {                  // frozen form: everything a visit needs in one 32-byte record
    uint32_t eoff = 0, ecnt = 0;  // its edges in `edges`
    uint32_t fail = kRoot;
    uint32_t hoff = 0, hlen = 0;  // its composed match list in hpool
    uint32_t own_head = kNone;
    uint32_t dense_of = kNone;
    uint32_t depth = 0;
  };
  std::vector<BuildNode> bn;
  std::vector<PoolEdge> pool;
  std::vector<uint32_t> dense;     // [tables][256]: child + 1, 0 = no edge
  std::vector<uint32_t> pid_next;  // per pattern: next pattern ending at the same node
  std::vector<Node> nodes;
  std::vector<Edge> edges;         // sorted by byte within a node
  std::vector<uint32_t> hpool;
  bool root_loop_closed = false;  // src/nfa/noncontiguous.rs:1620-1638

  static constexpr unsigned kDenseAt = 8;  // children beyond which a node gets a direct table

  size_t size() const { return bn.size(); }
  uint32_t add_node() {
    bn.emplace_back();
    return uint32_t(bn.size() - 1);
  }
  void reserve(size_t n_nodes, size_t pats) {
    bn.reserve(n_nodes);
    pool.reserve(n_nodes);
    pid_next.reserve(pats);
  }
  // ---- insertion phase ----
  uint32_t child(uint32_t s, uint8_t b) const {  // kNone if there is no edge
    const BuildNode& n = bn[s];
    if (n.dense_of != kNone) return dense[size_t(n.dense_of) * 256 + b] - 1u;  // 0 - 1 == kNone
    for (uint32_t e = n.first_edge; e != kNone; e = pool[e].next) {
      const PoolEdge& pe = pool[e];
      if (pe.byte >= b) return pe.byte == b ? pe.to : kNone;
    }
    return kNone;
  }
  void link(uint32_t s, uint8_t b, uint32_t to) {
    BuildNode& n = bn[s];
    uint32_t prev = kNone, e = n.first_edge;
    while (e != kNone && pool[e].byte < b) { prev = e; e = pool[e].next; }
    if (e != kNone && pool[e].byte == b) {
      pool[e].to = to;
    } else {
      const uint32_t ne = uint32_t(pool.size());
      pool.push_back(PoolEdge{to, e, b});
      if (prev == kNone) n.first_edge = ne; else pool[prev].next = ne;
      ++n.n_edges;
    }
    if (n.dense_of == kNone && n.n_edges > kDenseAt) {
      n.dense_of = uint32_t(dense.size() / 256);
      dense.resize(dense.size() + 256, 0);
      for (uint32_t x = n.first_edge; x != kNone; x = pool[x].next) dense[size_t(n.dense_of) * 256 + pool[x].byte] = pool[x].to + 1;
    } else if (n.dense_of != kNone) {
      dense[size_t(n.dense_of) * 256 + b] = to + 1;
    }
  }
  bool has_own(uint32_t s) const { return bn[s].own_head != kNone; }
  void add_own(uint32_t s, uint32_t pid) {  // pids arrive in ascending order, one call per pattern
    if (pid_next.size() <= pid) pid_next.resize(size_t(pid) + 1, kNone);
    BuildNode& n = bn[s];
    if (n.own_tail == kNone) n.own_head = pid; else pid_next[n.own_tail] = pid;
    n.own_tail = pid;
  }
  // ---- frozen form ----
  void freeze() {
    const size_t ns = size();
    nodes.assign(ns, Node{});
    edges.clear();
    edges.reserve(pool.size() + 256);
    for (size_t v = 0; v < ns; ++v) {
      // (a node's edges sit wherever its children were created: ask for them a few nodes ahead)
      if (v + 16 < ns && bn[v + 16].first_edge != kNone) __builtin_prefetch(&pool[bn[v + 16].first_edge]);
      // the anchored start mirrors the root's edges (src/nfa/noncontiguous.rs:1561-1586)
      const BuildNode& src = bn[v == kAnchoredRoot ? kRoot : v];
      Node& n = nodes[v];
      n.eoff = uint32_t(edges.size());
      for (uint32_t e = src.first_edge; e != kNone; e = pool[e].next) edges.push_back(Edge{pool[e].byte, pool[e].to});
      n.ecnt = uint32_t(edges.size()) - n.eoff;
      n.own_head = bn[v].own_head;
      n.depth = bn[v].depth;
      n.dense_of = src.dense_of;
    }
    std::vector<BuildNode>().swap(bn);
    std::vector<PoolEdge>().swap(pool);
  }
  struct EdgeSpan {
    const Edge* b; const Edge* e;
    const Edge* begin() const { return b; }
    const Edge* end() const { return e; }
  };
  EdgeSpan edges_of(uint32_t s) const {
    const Edge* b = edges.data() + nodes[s].eoff;
    return EdgeSpan{b, b + nodes[s].ecnt};
  }
  uint32_t& fail(uint32_t s) { return nodes[s].fail; }
  uint32_t fail(uint32_t s) const { return nodes[s].fail; }
  uint32_t find(uint32_t s, uint8_t b) const {  // kNone if there is no edge
    const Node& n = nodes[s];
    if (n.dense_of != kNone) return dense[size_t(n.dense_of) * 256 + b] - 1u;
    for (const Edge* e = edges.data() + n.eoff, *end = e + n.ecnt; e != end; ++e)
      if (e->byte >= b) return e->byte == b ? e->to : kNone;
    return kNone;
  }
  // goto function used while computing failure links: the root loops on
  // undefined bytes (src/nfa/noncontiguous.rs:1597-1606), DEAD is absorbing
  // (:1643-1646), everything else reports "undefined" as kFailId.
  uint32_t step(uint32_t s, uint8_t b) const {
    if (s == kDead) return kDead;
    const uint32_t to = find(s, b);
    if (to != kNone) return to;
    return s == kRoot ? kRoot : kFailId;
  }
  // composed match lists
  bool has_hits(uint32_t s) const { return nodes[s].hlen != 0; }
  uint32_t hlen(uint32_t s) const { return nodes[s].hlen; }
  void share_list(uint32_t s, uint32_t of) { nodes[s].hoff = nodes[of].hoff; nodes[s].hlen = nodes[of].hlen; }
  void start_list(uint32_t s) {  // list(s) = own(s), at the end of the pool
    Node& n = nodes[s];
    n.hoff = uint32_t(hpool.size());
    uint32_t c = 0;
    for (uint32_t p = n.own_head; p != kNone; p = pid_next[p]) { hpool.push_back(p); ++c; }
    n.hlen = c;
  }
  // list(s) ++= list(f).  list(s) is either empty or the pool's last segment (start_list just ran).
  void append_list(uint32_t s, uint32_t f) {
    if (nodes[f].hlen == 0) return;
    if (nodes[s].hlen == 0) { share_list(s, f); return; }
    append_copy(s, f);
  }
  // the same for a list(s) that may sit anywhere: a fresh segment list(s) ++ list(f)
  void extend_list(uint32_t s, uint32_t f) {
    if (nodes[f].hlen == 0) return;
    Node& n = nodes[s];
    if (n.hlen == 0) { share_list(s, f); return; }
    if (size_t(n.hoff) + n.hlen != hpool.size()) {
      const size_t from = n.hoff, cnt = n.hlen;
      n.hoff = uint32_t(hpool.size());
      hpool.resize(hpool.size() + cnt);
      std::memmove(hpool.data() + n.hoff, hpool.data() + from, cnt * sizeof(uint32_t));
    }
    append_copy(s, f);
  }
  const uint32_t* list(uint32_t s) const { return hpool.data() + nodes[s].hoff; }

 private:
  void append_copy(uint32_t s, uint32_t f) {
    const size_t from = nodes[f].hoff, cnt = nodes[f].hlen, at = hpool.size();
    hpool.resize(at + cnt);  // may move the pool: copy by index afterwards
    std::memmove(hpool.data() + at, hpool.data() + from, cnt * sizeof(uint32_t));
    nodes[s].hlen += uint32_t(cnt);
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
  {
    size_t total_bytes = 0;
    for (const PatternRef& pr : patterns) total_bytes += size_t(pr.n);
    t.reserve(total_bytes + 4, patterns.size());
  }
  for (int i = 0; i < 4; ++i) t.add_node();

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
      passed_match = passed_match || t.has_own(cur);
      if (opts.match_kind == kLeftmostFirst && passed_match) { shadowed = true; break; }  // :1109-1114
      const uint8_t b = p[i];
      mark_byte(b);
      if (ci) mark_byte(flip_ascii_case(b));
      const uint32_t to = t.child(cur, b);
      if (to != kNone) {
        cur = to;
      } else {
        if (t.size() > kMaxIndex) return ACG_E_STATE_ID_OVERFLOW;
        uint32_t nn = t.add_node();
        t.bn[nn].depth = uint32_t(i + 1);
        t.link(cur, b, nn);
        if (ci) t.link(cur, flip_ascii_case(b), nn);
        cur = nn;
      }
    }
    if (!shadowed) t.add_own(cur, uint32_t(pid));
  }
  const size_t ns = t.size();
  lap("trie");
  t.freeze();  // the anchored start mirrors the root's edges (:1561-1586)
  t.fail(kDead) = kDead;
  t.fail(kFailId) = kDead;       // allocated before the start id was known (:977-982)
  t.fail(kRoot) = kDead;
  t.fail(kAnchoredRoot) = kDead;  // :1584
  lap("trie: contiguous form");

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

  // ... and its matches
  t.start_list(kRoot);
  t.share_list(kAnchoredRoot, kRoot);

  // ---- failure links + match propagation (:1275-1374), BFS in ascending byte order ----
  std::vector<uint32_t> bfs;  // visit order, reused to fill DFA rows
  bfs.reserve(ns);
  {
    std::vector<uint8_t> seen(ci ? ns : 0, 0);
    for (const Edge& e : t.edges_of(kRoot)) {
      if (e.to == kRoot || (ci && seen[e.to])) continue;
      bfs.push_back(e.to);
      if (ci) seen[e.to] = 1;
      t.start_list(e.to);
      if (leftmost && t.has_hits(e.to)) t.fail(e.to) = kDead;
    }
    const bool root_hits = !leftmost && t.has_hits(kRoot);
    for (size_t qi = 0; qi < bfs.size(); ++qi) {
      // the queue is known ahead: node records eight entries on, their edges four on, the records of
      // those edges' targets two on -- by the time a node is dequeued its lines are in flight or here
      if (qi + 8 < bfs.size()) __builtin_prefetch(&t.nodes[bfs[qi + 8]]);
      if (qi + 4 < bfs.size()) __builtin_prefetch(t.edges.data() + t.nodes[bfs[qi + 4]].eoff);
      if (qi + 2 < bfs.size())
        for (const Edge& e : t.edges_of(bfs[qi + 2])) __builtin_prefetch(&t.nodes[e.to], 1);
      const uint32_t v = bfs[qi];
      const uint32_t fv = t.fail(v);
      for (const Edge& e : t.edges_of(v)) {
        if (ci && seen[e.to]) continue;
        bfs.push_back(e.to);
        if (ci) seen[e.to] = 1;
        t.start_list(e.to);
        if (leftmost && t.has_hits(e.to)) { t.fail(e.to) = kDead; continue; }
        uint32_t f = fv;
        while (t.step(f, e.byte) == kFailId) f = t.fail(f);
        f = t.step(f, e.byte);
        t.fail(e.to) = f;
        t.append_list(e.to, f);
        if (t.hlen(e.to) > kMaxIndex || t.hpool.size() > kMaxIndex) return ACG_E_STATE_ID_OVERFLOW;
      }
      if (root_hits) t.extend_list(v, kRoot);  // :1368-1371
    }
  }
  t.root_loop_closed = leftmost && t.has_hits(kRoot);
  lap("failure links + match lists");

  // ---- state permutation: DEAD, FAIL, MATCH.., START_U, START_A, NON-MATCH.. (:1399-1481) ----
  // `slot[pos]` = trie node sitting at state index pos after the reference's swap sequence.
  std::vector<uint32_t> slot(ns), newid(ns);
  for (size_t i = 0; i < ns; ++i) slot[i] = uint32_t(i);
  uint32_t next_avail = 4;
  for (size_t i = 4; i < ns; ++i) {
    if (!t.has_hits(slot[i])) continue;
    std::swap(slot[i], slot[next_avail]);
    ++next_avail;
  }
  const uint32_t n_start_a = next_avail - 1, n_start_u = next_avail - 2;
  std::swap(slot[3], slot[n_start_a]);
  std::swap(slot[2], slot[n_start_u]);
  uint32_t n_max_match = next_avail - 3;
  if (t.has_hits(kAnchoredRoot)) n_max_match = n_start_a;
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
  std::vector<uint32_t> mnode(n_match_states, kNone);  // trie node whose list the match state reports

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
      for (const Edge& e : t.edges_of(node)) {
        f.edge_class.push_back(rep_class(e.byte));
        f.edge_to.push_back(newid[e.to] << s2);
      }
    };
    f.level_off.push_back(0);
    add_row(kRoot, UINT32_MAX, t.root_loop_closed ? kDead : (newid[kRoot] << s2));
    add_row(kAnchoredRoot, UINT32_MAX, kDead);
    uint32_t cur_depth = 0;
    f.edge_class.reserve(t.edges.size());
    f.edge_to.reserve(t.edges.size());
    for (size_t qi = 0; qi < bfs.size(); ++qi) {
      if (qi + 8 < bfs.size()) { __builtin_prefetch(&t.nodes[bfs[qi + 8]]); __builtin_prefetch(&newid[bfs[qi + 8]]); }
      if (qi + 4 < bfs.size()) __builtin_prefetch(t.edges.data() + t.nodes[bfs[qi + 4]].eoff);
      if (qi + 2 < bfs.size())
        for (const Edge& e : t.edges_of(bfs[qi + 2])) __builtin_prefetch(&newid[e.to]);
      const uint32_t v = bfs[qi];
      const uint32_t dv = t.nodes[v].depth;
      if (dv != cur_depth) {  // BFS order: depths never decrease
        cur_depth = dv;
        f.level_off.push_back(uint32_t(f.row.size()));
      }
      add_row(v, t.fail(v) != kDead ? newid[t.fail(v)] : UINT32_MAX, kDead);
    }
    f.level_off.push_back(uint32_t(f.row.size()));
    f.edge_off.push_back(uint32_t(f.edge_to.size()));
    auto add_shallow = [&](uint32_t node) {
      for (const Edge& e : t.edges_of(node))
        f.shallow.push_back(DenseFillPlan::ShallowEdge{newid[node], e.byte, newid[e.to]});
    };
    add_shallow(kRoot);
    for (uint32_t v : bfs) {
      if (t.nodes[v].depth >= 5) break;
      add_shallow(v);
    }
    for (size_t pos = 2; pos <= n_max_match && pos < ns; ++pos)
      if (t.has_hits(slot[pos])) mnode[pos - 2] = slot[pos];
    d.row_depth.assign(ns, 0xFFFF);
    d.row_depth[newid[kRoot]] = 0;
    for (size_t v = 4; v < ns; ++v) d.row_depth[newid[v]] = uint16_t(std::min<uint32_t>(t.nodes[v].depth, 0xFFFE));
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
      for (const Edge& e : t.edges_of(node)) row[rep_class(e.byte)] = newid[e.to] << s2;
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
      if (!anchored && t.fail(v) != kDead)
        std::memcpy(row, row_of(t.fail(v)), sizeof(uint32_t) * d.alphabet_len);  // inherit delta(fail, .)
      overlay_edges(v, row);
    }
    for (size_t pos = 2; pos <= n_max_match && pos < ns; ++pos)
      if (t.has_hits(slot[pos])) mnode[pos - 2] = slot[pos];
    if (!anchored) {
      d.row_depth.assign(ns, 0xFFFF);
      d.row_depth[newid[kRoot]] = 0;
      for (size_t v = 4; v < ns; ++v) d.row_depth[newid[v]] = uint16_t(std::min<uint32_t>(t.nodes[v].depth, 0xFFFE));
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
      for (const Edge& e : t.edges_of(kRoot)) row[rep_class(e.byte)] = newid[e.to];
    }
    for (uint32_t v : bfs) {
      uint32_t* row = urow(v);
      if (t.fail(v) != kDead) std::memcpy(row, urow(t.fail(v)), sizeof(uint32_t) * alen);
      for (const Edge& e : t.edges_of(v)) row[rep_class(e.byte)] = newid[e.to];
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
        if (t.has_hits(node)) mnode[(map_u[pos] >> s2) - 2] = node;
      }
      if (pos != n_start_u) {
        uint32_t* row = T + map_a[pos];
        for (const Edge& e : t.edges_of(node)) row[rep_class(e.byte)] = map_a[newid[e.to]];
        if (t.has_hits(node)) mnode[(map_a[pos] >> s2) - 2] = node;
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
  for (size_t i = 0; i < n_match_states; ++i) {
    d.match_offsets[i] = uint32_t(total);
    if (mnode[i] != kNone) total += t.hlen(mnode[i]);
  }
  d.match_offsets[n_match_states] = uint32_t(total);
  d.match_pids.resize(total);
  for (size_t i = 0; i < n_match_states; ++i)
    if (mnode[i] != kNone)
      std::memcpy(d.match_pids.data() + d.match_offsets[i], t.list(mnode[i]), size_t(t.hlen(mnode[i])) * sizeof(uint32_t));
  lap("match CSR");
  return 0;
}

}  // namespace acb
