/*
 * ac_oracle.c -- CPU oracle (plain C restatement of BurntSushi/aho-corasick
 * 1.1.3) for the DFA-scan / packed hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see ac_oracle.h.  Parity status: pinned against
 * the reference's golden vectors (tests/test_oracle_golden.py).
 *
 * Layout of this file (reference file:line each part follows):
 *   1. byte classes            src/util/alphabet.rs:207-250, :45-62
 *   2. prefilter *decision*    src/util/prefilter.rs:163-305, 397-610, 760-853
 *   3. packed: Patterns        src/packed/pattern.rs:84-99, 265-281
 *              Rabin-Karp      src/packed/rabinkarp.rs:41-152
 *              Teddy (scalar emulation of the vector algorithm)
 *                              src/packed/teddy/builder.rs:98-231
 *                              src/packed/teddy/generic.rs:114-713, 751-870,
 *                              911-997, 1039-1162, 1178-1368
 *              Searcher        src/packed/api.rs:253-322, 529-546, 661-687
 *   4. noncontiguous NFA build src/nfa/noncontiguous.rs:381-585, 963-1646
 *      remapper                src/util/remapper.rs:86-150
 *   5. DFA build               src/dfa.rs:431-724, 801-835
 *   6. search loops/iterators  src/automaton.rs:844-970, 1259-1549
 *   7. facade                  src/ahocorasick.rs:2171-2261, 2778-2789
 */
#include "ac_oracle.h"

#include <stdlib.h>
#include <string.h>

#define DEAD 0u
#define FAIL 1u
#define SMALL_INDEX_MAX 0x7FFFFFFEu /* src/util/primitives.rs:96-111: i32::MAX - 1 */

#define VEC(T) struct { T* p; size_t n, cap; }
#define VPUSH(v, x)                                                        \
  do {                                                                     \
    if ((v).n == (v).cap) {                                                \
      (v).cap = (v).cap ? (v).cap * 2 : 16;                                \
      (v).p = realloc((v).p, (v).cap * sizeof(*(v).p));                    \
    }                                                                      \
    (v).p[(v).n++] = (x);                                                  \
  } while (0)

/* ------------------------------------------------------------------ */
/* 1. byte classes                                                     */
/* ------------------------------------------------------------------ */

typedef struct { uint8_t bits[32]; } byteset_t;
static void bs_add(byteset_t* s, uint8_t b) { s->bits[b >> 3] |= (uint8_t)(1u << (b & 7)); }
static int bs_has(const byteset_t* s, uint8_t b) { return (s->bits[b >> 3] >> (b & 7)) & 1; }

/* ByteClassSet::set_range, src/util/alphabet.rs:224-230 */
static void bcs_set_range(byteset_t* s, uint8_t start, uint8_t end) {
  if (start > 0) bs_add(s, (uint8_t)(start - 1));
  bs_add(s, end);
}
/* ByteClassSet::byte_classes, src/util/alphabet.rs:235-250 */
static void bcs_byte_classes(const byteset_t* s, uint8_t out[256]) {
  unsigned cls = 0;
  for (unsigned b = 0;; b++) {
    out[b] = (uint8_t)cls;
    if (b == 255) break;
    if (bs_has(s, (uint8_t)b)) cls++;
  }
}
/* ByteClasses::alphabet_len / stride2, src/util/alphabet.rs:45-62 */
static unsigned bc_alphabet_len(const uint8_t c[256]) { return (unsigned)c[255] + 1; }
static unsigned bc_stride2(const uint8_t c[256]) {
  unsigned a = bc_alphabet_len(c), p = 1, z = 0;
  while (p < a) { p <<= 1; z++; }
  return z;
}

/* opposite_ascii_case, src/util/prefilter.rs:906-914 */
static uint8_t opposite_ascii_case(uint8_t b) {
  if (b >= 'A' && b <= 'Z') return (uint8_t)(b + 32);
  if (b >= 'a' && b <= 'z') return (uint8_t)(b - 32);
  return b;
}

/* ------------------------------------------------------------------ */
/* 2. prefilter decision                                               */
/* ------------------------------------------------------------------ */

/* Heuristic byte-frequency ranks: the 256-entry data table of
 * src/util/byte_frequencies.rs:1-258, hex encoded (index = byte value). */
static const char FREQ_HEX[] =
    "3734333231302f2e2d67f24243e52c2b2a29282726252423222138201f1e1d1cff94a49588a09badddde867ae8cad7e0"
    "d0dcccbbb7b3b1a8b2c8e2c39ab8ae7e78bf9dc2aabda2a196c18e89abb0b9a7ba70afc0bc9c8c8f7b8580938a9272df"
    "97f9d8eeecfde3dae6f787b4f1e9f6f4e78bf5f3fbebc9c4f0d698b6cdb57f1bd4d3d2d5e4c5a99f83ac695062606151"
    "cf917473908299796b846d6e7c6f526c768d7181777da5755c6a5348635d414fa6eda3c7bee1d1cbc6d9dbceeaf89eef"
    "ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff"
    "ffffffffffffffffffffffffffffffff";
static uint8_t freq_rank(uint8_t b) {
  static uint8_t tab[256];
  static int init = 0;
  if (!init) {
    for (int i = 0; i < 256; i++) {
      char h = FREQ_HEX[2 * i], l = FREQ_HEX[2 * i + 1];
      int hv = h <= '9' ? h - '0' : h - 'a' + 10, lv = l <= '9' ? l - '0' : l - 'a' + 10;
      tab[i] = (uint8_t)(hv * 16 + lv);
    }
    init = 1;
  }
  return tab[b];
}

typedef struct {
  int ascii_ci;
  uint8_t byteset[256];
  size_t count;
  uint16_t rank_sum;
} start_bytes_t;

/* StartBytesBuilder::add / add_one_byte, src/util/prefilter.rs:826-852 */
static void sb_add_one(start_bytes_t* s, uint8_t b) {
  if (!s->byteset[b]) {
    s->byteset[b] = 1;
    s->count++;
    s->rank_sum = (uint16_t)(s->rank_sum + freq_rank(b));
  }
}
static void sb_add(start_bytes_t* s, const uint8_t* p, size_t n) {
  if (s->count > 3) return;
  if (n > 0) {
    sb_add_one(s, p[0]);
    if (s->ascii_ci) sb_add_one(s, opposite_ascii_case(p[0]));
  }
}
/* StartBytesBuilder::build, src/util/prefilter.rs:784-824: Some iff 1..=3 bytes, all ASCII */
static int sb_available(const start_bytes_t* s) {
  if (s->count > 3) return 0;
  size_t len = 0;
  for (int b = 0; b < 256; b++) {
    if (!s->byteset[b]) continue;
    if (b > 0x7F) return 0;
    len++;
  }
  return len != 0;
}

typedef struct {
  int ascii_ci;
  byteset_t rare_set;
  int available;
  size_t count;
  uint16_t rank_sum;
} rare_bytes_t;

/* RareBytesBuilder::add_one_rare_byte / add_rare_byte, src/util/prefilter.rs:640-657 */
static void rb_add_one_rare(rare_bytes_t* r, uint8_t b) {
  if (!bs_has(&r->rare_set, b)) {
    bs_add(&r->rare_set, b);
    r->count++;
    r->rank_sum = (uint16_t)(r->rank_sum + freq_rank(b));
  }
}
/* RareBytesBuilder::add, src/util/prefilter.rs:585-630 (byte offsets only steer
 * the memchr skip distance and are not tracked here) */
static void rb_add(rare_bytes_t* r, const uint8_t* p, size_t n) {
  if (!r->available) return;
  if (r->count > 3) { r->available = 0; return; }
  if (n >= 256) { r->available = 0; return; }
  if (n == 0) return;
  uint8_t rarest = p[0];
  uint8_t rarest_rank = freq_rank(p[0]);
  int found = 0;
  for (size_t i = 0; i < n; i++) {
    uint8_t b = p[i];
    if (found) continue;
    if (bs_has(&r->rare_set, b)) { found = 1; continue; }
    uint8_t rank = freq_rank(b);
    if (rank < rarest_rank) { rarest = b; rarest_rank = rank; }
  }
  if (!found) {
    rb_add_one_rare(r, rarest);
    if (r->ascii_ci) rb_add_one_rare(r, opposite_ascii_case(rarest));
  }
}
/* RareBytesBuilder::build, src/util/prefilter.rs:535-575 */
static int rb_available(const rare_bytes_t* r) {
  if (!r->available || r->count > 3) return 0;
  size_t len = 0;
  for (int b = 0; b < 256; b++) if (bs_has(&r->rare_set, (uint8_t)b)) len++;
  return len != 0;
}

/* ------------------------------------------------------------------ */
/* 3. packed                                                           */
/* ------------------------------------------------------------------ */

#define PACKED_PATTERN_LIMIT 128 /* src/packed/api.rs:11 */
#define RK_BUCKETS 64           /* src/packed/rabinkarp.rs:20 */

typedef struct { uint64_t hash; uint32_t pid; } rk_entry;

struct orc_packed {
  int kind;
  /* Patterns, src/packed/pattern.rs:21-60 */
  uint8_t** by_id;
  size_t* lens;
  size_t n;
  uint32_t* order;
  size_t minimum_len_pats;
  /* Rabin-Karp */
  VEC(rk_entry) rk[RK_BUCKETS];
  size_t hash_len;
  uint64_t hash_2pow;
  /* Teddy */
  int has_teddy;
  int nbuckets;  /* 8 slim, 16 fat */
  int width;     /* positions scanned per chunk: 16 or 32 */
  int mask_len;  /* 1..4 */
  VEC(uint32_t) buckets[16];
  uint16_t lo[4][16], hi[4][16];
  size_t minimum_len; /* Searcher.minimum_len, api.rs:268-278 */
};

static int is_prefix(const uint8_t* hay, size_t hay_len, const uint8_t* pat, size_t pat_len) {
  /* Pattern::is_prefix(_raw), src/packed/pattern.rs:249-281 */
  return pat_len <= hay_len && memcmp(hay, pat, pat_len) == 0;
}

static uint64_t rk_hash(const uint8_t* b, size_t n) {
  /* RabinKarp::hash, src/packed/rabinkarp.rs:135-143 (usize == u64) */
  uint64_t h = 0;
  for (size_t i = 0; i < n; i++) h = (h << 1) + b[i];
  return h;
}

/* RabinKarp::find_at, src/packed/rabinkarp.rs:86-117 */
static int rk_find_at(const orc_packed* p, const uint8_t* hay, size_t hay_len, size_t at,
                      orc_match* out) {
  if (at + p->hash_len > hay_len) return 0;
  uint64_t hash = rk_hash(hay + at, p->hash_len);
  for (;;) {
    size_t b = (size_t)(hash % RK_BUCKETS);
    for (size_t i = 0; i < p->rk[b].n; i++) {
      if (p->rk[b].p[i].hash == hash) {
        uint32_t pid = p->rk[b].p[i].pid;
        if (is_prefix(hay + at, hay_len - at, p->by_id[pid], p->lens[pid])) {
          out->pid = pid; out->start = at; out->end = at + p->lens[pid];
          return 1;
        }
      }
    }
    if (at + p->hash_len >= hay_len) return 0;
    /* update_hash, :145-151 */
    hash = ((hash - (uint64_t)hay[at] * p->hash_2pow) << 1) + hay[at + p->hash_len];
    at++;
  }
}

/* Teddy::verify_bucket, src/packed/teddy/generic.rs:849-870 */
static int teddy_verify_bucket(const orc_packed* p, const uint8_t* hay, size_t cur, size_t end,
                               int bucket, orc_match* out) {
  for (size_t i = 0; i < p->buckets[bucket].n; i++) {
    uint32_t pid = p->buckets[bucket].p[i];
    if (is_prefix(hay + cur, end - cur, p->by_id[pid], p->lens[pid])) {
      out->pid = pid; out->start = cur; out->end = cur + p->lens[pid];
      return 1;
    }
  }
  return 0;
}

/* One chunk: Slim/Fat::candidate + Teddy::verify (generic.rs:164-178, 216-236,
 * 283-304, 354-380, 911-997, 1039-1162).  `cur` is the haystack offset of the
 * chunk's lane 0; prev[k][*] carry res_k of the previous chunk. */
static int teddy_find_one(const orc_packed* p, const uint8_t* hay, size_t cur, size_t end,
                          uint16_t prev[3][32], orc_match* out) {
  const int W = p->width, N = p->mask_len;
  uint16_t res[4][32];
  for (int k = 0; k < N; k++)
    for (int j = 0; j < W; j++) {
      uint8_t b = hay[cur + (size_t)j];
      res[k][j] = p->lo[k][b & 0xF] & p->hi[k][b >> 4];
    }
  uint16_t cand[32];
  int any = 0;
  for (int j = 0; j < W; j++) {
    uint16_t c = res[N - 1][j];
    for (int k = 0; k < N - 1; k++) {
      int s = N - 1 - k; /* shift_in_{s}_bytes */
      uint16_t v = (j >= s) ? res[k][j - s] : prev[k][W - s + j];
      c &= v;
    }
    cand[j] = c;
    any |= c;
  }
  for (int k = 0; k < N - 1; k++) memcpy(prev[k], res[k], sizeof(uint16_t) * (size_t)W);
  if (!any) return 0;
  size_t base = cur - (size_t)(N - 1);
  for (int j = 0; j < W; j++) {
    uint16_t c = cand[j];
    for (int b = 0; b < p->nbuckets; b++)
      if ((c >> b) & 1)
        if (teddy_verify_bucket(p, hay, base + (size_t)j, end, b, out)) return 1;
  }
  return 0;
}

/* Slim/Fat::<N>::find, src/packed/teddy/generic.rs:114-160 (N=1), 180-214,
 * 238-281, 306-352 and the Fat twins 432-713 */
static int teddy_find(const orc_packed* p, const uint8_t* hay, size_t start, size_t end,
                      orc_match* out) {
  const size_t W = (size_t)p->width;
  uint16_t prev[3][32];
  memset(prev, 0xFF, sizeof(prev));
  size_t cur = start + (size_t)(p->mask_len - 1);
  while (cur + W <= end) {
    if (teddy_find_one(p, hay, cur, end, prev, out)) return 1;
    cur += W;
  }
  if (cur < end) {
    cur = end - W;
    memset(prev, 0xFF, sizeof(prev));
    if (teddy_find_one(p, hay, cur, end, prev, out)) return 1;
  }
  return 0;
}

void orc_packed_config_default(orc_packed_config* c) {
  c->kind = ORC_PACKED_LEFTMOST_FIRST;
  c->force = ORC_FORCE_NONE;
  c->only_teddy_fat = -1;
  c->only_teddy_256bit = -1;
  c->heuristic_pattern_limits = 1;
}

void orc_packed_free(orc_packed* p) {
  if (!p) return;
  for (size_t i = 0; i < p->n; i++) free(p->by_id[i]);
  free(p->by_id); free(p->lens); free(p->order);
  for (int i = 0; i < RK_BUCKETS; i++) free(p->rk[i].p);
  for (int i = 0; i < 16; i++) free(p->buckets[i].p);
  free(p);
}

/* stable insertion sort by descending length: Patterns::set_match_kind,
 * src/packed/pattern.rs:84-99 (slice::sort_by is stable) */
static void order_by_len_desc(uint32_t* order, size_t n, const size_t* lens) {
  for (size_t i = 1; i < n; i++) {
    uint32_t x = order[i];
    size_t j = i;
    while (j > 0 && lens[order[j - 1]] < lens[x]) { order[j] = order[j - 1]; j--; }
    order[j] = x;
  }
}

/* Teddy::new bucket assignment, src/packed/teddy/generic.rs:751-809 */
static void teddy_assign_buckets(orc_packed* p) {
  const int B = p->nbuckets, N = p->mask_len;
  /* map: low-nybble N-prefix -> bucket (BTreeMap in the reference; any map works) */
  typedef struct { uint8_t key[4]; int bucket; } ent;
  VEC(ent) map = {0};
  for (size_t oi = 0; oi < p->n; oi++) {
    uint32_t id = p->order[oi];
    uint8_t key[4] = {0, 0, 0, 0};
    for (int i = 0; i < N && (size_t)i < p->lens[id]; i++) key[i] = p->by_id[id][i] & 0xF;
    int bucket = -1;
    for (size_t m = 0; m < map.n; m++)
      if (memcmp(map.p[m].key, key, 4) == 0) { bucket = map.p[m].bucket; break; }
    if (bucket < 0) {
      bucket = (B - 1) - (int)(id % (uint32_t)B);
      ent e; memcpy(e.key, key, 4); e.bucket = bucket;
      VPUSH(map, e);
    }
    VPUSH(p->buckets[bucket], id);
  }
  free(map.p);
  /* Slim/FatMaskBuilder::from_teddy, generic.rs:1178-1256, 1288-1368 */
  memset(p->lo, 0, sizeof(p->lo));
  memset(p->hi, 0, sizeof(p->hi));
  for (int b = 0; b < B; b++)
    for (size_t i = 0; i < p->buckets[b].n; i++) {
      uint32_t pid = p->buckets[b].p[i];
      for (int k = 0; k < N; k++) {
        uint8_t byte = p->by_id[pid][k];
        p->lo[k][byte & 0xF] |= (uint16_t)(1u << b);
        p->hi[k][byte >> 4] |= (uint16_t)(1u << b);
      }
    }
}

/* packed::Builder::add + build, src/packed/api.rs:253-322; teddy::Builder::build_imp
 * for x86_64 with AVX2 available, src/packed/teddy/builder.rs:98-231 */
int orc_packed_build(const uint8_t* const* pats, const size_t* lens, size_t n,
                     const orc_packed_config* cfg, orc_packed** out) {
  *out = NULL;
  /* Builder::add: inert on > PATTERN_LIMIT patterns or an empty pattern */
  size_t added = 0;
  for (size_t i = 0; i < n; i++) {
    if (added >= PACKED_PATTERN_LIMIT) return ORC_OK; /* inert */
    if (lens[i] == 0) return ORC_OK;                  /* inert */
    added++;
  }
  if (added == 0) return ORC_OK;
  orc_packed* p = calloc(1, sizeof(*p));
  p->kind = cfg->kind;
  p->n = n;
  p->by_id = calloc(n, sizeof(uint8_t*));
  p->lens = calloc(n, sizeof(size_t));
  p->order = calloc(n, sizeof(uint32_t));
  p->minimum_len_pats = (size_t)-1;
  for (size_t i = 0; i < n; i++) {
    p->by_id[i] = malloc(lens[i]);
    memcpy(p->by_id[i], pats[i], lens[i]);
    p->lens[i] = lens[i];
    p->order[i] = (uint32_t)i;
    if (lens[i] < p->minimum_len_pats) p->minimum_len_pats = lens[i];
  }
  if (cfg->kind == ORC_PACKED_LEFTMOST_LONGEST) order_by_len_desc(p->order, n, p->lens);
  /* RabinKarp::new, rabinkarp.rs:41-66 */
  p->hash_len = p->minimum_len_pats;
  p->hash_2pow = 1;
  for (size_t i = 1; i < p->hash_len; i++) p->hash_2pow <<= 1;
  for (size_t oi = 0; oi < n; oi++) {
    uint32_t id = p->order[oi];
    rk_entry e = { rk_hash(p->by_id[id], p->hash_len), id };
    VPUSH(p->rk[e.hash % RK_BUCKETS], e);
  }
  if (cfg->force == ORC_FORCE_RABINKARP) {
    p->has_teddy = 0;
    p->minimum_len = 0;
    *out = p;
    return ORC_OK;
  }
  /* teddy::Builder::build_imp (x86_64, ssse3+avx2 present) */
  int patlimit = cfg->heuristic_pattern_limits;
  if (patlimit && n > 64) { orc_packed_free(p); return ORC_OK; }
  int mask_len = (int)(p->minimum_len_pats < 4 ? p->minimum_len_pats : 4);
  int beefy = n > 32;
  int use_avx2 = cfg->only_teddy_256bit == 1 ? 1 : cfg->only_teddy_256bit == 0 ? 0 : 1;
  int fat;
  if (cfg->only_teddy_fat < 0) fat = use_avx2 && beefy;
  else if (cfg->only_teddy_fat == 0) fat = 0;
  else { if (!use_avx2) { orc_packed_free(p); return ORC_OK; } fat = 1; }
  if (patlimit && mask_len == 1 && n > 16) { orc_packed_free(p); return ORC_OK; }
  p->has_teddy = 1;
  p->mask_len = mask_len;
  if (!use_avx2) { p->nbuckets = 8; p->width = 16; }
  else if (!fat) { p->nbuckets = 8; p->width = 32; }
  else { p->nbuckets = 16; p->width = 16; }
  teddy_assign_buckets(p);
  p->minimum_len = (size_t)p->width + (size_t)(mask_len - 1); /* generic.rs:94-96, 427-429 */
  *out = p;
  return ORC_OK;
}

size_t orc_packed_minimum_len(const orc_packed* p) { return p->minimum_len; }

/* Searcher::find_in, src/packed/api.rs:529-546 */
int orc_packed_find_in(const orc_packed* p, const uint8_t* hay, size_t hay_len,
                       size_t span_start, size_t span_end, orc_match* out, int* found) {
  if (span_end > hay_len || span_start > span_end) return ORC_E_INVALID_SPAN;
  if (p->has_teddy) {
    if (span_end - span_start < p->minimum_len)
      *found = rk_find_at(p, hay, span_end, span_start, out);
    else
      *found = teddy_find(p, hay, span_start, span_end, out);
  } else {
    *found = rk_find_at(p, hay, span_end, span_start, out);
  }
  return ORC_OK;
}

/* packed FindIter::next, src/packed/api.rs:661-687 */
int orc_packed_find_iter(const orc_packed* p, const uint8_t* hay, size_t hay_len,
                         orc_match* out, size_t cap, size_t* n_out) {
  size_t start = 0, end = hay_len, n = 0;
  for (;;) {
    if (start > end) break;
    orc_match m; int found = 0;
    orc_packed_find_in(p, hay, hay_len, start, end, &m, &found);
    if (!found) break;
    start = (size_t)m.end;
    if (n < cap) out[n] = m;
    n++;
  }
  *n_out = n;
  return n > cap ? ORC_E_OVERFLOW : ORC_OK;
}

/* prefilter::Builder, src/util/prefilter.rs:91-326 */
typedef struct {
  size_t count;
  int ascii_ci;
  start_bytes_t start_bytes;
  rare_bytes_t rare_bytes;
  size_t memmem_count;
  int has_packed_builder; /* MatchKind::as_packed, src/util/search.rs:1103-1114 */
  int packed_kind;
  int enabled;
  /* patterns handed to packed::Builder::add */
  VEC(const uint8_t*) ppats;
  VEC(size_t) plens;
  int packed_inert;
} pre_builder_t;

static void pre_init(pre_builder_t* b, int match_kind, int ascii_ci) {
  memset(b, 0, sizeof(*b));
  b->ascii_ci = ascii_ci;
  b->start_bytes.ascii_ci = ascii_ci;
  b->rare_bytes.ascii_ci = ascii_ci;
  b->rare_bytes.available = 1;
  b->enabled = 1;
  b->has_packed_builder = match_kind != ORC_STANDARD;
  b->packed_kind = match_kind == ORC_LEFTMOST_LONGEST ? ORC_PACKED_LEFTMOST_LONGEST
                                                      : ORC_PACKED_LEFTMOST_FIRST;
}
/* Builder::add, prefilter.rs:308-323 (+ packed::Builder::add, api.rs:294-322) */
static void pre_add(pre_builder_t* b, const uint8_t* p, size_t n) {
  if (n == 0) b->enabled = 0;
  if (!b->enabled) return;
  b->count++;
  sb_add(&b->start_bytes, p, n);
  rb_add(&b->rare_bytes, p, n);
  b->memmem_count++;
  if (b->has_packed_builder && !b->packed_inert) {
    if (b->ppats.n >= PACKED_PATTERN_LIMIT) { b->packed_inert = 1; b->ppats.n = b->plens.n = 0; }
    else { VPUSH(b->ppats, p); VPUSH(b->plens, n); }
  }
}
/* Builder::build, prefilter.rs:163-305. Returns the prefilter kind; *packed_out
 * receives the packed searcher when that is what is chosen. */
static int pre_build(pre_builder_t* b, orc_packed** packed_out) {
  *packed_out = NULL;
  if (!b->enabled) return ORC_PRE_NONE;
  if (!b->ascii_ci && b->memmem_count == 1) return ORC_PRE_MEMMEM;
  orc_packed* packed = NULL;
  size_t patlen = (size_t)-1, minlen = 0;
  if (!b->ascii_ci && b->has_packed_builder) {
    /* packed::Builder::len/minimum_len reflect the (possibly reset) pattern set */
    patlen = b->packed_inert ? 0 : b->ppats.n;
    minlen = (size_t)-1;
    for (size_t i = 0; i < b->plens.n; i++) if (b->plens.p[i] < minlen) minlen = b->plens.p[i];
    if (!b->packed_inert && b->ppats.n > 0) {
      orc_packed_config cfg; orc_packed_config_default(&cfg);
      cfg.kind = b->packed_kind;
      orc_packed_build(b->ppats.p, b->plens.p, b->ppats.n, &cfg, &packed);
    }
  }
  int has_start = sb_available(&b->start_bytes);
  int has_rare = rb_available(&b->rare_bytes);
  int kind;
  if (has_start && has_rare) {
    if (patlen <= 16 && minlen >= 2 && b->start_bytes.count >= 3 && b->rare_bytes.count >= 3) {
      kind = packed ? ORC_PRE_PACKED : ORC_PRE_NONE;
    } else {
      int fewer = b->start_bytes.count < b->rare_bytes.count;
      int rarer = b->start_bytes.rank_sum <= b->rare_bytes.rank_sum + 50;
      kind = (fewer || rarer) ? ORC_PRE_START_BYTES : ORC_PRE_RARE_BYTES;
    }
  } else if (has_start) {
    if (patlen <= 16 && minlen >= 2 && b->start_bytes.count >= 3) kind = packed ? ORC_PRE_PACKED : ORC_PRE_NONE;
    else kind = ORC_PRE_START_BYTES;
  } else if (has_rare) {
    if (patlen <= 16 && minlen >= 2 && b->rare_bytes.count >= 3) kind = packed ? ORC_PRE_PACKED : ORC_PRE_NONE;
    else kind = ORC_PRE_RARE_BYTES;
  } else if (b->ascii_ci) {
    kind = ORC_PRE_NONE;
  } else {
    kind = packed ? ORC_PRE_PACKED : ORC_PRE_NONE;
  }
  if (kind == ORC_PRE_PACKED) *packed_out = packed; else orc_packed_free(packed);
  return kind;
}
static void pre_free(pre_builder_t* b) { free(b->ppats.p); free(b->plens.p); }

/* ------------------------------------------------------------------ */
/* 4. noncontiguous NFA                                                */
/* ------------------------------------------------------------------ */

typedef struct { uint32_t sparse, dense, matches, fail, depth; } nstate;
typedef struct { uint8_t byte; uint32_t next, link; } ntrans;
typedef struct { uint32_t pid, link; } nmatch;

typedef struct {
  int match_kind;
  VEC(nstate) states;
  VEC(ntrans) sparse;
  VEC(uint32_t) dense;
  VEC(nmatch) matches;
  VEC(uint32_t) pattern_lens;
  uint8_t byte_classes[256];
  size_t min_pattern_len, max_pattern_len;
  uint32_t max_special_id, max_match_id, start_unanchored_id, start_anchored_id;
  int prefilter_kind;
  orc_packed* packed;
} nfa_t;

static int n_is_match(const nfa_t* n, uint32_t sid) { return n->states.p[sid].matches != 0; }

/* NFA::follow_transition(_sparse), src/nfa/noncontiguous.rs:339-374 */
static uint32_t n_follow(const nfa_t* n, uint32_t sid, uint8_t byte) {
  const nstate* s = &n->states.p[sid];
  if (s->dense == 0) {
    for (uint32_t link = s->sparse; link != 0; link = n->sparse.p[link].link) {
      const ntrans* t = &n->sparse.p[link];
      if (byte <= t->byte) {
        if (byte == t->byte) return t->next;
        break;
      }
    }
    return FAIL;
  }
  return n->dense.p[s->dense + n->byte_classes[byte]];
}

/* alloc_* with the StateID overflow checks of noncontiguous.rs:527-585 */
static int n_alloc_transition(nfa_t* n, uint32_t* id) {
  if (n->sparse.n > SMALL_INDEX_MAX) return ORC_E_STATE_ID_OVERFLOW;
  *id = (uint32_t)n->sparse.n;
  ntrans t = {0, 0, 0};
  VPUSH(n->sparse, t);
  return ORC_OK;
}
static int n_alloc_state(nfa_t* n, size_t depth, uint32_t* id) {
  if (n->states.n > SMALL_INDEX_MAX) return ORC_E_STATE_ID_OVERFLOW;
  *id = (uint32_t)n->states.n;
  nstate s = {0, 0, 0, n->start_unanchored_id, (uint32_t)depth};
  VPUSH(n->states, s);
  return ORC_OK;
}

/* NFA::add_transition, noncontiguous.rs:381-423 */
static int n_add_transition(nfa_t* n, uint32_t prev, uint8_t byte, uint32_t next) {
  int rc;
  if (n->states.p[prev].dense != 0)
    n->dense.p[n->states.p[prev].dense + n->byte_classes[byte]] = next;
  uint32_t head = n->states.p[prev].sparse;
  if (head == 0 || byte < n->sparse.p[head].byte) {
    uint32_t nl;
    if ((rc = n_alloc_transition(n, &nl))) return rc;
    n->sparse.p[nl].byte = byte; n->sparse.p[nl].next = next; n->sparse.p[nl].link = head;
    n->states.p[prev].sparse = nl;
    return ORC_OK;
  } else if (byte == n->sparse.p[head].byte) {
    n->sparse.p[head].next = next;
    return ORC_OK;
  }
  uint32_t link_prev = head, link_next = n->sparse.p[head].link;
  while (link_next != 0 && byte > n->sparse.p[link_next].byte) {
    link_prev = link_next;
    link_next = n->sparse.p[link_next].link;
  }
  if (link_next == 0 || byte < n->sparse.p[link_next].byte) {
    uint32_t nl;
    if ((rc = n_alloc_transition(n, &nl))) return rc;
    n->sparse.p[nl].byte = byte; n->sparse.p[nl].next = next; n->sparse.p[nl].link = link_next;
    n->sparse.p[link_prev].link = nl;
  } else {
    n->sparse.p[link_next].next = next;
  }
  return ORC_OK;
}

/* NFA::init_full_state, noncontiguous.rs:435-463 */
static int n_init_full_state(nfa_t* n, uint32_t prev, uint32_t next) {
  uint32_t prev_link = 0;
  for (unsigned b = 0; b <= 255; b++) {
    uint32_t nl; int rc;
    if ((rc = n_alloc_transition(n, &nl))) return rc;
    n->sparse.p[nl].byte = (uint8_t)b; n->sparse.p[nl].next = next; n->sparse.p[nl].link = 0;
    if (prev_link == 0) n->states.p[prev].sparse = nl; else n->sparse.p[prev_link].link = nl;
    prev_link = nl;
  }
  return ORC_OK;
}

/* NFA::add_match, noncontiguous.rs:466-484 */
static int n_add_match(nfa_t* n, uint32_t sid, uint32_t pid) {
  uint32_t link = n->states.p[sid].matches;
  while (n->matches.p[link].link != 0) link = n->matches.p[link].link;
  if (n->matches.n > SMALL_INDEX_MAX) return ORC_E_STATE_ID_OVERFLOW;
  uint32_t nl = (uint32_t)n->matches.n;
  nmatch m = {pid, 0};
  VPUSH(n->matches, m);
  if (link == 0) n->states.p[sid].matches = nl; else n->matches.p[link].link = nl;
  return ORC_OK;
}

/* NFA::copy_matches, noncontiguous.rs:490-523 */
static int n_copy_matches(nfa_t* n, uint32_t src, uint32_t dst) {
  uint32_t link_dst = n->states.p[dst].matches;
  while (n->matches.p[link_dst].link != 0) link_dst = n->matches.p[link_dst].link;
  uint32_t link_src = n->states.p[src].matches;
  while (link_src != 0) {
    if (n->matches.n > SMALL_INDEX_MAX) return ORC_E_STATE_ID_OVERFLOW;
    uint32_t nl = (uint32_t)n->matches.n;
    nmatch m = {n->matches.p[link_src].pid, 0};
    VPUSH(n->matches, m);
    if (link_dst == 0) n->states.p[dst].matches = nl; else n->matches.p[link_dst].link = nl;
    link_dst = nl;
    link_src = n->matches.p[link_src].link;
  }
  return ORC_OK;
}

/* NFA::next_state, noncontiguous.rs:601-626 */
static uint32_t n_next_state(const nfa_t* n, int anchored, uint32_t sid, uint8_t byte) {
  for (;;) {
    uint32_t next = n_follow(n, sid, byte);
    if (next != FAIL) return next;
    if (anchored) return DEAD;
    sid = n->states.p[sid].fail;
  }
}

static void nfa_free(nfa_t* n) {
  free(n->states.p); free(n->sparse.p); free(n->dense.p); free(n->matches.p);
  free(n->pattern_lens.p);
  orc_packed_free(n->packed);
}

/* Compiler::compile, noncontiguous.rs:963-1051 */
static int nfa_build(nfa_t* n, const uint8_t* const* pats, const size_t* lens, size_t npat,
                     const orc_opts* o) {
  int rc;
  memset(n, 0, sizeof(*n));
  n->match_kind = o->match_kind;
  n->min_pattern_len = (size_t)-1;
  for (int b = 0; b < 256; b++) n->byte_classes[b] = (uint8_t)b; /* ByteClasses::singletons */
  const int is_leftmost = o->match_kind != ORC_STANDARD;
  const int ci = o->ascii_case_insensitive;
  byteset_t byteset; memset(&byteset, 0, sizeof(byteset));
  pre_builder_t pre; pre_init(&pre, o->match_kind, ci);

  ntrans t0 = {0, 0, 0}; VPUSH(n->sparse, t0);
  nmatch m0 = {0, 0}; VPUSH(n->matches, m0);
  VPUSH(n->dense, DEAD);
  uint32_t id;
  if ((rc = n_alloc_state(n, 0, &id))) goto fail; /* DEAD */
  if ((rc = n_alloc_state(n, 0, &id))) goto fail; /* FAIL */
  if ((rc = n_alloc_state(n, 0, &id))) goto fail;
  n->start_unanchored_id = id;
  if ((rc = n_alloc_state(n, 0, &id))) goto fail;
  n->start_anchored_id = id;
  /* init_unanchored_start_state :1549-1555, add_dead_state_loop :1643-1646 */
  if ((rc = n_init_full_state(n, n->start_unanchored_id, FAIL))) goto fail;
  if ((rc = n_init_full_state(n, n->start_anchored_id, FAIL))) goto fail;
  if ((rc = n_init_full_state(n, DEAD, DEAD))) goto fail;

  /* build_trie :1057-1150 */
  for (size_t i = 0; i < npat; i++) {
    if (i > SMALL_INDEX_MAX) { rc = ORC_E_PATTERN_ID_OVERFLOW; goto fail; }
    if (lens[i] > SMALL_INDEX_MAX) { rc = ORC_E_PATTERN_TOO_LONG; goto fail; }
    if (lens[i] < n->min_pattern_len) n->min_pattern_len = lens[i];
    if (lens[i] > n->max_pattern_len) n->max_pattern_len = lens[i];
    VPUSH(n->pattern_lens, (uint32_t)lens[i]);
    if (o->prefilter) pre_add(&pre, pats[i], lens[i]);
    uint32_t prev = n->start_unanchored_id;
    int saw_match = 0, skip = 0;
    for (size_t depth = 0; depth < lens[i]; depth++) {
      uint8_t b = pats[i][depth];
      saw_match = saw_match || n_is_match(n, prev);
      if (o->match_kind == ORC_LEFTMOST_FIRST && saw_match) { skip = 1; break; }
      bcs_set_range(&byteset, b, b);
      if (ci) { uint8_t ob = opposite_ascii_case(b); bcs_set_range(&byteset, ob, ob); }
      uint32_t next = n_follow(n, prev, b);
      if (next != FAIL) {
        prev = next;
      } else {
        if ((rc = n_alloc_state(n, depth, &next))) goto fail;
        if ((rc = n_add_transition(n, prev, b, next))) goto fail;
        if (ci) if ((rc = n_add_transition(n, prev, opposite_ascii_case(b), next))) goto fail;
        prev = next;
      }
    }
    if (skip) continue;
    if ((rc = n_add_match(n, prev, (uint32_t)i))) goto fail;
  }
  bcs_byte_classes(&byteset, n->byte_classes);

  /* set_anchored_start_state :1561-1586 */
  {
    uint32_t ul = n->states.p[n->start_unanchored_id].sparse;
    uint32_t al = n->states.p[n->start_anchored_id].sparse;
    while (ul != 0 && al != 0) {
      n->sparse.p[al].next = n->sparse.p[ul].next;
      ul = n->sparse.p[ul].link;
      al = n->sparse.p[al].link;
    }
    if ((rc = n_copy_matches(n, n->start_unanchored_id, n->start_anchored_id))) goto fail;
    n->states.p[n->start_anchored_id].fail = DEAD;
  }
  /* add_unanchored_start_state_loop :1597-1606 */
  for (uint32_t l = n->states.p[n->start_unanchored_id].sparse; l != 0; l = n->sparse.p[l].link)
    if (n->sparse.p[l].next == FAIL) n->sparse.p[l].next = n->start_unanchored_id;

  /* densify :1500-1526 */
  {
    unsigned alen = bc_alphabet_len(n->byte_classes);
    for (size_t i = 0; i < n->states.n; i++) {
      if (i == DEAD || i == FAIL) continue;
      if (o->dense_depth >= 0 && (int64_t)n->states.p[i].depth >= o->dense_depth) continue;
      if (n->dense.n > SMALL_INDEX_MAX) { rc = ORC_E_STATE_ID_OVERFLOW; goto fail; }
      uint32_t dense = (uint32_t)n->dense.n;
      for (unsigned c = 0; c < alen; c++) VPUSH(n->dense, FAIL);
      for (uint32_t l = n->states.p[i].sparse; l != 0; l = n->sparse.p[l].link)
        n->dense.p[dense + n->byte_classes[n->sparse.p[l].byte]] = n->sparse.p[l].next;
      n->states.p[i].dense = dense;
    }
  }

  /* fill_failure_transitions :1275-1374 */
  {
    const uint32_t start_uid = n->start_unanchored_id;
    VEC(uint32_t) queue = {0};
    size_t qhead = 0;
    uint8_t* seen = ci ? calloc(n->states.n, 1) : NULL; /* QueuedSet :1657-1691 */
    for (uint32_t l = n->states.p[start_uid].sparse; l != 0; l = n->sparse.p[l].link) {
      uint32_t nx = n->sparse.p[l].next;
      if (nx == start_uid || (seen && seen[nx])) continue;
      VPUSH(queue, nx);
      if (seen) seen[nx] = 1;
      if (is_leftmost && n_is_match(n, nx)) n->states.p[nx].fail = DEAD;
    }
    while (qhead < queue.n) {
      uint32_t sid = queue.p[qhead++];
      for (uint32_t l = n->states.p[sid].sparse; l != 0; l = n->sparse.p[l].link) {
        uint32_t nx = n->sparse.p[l].next;
        uint8_t byte = n->sparse.p[l].byte;
        if (seen && seen[nx]) continue;
        VPUSH(queue, nx);
        if (seen) seen[nx] = 1;
        if (is_leftmost && n_is_match(n, nx)) { n->states.p[nx].fail = DEAD; continue; }
        uint32_t fail = n->states.p[sid].fail;
        while (n_follow(n, fail, byte) == FAIL) fail = n->states.p[fail].fail;
        fail = n_follow(n, fail, byte);
        n->states.p[nx].fail = fail;
        if ((rc = n_copy_matches(n, fail, nx))) { free(queue.p); free(seen); goto fail; }
      }
      if (!is_leftmost)
        if ((rc = n_copy_matches(n, n->start_unanchored_id, sid))) { free(queue.p); free(seen); goto fail; }
    }
    free(queue.p);
    free(seen);
  }

  /* close_start_state_loop_for_leftmost :1620-1638 */
  {
    uint32_t su = n->start_unanchored_id;
    uint32_t dense = n->states.p[su].dense;
    if (is_leftmost && n_is_match(n, su)) {
      for (uint32_t l = n->states.p[su].sparse; l != 0; l = n->sparse.p[l].link)
        if (n->sparse.p[l].next == su) {
          n->sparse.p[l].next = DEAD;
          if (dense != 0) n->dense.p[dense + n->byte_classes[n->sparse.p[l].byte]] = DEAD;
        }
    }
  }

  /* shuffle :1399-1481 with Remapper, src/util/remapper.rs:86-150 */
  {
    size_t ns = n->states.n;
    uint32_t* map = malloc(ns * sizeof(uint32_t));
    for (size_t i = 0; i < ns; i++) map[i] = (uint32_t)i;
#define RSWAP(a, b)                                                      \
  do {                                                                   \
    uint32_t a_ = (a), b_ = (b);                                         \
    if (a_ != b_) {                                                      \
      nstate ts = n->states.p[a_]; n->states.p[a_] = n->states.p[b_]; n->states.p[b_] = ts; \
      uint32_t tm = map[a_]; map[a_] = map[b_]; map[b_] = tm;            \
    }                                                                    \
  } while (0)
    uint32_t old_uid = n->start_unanchored_id, old_aid = n->start_anchored_id;
    uint32_t next_avail = 4;
    for (size_t i = 4; i < ns; i++) {
      if (!n_is_match(n, (uint32_t)i)) continue;
      RSWAP((uint32_t)i, next_avail);
      next_avail++;
    }
    uint32_t new_aid = next_avail - 1;
    RSWAP(old_aid, new_aid);
    uint32_t new_uid = next_avail - 2;
    RSWAP(old_uid, new_uid);
    n->max_match_id = next_avail - 3;
    n->start_unanchored_id = new_uid;
    n->start_anchored_id = new_aid;
    if (n_is_match(n, n->start_anchored_id)) n->max_match_id = n->start_anchored_id;
#undef RSWAP
    /* Remapper::remap :119-150 */
    uint32_t* oldmap = malloc(ns * sizeof(uint32_t));
    memcpy(oldmap, map, ns * sizeof(uint32_t));
    for (size_t i = 0; i < ns; i++) {
      uint32_t cur_id = (uint32_t)i, new_id = oldmap[i];
      if (cur_id == new_id) continue;
      for (;;) {
        uint32_t idn = oldmap[new_id];
        if (cur_id == idn) { map[i] = new_id; break; }
        new_id = idn;
      }
    }
    /* NFA::remap :261-278 */
    unsigned alen = bc_alphabet_len(n->byte_classes);
    for (size_t i = 0; i < ns; i++) {
      nstate* s = &n->states.p[i];
      s->fail = map[s->fail];
      for (uint32_t l = s->sparse; l != 0; l = n->sparse.p[l].link)
        n->sparse.p[l].next = map[n->sparse.p[l].next];
      if (s->dense != 0)
        for (unsigned c = 0; c < alen; c++) n->dense.p[s->dense + c] = map[n->dense.p[s->dense + c]];
    }
    free(map); free(oldmap);
  }

  n->prefilter_kind = pre_build(&pre, &n->packed);
  n->max_special_id = n->prefilter_kind != ORC_PRE_NONE ? n->start_anchored_id : n->max_match_id;
  pre_free(&pre);
  return ORC_OK;
fail:
  pre_free(&pre);
  nfa_free(n);
  memset(n, 0, sizeof(*n));
  return rc;
}

/* ------------------------------------------------------------------ */
/* 5. DFA                                                              */
/* ------------------------------------------------------------------ */

typedef struct {
  uint32_t* trans;
  uint64_t trans_len;
  VEC(uint32_t) * matches; /* per match state */
  size_t num_match_states;
  uint32_t* match_offsets; /* CSR view built after construction */
  uint32_t* match_pids;
  size_t state_len;
  unsigned alphabet_len, stride2;
  uint8_t byte_classes[256];
  uint32_t max_special_id, max_match_id, start_unanchored_id, start_anchored_id;
} dfa_t;

/* DFA::set_matches, src/dfa.rs:171-184 */
static void d_set_matches(dfa_t* d, uint32_t sid, const nfa_t* n, uint32_t oldsid) {
  size_t index = (sid >> d->stride2) - 2;
  for (uint32_t l = n->states.p[oldsid].matches; l != 0; l = n->matches.p[l].link)
    VPUSH(d->matches[index], n->matches.p[l].pid);
}

/* sparse_iter, src/dfa.rs:801-835 */
typedef void (*sparse_cb)(void* ctx, uint8_t byte, uint8_t cls, uint32_t next);
static void sparse_iter(const nfa_t* n, uint32_t oldsid, const uint8_t classes[256], sparse_cb f,
                        void* ctx) {
  int prev_class = -1;
  unsigned byte = 0;
  for (uint32_t l = n->states.p[oldsid].sparse; l != 0; l = n->sparse.p[l].link) {
    const ntrans* t = &n->sparse.p[l];
    while (byte < t->byte) {
      uint8_t rep = (uint8_t)byte, cls = classes[rep];
      byte++;
      if (prev_class != cls) { f(ctx, rep, cls, FAIL); prev_class = cls; }
    }
    uint8_t rep = t->byte, cls = classes[rep];
    byte++;
    if (prev_class != cls) { f(ctx, rep, cls, t->next); prev_class = cls; }
  }
  for (unsigned b = byte; b <= 255; b++) {
    uint8_t rep = (uint8_t)b, cls = classes[rep];
    if (prev_class != cls) { f(ctx, rep, cls, FAIL); prev_class = cls; }
  }
}

typedef struct {
  const nfa_t* n; dfa_t* d; int anchored; uint32_t oldsid; uint32_t newsid, anewsid; int mode;
} fb_ctx;

/* closure of finish_build_one_start, src/dfa.rs:565-591 */
static void cb_one_start(void* vctx, uint8_t byte, uint8_t cls, uint32_t oldnext) {
  fb_ctx* c = vctx;
  const nstate* st = &c->n->states.p[c->oldsid];
  if (oldnext == FAIL) {
    if (c->anchored) oldnext = DEAD;
    else if (st->fail == DEAD) oldnext = DEAD;
    else oldnext = n_next_state(c->n, 0, st->fail, byte);
  }
  c->d->trans[c->newsid + cls] = oldnext << c->d->stride2;
}
/* closures of finish_build_both_starts, src/dfa.rs:651-701 (store OLD ids; remapped later) */
static void cb_both_start(void* vctx, uint8_t byte, uint8_t cls, uint32_t oldnext) {
  (void)byte;
  fb_ctx* c = vctx;
  c->d->trans[c->newsid + cls] = (oldnext == FAIL) ? DEAD : oldnext;
}
static void cb_both_other(void* vctx, uint8_t byte, uint8_t cls, uint32_t oldnext) {
  fb_ctx* c = vctx;
  const nstate* st = &c->n->states.p[c->oldsid];
  if (oldnext == FAIL) {
    uint32_t nx = st->fail == DEAD ? DEAD : n_next_state(c->n, 0, st->fail, byte);
    c->d->trans[c->newsid + cls] = nx;
  } else {
    c->d->trans[c->newsid + cls] = oldnext;
    c->d->trans[c->anewsid + cls] = oldnext;
  }
}

static void dfa_free(dfa_t* d) {
  free(d->trans);
  if (d->matches) for (size_t i = 0; i < d->num_match_states; i++) free(d->matches[i].p);
  free(d->matches); free(d->match_offsets); free(d->match_pids);
}

/* dfa::Builder::build_from_noncontiguous, src/dfa.rs:431-540 */
static int dfa_build(dfa_t* d, const nfa_t* n, int start_kind, int byte_classes) {
  memset(d, 0, sizeof(*d));
  if (byte_classes) memcpy(d->byte_classes, n->byte_classes, 256);
  else for (int b = 0; b < 256; b++) d->byte_classes[b] = (uint8_t)b;
  d->alphabet_len = bc_alphabet_len(d->byte_classes);
  d->stride2 = bc_stride2(d->byte_classes);
  size_t ns = n->states.n;
  d->state_len = start_kind == ORC_START_BOTH ? ns * 2 - 4 : ns;
  uint64_t trans_len = (uint64_t)d->state_len << d->stride2;
  if (trans_len - ((uint64_t)1 << d->stride2) > SMALL_INDEX_MAX) return ORC_E_STATE_ID_OVERFLOW;
  d->trans_len = trans_len;
  d->num_match_states = start_kind == ORC_START_BOTH ? ((size_t)n->max_match_id - 1) * 2
                                                      : (size_t)n->max_match_id - 1;
  d->trans = calloc(trans_len ? trans_len : 1, sizeof(uint32_t));
  d->matches = calloc(d->num_match_states ? d->num_match_states : 1, sizeof(*d->matches));
  const unsigned s2 = d->stride2;
  if (start_kind != ORC_START_BOTH) {
    /* finish_build_one_start :544-607 */
    int anchored = start_kind == ORC_START_ANCHORED;
    for (size_t old = 0; old < ns; old++) {
      uint32_t newsid = (uint32_t)old << s2;
      if (n_is_match(n, (uint32_t)old)) d_set_matches(d, newsid, n, (uint32_t)old);
      fb_ctx c = {n, d, anchored, (uint32_t)old, newsid, 0, 0};
      sparse_iter(n, (uint32_t)old, d->byte_classes, cb_one_start, &c);
    }
    d->max_special_id = n->max_special_id << s2;
    d->max_match_id = n->max_match_id << s2;
    if (anchored) { d->start_unanchored_id = DEAD; d->start_anchored_id = n->start_anchored_id << s2; }
    else { d->start_unanchored_id = n->start_unanchored_id << s2; d->start_anchored_id = DEAD; }
  } else {
    /* finish_build_both_starts :617-724 */
    const uint32_t stride = 1u << s2;
    uint32_t* remap_u = calloc(ns, sizeof(uint32_t));
    uint32_t* remap_a = calloc(ns, sizeof(uint32_t));
    uint8_t* is_anch = calloc(d->state_len, 1);
    uint32_t newsid = DEAD;
    for (size_t old = 0; old < ns; old++) {
      if (old == DEAD || old == FAIL) {
        remap_u[old] = newsid; remap_a[old] = newsid; newsid += stride;
      } else if (old == n->start_unanchored_id || old == n->start_anchored_id) {
        if (old == n->start_unanchored_id) { remap_u[old] = newsid; remap_a[old] = DEAD; }
        else { remap_u[old] = DEAD; remap_a[old] = newsid; is_anch[newsid >> s2] = 1; }
        if (n_is_match(n, (uint32_t)old)) d_set_matches(d, newsid, n, (uint32_t)old);
        fb_ctx c = {n, d, 0, (uint32_t)old, newsid, 0, 0};
        sparse_iter(n, (uint32_t)old, d->byte_classes, cb_both_start, &c);
        newsid += stride;
      } else {
        uint32_t un = newsid; newsid += stride;
        uint32_t an = newsid; newsid += stride;
        remap_u[old] = un; remap_a[old] = an; is_anch[an >> s2] = 1;
        if (n_is_match(n, (uint32_t)old)) {
          d_set_matches(d, un, n, (uint32_t)old);
          d_set_matches(d, an, n, (uint32_t)old);
        }
        fb_ctx c = {n, d, 0, (uint32_t)old, un, an, 0};
        sparse_iter(n, (uint32_t)old, d->byte_classes, cb_both_other, &c);
      }
    }
    for (size_t i = 0; i < d->state_len; i++) {
      uint32_t* row = d->trans + (i << s2);
      const uint32_t* rm = is_anch[i] ? remap_a : remap_u;
      for (uint32_t c = 0; c < stride; c++) row[c] = rm[row[c]];
    }
    d->max_special_id = remap_a[n->max_special_id];
    d->max_match_id = remap_a[n->max_match_id];
    d->start_unanchored_id = remap_u[n->start_unanchored_id];
    d->start_anchored_id = remap_a[n->start_anchored_id];
    free(remap_u); free(remap_a); free(is_anch);
  }
  /* CSR view of `matches: Vec<Vec<PatternID>>` */
  size_t total = 0;
  for (size_t i = 0; i < d->num_match_states; i++) total += d->matches[i].n;
  d->match_offsets = calloc(d->num_match_states + 1, sizeof(uint32_t));
  d->match_pids = calloc(total ? total : 1, sizeof(uint32_t));
  size_t off = 0;
  for (size_t i = 0; i < d->num_match_states; i++) {
    d->match_offsets[i] = (uint32_t)off;
    memcpy(d->match_pids + off, d->matches[i].p, d->matches[i].n * sizeof(uint32_t));
    off += d->matches[i].n;
  }
  d->match_offsets[d->num_match_states] = (uint32_t)off;
  return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* 6/7. facade + search loops                                          */
/* ------------------------------------------------------------------ */

struct orc_ac {
  int kind;       /* what the reference would have built (NFA / CONTIGUOUS / DFA) */
  int use_dfa;
  int start_kind;
  nfa_t nfa;
  dfa_t dfa;
};

void orc_opts_default(orc_opts* o) {
  o->match_kind = ORC_STANDARD;
  o->start_kind = ORC_START_UNANCHORED;
  o->ascii_case_insensitive = 0;
  o->byte_classes = 1;
  o->prefilter = 1;
  o->kind = ORC_KIND_AUTO;
  o->dense_depth = 3;
}

int orc_build(const uint8_t* const* pats, const size_t* lens, size_t n, const orc_opts* opts,
              orc_ac** out) {
  *out = NULL;
  orc_ac* ac = calloc(1, sizeof(*ac));
  ac->start_kind = opts->start_kind;
  int rc = nfa_build(&ac->nfa, pats, lens, n, opts);
  if (rc) { free(ac); return rc; }
  int kind = opts->kind;
  if (kind == ORC_KIND_AUTO) {
    /* build_auto, src/ahocorasick.rs:2213-2261 */
    int try_dfa = opts->start_kind != ORC_START_BOTH && n <= 100;
    if (try_dfa && dfa_build(&ac->dfa, &ac->nfa, opts->start_kind, opts->byte_classes) == ORC_OK) {
      ac->kind = ORC_KIND_DFA; ac->use_dfa = 1;
    } else {
      dfa_free(&ac->dfa); memset(&ac->dfa, 0, sizeof(ac->dfa));
      ac->kind = ORC_KIND_CONTIGUOUS; ac->use_dfa = 0;
    }
  } else if (kind == ORC_KIND_DFA) {
    rc = dfa_build(&ac->dfa, &ac->nfa, opts->start_kind, opts->byte_classes);
    if (rc) { dfa_free(&ac->dfa); nfa_free(&ac->nfa); free(ac); return rc; }
    ac->kind = ORC_KIND_DFA; ac->use_dfa = 1;
  } else {
    ac->kind = kind; ac->use_dfa = 0;
  }
  *out = ac;
  return ORC_OK;
}

void orc_free(orc_ac* ac) {
  if (!ac) return;
  dfa_free(&ac->dfa);
  nfa_free(&ac->nfa);
  free(ac);
}

int orc_kind(const orc_ac* ac) { return ac->kind; }
int orc_match_kind(const orc_ac* ac) { return ac->nfa.match_kind; }
int orc_start_kind(const orc_ac* ac) { return ac->start_kind; }
size_t orc_patterns_len(const orc_ac* ac) { return ac->nfa.pattern_lens.n; }
size_t orc_min_pattern_len(const orc_ac* ac) { return ac->nfa.min_pattern_len; }
size_t orc_max_pattern_len(const orc_ac* ac) { return ac->nfa.max_pattern_len; }
int orc_prefilter_kind(const orc_ac* ac) { return ac->nfa.prefilter_kind; }
int orc_packed_variant(const orc_ac* ac, int* fat, int* mask_len, int* vbytes) {
  const orc_packed* p = ac->nfa.packed;
  if (!p || !p->has_teddy) return 0;
  *fat = p->nbuckets == 16; *mask_len = p->mask_len; *vbytes = p->nbuckets == 16 ? 32 : p->width;
  return 1;
}

int orc_dfa_get(const orc_ac* ac, orc_dfa_view* v) {
  if (!ac->use_dfa) return ORC_E_UNSUPPORTED_KIND;
  const dfa_t* d = &ac->dfa;
  v->trans = d->trans; v->trans_len = d->trans_len; v->stride2 = d->stride2;
  v->alphabet_len = d->alphabet_len; v->byte_classes = d->byte_classes;
  v->max_special_id = d->max_special_id; v->max_match_id = d->max_match_id;
  v->start_unanchored_id = d->start_unanchored_id; v->start_anchored_id = d->start_anchored_id;
  v->match_offsets = d->match_offsets; v->match_pids = d->match_pids;
  v->num_match_states = (uint32_t)d->num_match_states;
  v->pattern_lens = ac->nfa.pattern_lens.p; v->n_patterns = (uint32_t)ac->nfa.pattern_lens.n;
  v->match_kind = (uint32_t)ac->nfa.match_kind;
  v->min_pattern_len = ac->nfa.min_pattern_len; v->max_pattern_len = ac->nfa.max_pattern_len;
  v->state_len = d->state_len;
  return ORC_OK;
}

/* --- Automaton trait surface (src/automaton.rs:198-637) over NFA or DFA --- */
static int a_start_state(const orc_ac* ac, int anchored, uint32_t* sid) {
  if (ac->use_dfa) { /* src/dfa.rs:192-215 */
    uint32_t s = anchored ? ac->dfa.start_anchored_id : ac->dfa.start_unanchored_id;
    if (s == DEAD) return anchored ? ORC_E_INVALID_INPUT_ANCHORED : ORC_E_INVALID_INPUT_UNANCHORED;
    *sid = s;
  } else { /* src/nfa/noncontiguous.rs:593-598 */
    *sid = anchored ? ac->nfa.start_anchored_id : ac->nfa.start_unanchored_id;
  }
  return ORC_OK;
}
static inline uint32_t a_next_state(const orc_ac* ac, int anchored, uint32_t sid, uint8_t byte) {
  if (ac->use_dfa) return ac->dfa.trans[sid + ac->dfa.byte_classes[byte]]; /* src/dfa.rs:218-226 */
  return n_next_state(&ac->nfa, anchored, sid, byte);
}
static inline int a_is_special(const orc_ac* ac, uint32_t sid) {
  return sid <= (ac->use_dfa ? ac->dfa.max_special_id : ac->nfa.max_special_id);
}
static inline int a_is_dead(uint32_t sid) { return sid == DEAD; }
static inline int a_is_match(const orc_ac* ac, uint32_t sid) {
  return sid != DEAD && sid <= (ac->use_dfa ? ac->dfa.max_match_id : ac->nfa.max_match_id);
}
static size_t a_match_len(const orc_ac* ac, uint32_t sid) {
  if (ac->use_dfa) { /* src/dfa.rs:275-279 */
    size_t off = (sid >> ac->dfa.stride2) - 2;
    return ac->dfa.match_offsets[off + 1] - ac->dfa.match_offsets[off];
  }
  size_t c = 0; /* src/nfa/noncontiguous.rs:679-681 */
  for (uint32_t l = ac->nfa.states.p[sid].matches; l != 0; l = ac->nfa.matches.p[l].link) c++;
  return c;
}
static uint32_t a_match_pattern(const orc_ac* ac, uint32_t sid, size_t index) {
  if (ac->use_dfa) { /* src/dfa.rs:282-286 */
    size_t off = (sid >> ac->dfa.stride2) - 2;
    return ac->dfa.match_pids[ac->dfa.match_offsets[off] + index];
  }
  uint32_t l = ac->nfa.states.p[sid].matches; /* src/nfa/noncontiguous.rs:684-686 */
  while (index--) l = ac->nfa.matches.p[l].link;
  return ac->nfa.matches.p[l].pid;
}
/* get_match, src/automaton.rs:1540-1549 */
static orc_match a_get_match(const orc_ac* ac, uint32_t sid, size_t index, size_t at) {
  uint32_t pid = a_match_pattern(ac, sid, index);
  size_t len = ac->nfa.pattern_lens.p[pid];
  orc_match m = {pid, 0, at - len, at};
  return m;
}

/* enforce_anchored_consistency, src/ahocorasick.rs:2778-2789 */
static int enforce_anchored(int have, int want_anchored) {
  if (have == ORC_START_BOTH) return ORC_OK;
  if (have == ORC_START_UNANCHORED) return want_anchored ? ORC_E_INVALID_INPUT_ANCHORED : ORC_OK;
  return want_anchored ? ORC_OK : ORC_E_INVALID_INPUT_UNANCHORED;
}

/* Input::set_span validity, src/util/search.rs:332-343 (the reference panics) */
static int span_ok(size_t hay_len, size_t s, size_t e) { return e <= hay_len && s <= e + 1; }

/* try_find_fwd + try_find_fwd_imp, src/automaton.rs:1259-1420.  Prefilters that
 * only skip ahead (memmem/start/rare bytes) never change results and are not
 * modelled; the packed prefilter returns confirmed matches (:1301-1304) and is. */
static int a_try_find(const orc_ac* ac, const uint8_t* hay, size_t hay_len, size_t start,
                      size_t end, int anchored, int earliest_in, orc_match* out, int* found) {
  (void)hay_len;
  *found = 0;
  if (start > end) return ORC_OK; /* Input::is_done */
  int earliest = ac->nfa.match_kind == ORC_STANDARD || earliest_in;
  uint32_t sid;
  int rc = a_start_state(ac, anchored, &sid);
  if (rc) return rc;
  size_t at = start;
  orc_match mat; int have = 0;
  if (a_is_match(ac, sid)) {
    mat = a_get_match(ac, sid, 0, at); have = 1;
    if (earliest) { *out = mat; *found = 1; return ORC_OK; }
  }
  if (!anchored && ac->nfa.prefilter_kind == ORC_PRE_PACKED) {
    return orc_packed_find_in(ac->nfa.packed, hay, end, start, end, out, found);
  }
  while (at < end) {
    sid = a_next_state(ac, anchored, sid, hay[at]);
    if (a_is_special(ac, sid)) {
      if (a_is_dead(sid)) break;
      if (a_is_match(ac, sid)) {
        orc_match m = a_get_match(ac, sid, 0, at + 1);
        if (!(anchored && m.start > start)) {
          mat = m; have = 1;
          if (earliest) break;
        }
      }
    }
    at++;
  }
  if (have) { *out = mat; *found = 1; }
  return ORC_OK;
}

int orc_try_find(const orc_ac* ac, const uint8_t* hay, size_t hay_len, size_t span_start,
                 size_t span_end, int anchored, int earliest, orc_match* out, int* found) {
  if (!span_ok(hay_len, span_start, span_end)) return ORC_E_INVALID_SPAN;
  int rc = enforce_anchored(ac->start_kind, anchored);
  if (rc) return rc;
  return a_try_find(ac, hay, hay_len, span_start, span_end, anchored, earliest, out, found);
}

/* FindIter::{new,next,handle_overlapping_empty_match}, src/automaton.rs:857-936 */
int orc_find_iter(const orc_ac* ac, const uint8_t* hay, size_t hay_len, size_t span_start,
                  size_t span_end, int anchored, orc_match* out, size_t cap, size_t* n_out) {
  *n_out = 0;
  if (!span_ok(hay_len, span_start, span_end)) return ORC_E_INVALID_SPAN;
  int rc = enforce_anchored(ac->start_kind, anchored);
  if (rc) return rc;
  uint32_t sid;
  if ((rc = a_start_state(ac, anchored, &sid))) return rc;
  size_t start = span_start, n = 0;
  int have_last = 0; size_t last_match_end = 0;
  for (;;) {
    orc_match m; int found;
    a_try_find(ac, hay, hay_len, start, span_end, anchored, 0, &m, &found);
    if (!found) break;
    if (m.start == m.end) {
      if (have_last && m.end == last_match_end) {
        start = start + 1;
        a_try_find(ac, hay, hay_len, start, span_end, anchored, 0, &m, &found);
        if (!found) break;
      }
    }
    start = (size_t)m.end;
    last_match_end = (size_t)m.end; have_last = 1;
    if (n < cap) out[n] = m;
    n++;
  }
  *n_out = n;
  return n > cap ? ORC_E_OVERFLOW : ORC_OK;
}

/* OverlappingState, src/automaton.rs:782-827 */
typedef struct {
  int has_mat; orc_match mat;
  int has_id; uint32_t id;
  size_t at;
  int has_nmi; size_t nmi;
} ov_state;

/* try_find_overlapping_fwd(_imp), src/automaton.rs:1423-1537 */
static int a_try_find_overlapping(const orc_ac* ac, const uint8_t* hay, size_t start, size_t end,
                                  int anchored, ov_state* st) {
  st->has_mat = 0;
  if (start > end) return ORC_OK;
  uint32_t sid;
  if (!st->has_id) {
    int rc = a_start_state(ac, anchored, &sid);
    if (rc) return rc;
    if (a_is_match(ac, sid)) {
      size_t i = st->has_nmi ? st->nmi : 0;
      if (i < a_match_len(ac, sid)) {
        st->has_nmi = 1; st->nmi = i + 1;
        st->mat = a_get_match(ac, sid, i, start); st->has_mat = 1;
        return ORC_OK;
      }
    }
    st->at = start; st->has_id = 1; st->id = sid; st->has_nmi = 0; st->has_mat = 0;
  } else {
    sid = st->id;
    if (st->has_nmi) {
      size_t i = st->nmi;
      if (i < a_match_len(ac, sid)) {
        st->nmi = i + 1;
        st->mat = a_get_match(ac, sid, i, st->at + 1); st->has_mat = 1;
        return ORC_OK;
      }
      st->at += 1; st->has_nmi = 0; st->has_mat = 0;
    }
  }
  while (st->at < end) {
    sid = a_next_state(ac, anchored, sid, hay[st->at]);
    if (a_is_special(ac, sid)) {
      st->id = sid; st->has_id = 1;
      if (a_is_dead(sid)) return ORC_OK;
      if (a_is_match(ac, sid)) {
        st->has_nmi = 1; st->nmi = 1;
        st->mat = a_get_match(ac, sid, 0, st->at + 1); st->has_mat = 1;
        return ORC_OK;
      }
    }
    st->at += 1;
  }
  st->id = sid; st->has_id = 1;
  return ORC_OK;
}

/* Automaton::try_find_overlapping_iter + FindOverlappingIter::next,
 * src/automaton.rs:397-423, 954-970 */
int orc_find_overlapping_iter(const orc_ac* ac, const uint8_t* hay, size_t hay_len,
                              size_t span_start, size_t span_end, int anchored, orc_match* out,
                              size_t cap, size_t* n_out) {
  *n_out = 0;
  if (!span_ok(hay_len, span_start, span_end)) return ORC_E_INVALID_SPAN;
  int rc = enforce_anchored(ac->start_kind, anchored);
  if (rc) return rc;
  if (ac->nfa.match_kind != ORC_STANDARD) return ORC_E_UNSUPPORTED_OVERLAPPING;
  if (anchored) return ORC_E_INVALID_INPUT_ANCHORED;
  uint32_t sid;
  if ((rc = a_start_state(ac, anchored, &sid))) return rc;
  ov_state st; memset(&st, 0, sizeof(st));
  size_t n = 0;
  for (;;) {
    a_try_find_overlapping(ac, hay, span_start, span_end, anchored, &st);
    if (!st.has_mat) break;
    if (n < cap) out[n] = st.mat;
    n++;
  }
  *n_out = n;
  return n > cap ? ORC_E_OVERFLOW : ORC_OK;
}

/* CPU baseline: the overlapping hot loop (src/automaton.rs:1491-1534 with
 * src/dfa.rs:218-226 inlined) over a DFA, counting matches. */
int orc_scan_overlapping_count(const orc_ac* ac, const uint8_t* hay, size_t hay_len,
                               size_t span_start, size_t span_end, uint64_t* n_matches,
                               uint64_t* fnv) {
  if (!ac->use_dfa) return ORC_E_UNSUPPORTED_KIND;
  if (!span_ok(hay_len, span_start, span_end)) return ORC_E_INVALID_SPAN;
  const dfa_t* d = &ac->dfa;
  const uint32_t* trans = d->trans;
  const uint8_t* cls = d->byte_classes;
  const uint32_t max_match = d->max_match_id;
  const uint32_t* plen = ac->nfa.pattern_lens.p;
  uint32_t sid = d->start_unanchored_id;
  uint64_t cnt = 0, h = 0xcbf29ce484222325ull;
#define FNV_MIX(x) do { uint64_t v_ = (x); for (int k_ = 0; k_ < 8; k_++) { h ^= (v_ >> (8 * k_)) & 0xFF; h *= 0x100000001b3ull; } } while (0)
  if (sid != DEAD && sid <= max_match) {
    size_t off = (sid >> d->stride2) - 2;
    for (uint32_t i = d->match_offsets[off]; i < d->match_offsets[off + 1]; i++) {
      uint32_t pid = d->match_pids[i];
      FNV_MIX(pid); FNV_MIX(span_start - plen[pid]); FNV_MIX(span_start); cnt++;
    }
  }
  for (size_t at = span_start; at < span_end; at++) {
    sid = trans[sid + cls[hay[at]]];
    if (sid <= max_match) {
      if (sid == DEAD) break;
      size_t off = (sid >> d->stride2) - 2;
      for (uint32_t i = d->match_offsets[off]; i < d->match_offsets[off + 1]; i++) {
        uint32_t pid = d->match_pids[i];
        FNV_MIX(pid); FNV_MIX(at + 1 - plen[pid]); FNV_MIX(at + 1); cnt++;
      }
    }
  }
#undef FNV_MIX
  *n_matches = cnt; *fnv = h;
  return ORC_OK;
}
