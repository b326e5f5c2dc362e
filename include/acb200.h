/*
 * acb200.h -- C ABI of the B200-native Aho-Corasick search path (libacb200.so).
 *
 * This is the drop-in boundary.  The reference has no
 * FFI in-tree; the seam its hot path sits behind is the sealed trait
 * `unsafe trait Automaton` (src/automaton.rs:198) reached from `AhoCorasick`
 * through `Arc<dyn AcAutomaton>` (src/ahocorasick.rs:177-180) with exactly two
 * virtual entry points on the search path -- `try_find` and
 * `try_find_overlapping` (src/ahocorasick.rs:2757-2772).  Every export below
 * cites the reference interface it replaces.  Plain pointers and sizes only,
 * no exceptions/panics cross the boundary, handles are immutable after
 * creation and safe for concurrent searches (the reference's automata are
 * Send + Sync, src/lib.rs:274-326).
 *
 * There is no CPU fallback: every search entry point returns ACG_E_CUDA /
 * ACG_E_NO_DEVICE if no CUDA device is usable.
 */
#ifndef ACB200_H
#define ACB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/util/search.rs:1052 MatchKind */
enum { ACG_STANDARD = 0, ACG_LEFTMOST_FIRST = 1, ACG_LEFTMOST_LONGEST = 2 };
/* src/util/search.rs:1133 StartKind */
enum { ACG_START_UNANCHORED = 0, ACG_START_ANCHORED = 1, ACG_START_BOTH = 2 };
/* src/ahocorasick.rs:2624 AhoCorasickKind (0 = None / auto, :2213-2261) */
enum { ACG_KIND_AUTO = 0, ACG_KIND_NONCONTIGUOUS_NFA = 1, ACG_KIND_CONTIGUOUS_NFA = 2, ACG_KIND_DFA = 3 };
/* which prefilter the reference would have built, src/util/prefilter.rs:163-305 */
enum { ACG_PRE_NONE = 0, ACG_PRE_MEMMEM = 1, ACG_PRE_START_BYTES = 2, ACG_PRE_RARE_BYTES = 3, ACG_PRE_PACKED = 4 };
/* device engines (acg_set_engine / acg_last_engine) */
enum {
  ACG_ENGINE_AUTO = 0,
  ACG_ENGINE_WALK = 1,      /* sharded DFA state-transition scan (K1) */
  ACG_ENGINE_PREFILTER = 2, /* k-gram prefilter + anchored DFA verify (K3/K3b), the packed/Teddy role */
  ACG_ENGINE_SEQUENTIAL = 3 /* single-lane restatement of the reference loop (anchored inputs, empty patterns) */
};

/* Error convention: 0 = OK; negative codes map 1:1 to BuildError
 * (src/util/error.rs:23-49) and MatchErrorKind (:200-223), plus boundary codes. */
enum {
  ACG_OK = 0,
  ACG_E_STATE_ID_OVERFLOW = -1,
  ACG_E_PATTERN_ID_OVERFLOW = -2,
  ACG_E_PATTERN_TOO_LONG = -3,
  ACG_E_INVALID_INPUT_ANCHORED = -10,
  ACG_E_INVALID_INPUT_UNANCHORED = -11,
  ACG_E_UNSUPPORTED_STREAM = -12,
  ACG_E_UNSUPPORTED_OVERLAPPING = -13,
  ACG_E_UNSUPPORTED_EMPTY = -14,
  ACG_E_INVALID_SPAN = -20, /* the reference panics, src/util/search.rs:332-343 */
  ACG_E_OVERFLOW = -21,     /* out buffer too small; *n_out holds the required count */
  ACG_E_INVALID_ARG = -22,
  ACG_E_CUDA = -30,
  ACG_E_NO_DEVICE = -31,
  ACG_E_NOMEM = -32
};

/* `Match { pattern: PatternID, span: Span }`, src/util/search.rs:825-830 */
typedef struct {
  uint32_t pid;
  uint32_t _pad;
  uint64_t start;
  uint64_t end;
} acg_match;

/* AhoCorasickBuilder knobs, src/ahocorasick.rs:2135-2617 */
typedef struct {
  int32_t match_kind;             /* default ACG_STANDARD */
  int32_t start_kind;             /* default ACG_START_UNANCHORED */
  int32_t ascii_case_insensitive; /* default 0 */
  int32_t byte_classes;           /* default 1 */
  int32_t prefilter;              /* default 1 */
  int32_t kind;                   /* default ACG_KIND_AUTO; the device always executes a DFA */
  int64_t dense_depth;            /* accepted for API parity; has no effect on a DFA */
} acg_build_opts;

/* The data `DFA` exposes through the Automaton trait (src/dfa.rs:91-132,
 * 192-302).  This is what a Rust `-sys` shim passes after building the
 * automaton with the reference's own builder. */
typedef struct {
  const uint32_t* trans;        /* premultiplied next-state ids, row-major [state_len][1<<stride2] */
  uint64_t trans_len;
  uint32_t stride2;
  uint32_t alphabet_len;
  uint8_t byte_classes[256];
  uint32_t max_special_id, max_match_id, start_unanchored_id, start_anchored_id;
  const uint32_t* match_offsets; /* CSR over match states 2..=max_match_id>>stride2: [n+1] */
  const uint32_t* match_pids;
  const uint32_t* pattern_lens;
  uint32_t n_patterns;
  uint32_t match_kind;
  uint32_t start_kind;
  uint32_t prefilter_kind;       /* informative (ACG_PRE_*) */
  uint64_t min_pattern_len, max_pattern_len;
} acg_dfa_desc;

typedef struct acg_dfa acg_dfa;

void acg_build_opts_default(acg_build_opts* o);

/* AhoCorasickBuilder::build (src/ahocorasick.rs:2171-2207) with kind=DFA:
 * noncontiguous construction (src/nfa/noncontiguous.rs:963-1051) followed by
 * dfa::Builder::build_from_noncontiguous (src/dfa.rs:431-540), then upload.
 * The host tables are bit-identical to the reference's. */
int acg_build(const uint8_t* const* patterns, const uint64_t* lens, uint64_t n,
              const acg_build_opts* opts, acg_dfa** out);
/* Same, host tables only (no CUDA needed): for table-parity checks. Searches
 * on such a handle return ACG_E_NO_DEVICE. */
int acg_build_host(const uint8_t* const* patterns, const uint64_t* lens, uint64_t n,
                   const acg_build_opts* opts, acg_dfa** out);
/* acg_build with the dense transition table produced on the GPU: the host
 * runs the noncontiguous construction (trie, failure links, match lists, state permutation:
 * src/nfa/noncontiguous.rs:963-1481) and ships that compact form; the cells of
 * dfa::Builder::finish_build_one_start (src/dfa.rs:544-593) -- next_state for every (state, class)
 * -- are filled by a kernel, one launch per trie level, each row inheriting the finished row of
 * its failure state.  Same table bit for bit (acg_dfa_table fetches it back on demand); nothing of
 * size state_len x stride is built on or copied from the host.  Applies to StartKind::Unanchored
 * (the other start kinds take the acg_build path). */
int acg_build_on_device(const uint8_t* const* patterns, const uint64_t* lens, uint64_t n,
                        const acg_build_opts* opts, acg_dfa** out);
/* Adopt a DFA built by the reference itself (pointers borrowed for the call). */
int acg_dfa_create(const acg_dfa_desc* desc, acg_dfa** out);
void acg_dfa_free(acg_dfa* dfa);

/* Host view of the tables held by the handle (borrowed until acg_dfa_free). */
int acg_dfa_table(const acg_dfa* dfa, acg_dfa_desc* out);
uint64_t acg_dfa_state_len(const acg_dfa* dfa);
/* getters, src/ahocorasick.rs:1867-2021 */
int acg_kind(const acg_dfa* dfa); /* what AhoCorasick::kind() would report for these options */
int acg_match_kind(const acg_dfa* dfa);
int acg_start_kind(const acg_dfa* dfa);
uint64_t acg_patterns_len(const acg_dfa* dfa);
uint64_t acg_min_pattern_len(const acg_dfa* dfa);
uint64_t acg_max_pattern_len(const acg_dfa* dfa);
uint64_t acg_memory_usage(const acg_dfa* dfa);
int acg_prefilter_kind(const acg_dfa* dfa);
/* Teddy variant the reference would pick (src/packed/teddy/builder.rs:98-231); 0 if not packed */
int acg_packed_variant(const acg_dfa* dfa, int* fat, int* mask_len);

/* Engine override (default AUTO) and introspection for tests/bench. */
int acg_set_engine(acg_dfa* dfa, int engine);
int acg_last_engine(const acg_dfa* dfa);

/* ---- searches over HOST buffers (copies are part of the call) ------------- */

/* AhoCorasick::try_find_overlapping_iter(...).collect()
 * (src/ahocorasick.rs:1350 -> src/automaton.rs:397-423, 954-970, 1423-1537):
 * all matches in the reference's iteration order. Two-call protocol on
 * ACG_E_OVERFLOW. */
int acg_find_overlapping(const acg_dfa* dfa, const uint8_t* hay, uint64_t hay_len,
                         uint64_t span_start, uint64_t span_end, int anchored,
                         acg_match* out, uint64_t cap, uint64_t* n_out);
/* AhoCorasick::try_find_iter(...).collect()
 * (src/ahocorasick.rs:1275 -> src/automaton.rs:844-936, 1259-1420). */
int acg_find_iter(const acg_dfa* dfa, const uint8_t* hay, uint64_t hay_len,
                  uint64_t span_start, uint64_t span_end, int anchored,
                  acg_match* out, uint64_t cap, uint64_t* n_out);
/* AhoCorasick::try_find (src/ahocorasick.rs:1021) / is_match (:311, earliest=1). */
int acg_find(const acg_dfa* dfa, const uint8_t* hay, uint64_t hay_len,
             uint64_t span_start, uint64_t span_end, int anchored, int earliest,
             acg_match* out, int* found);

/* ---- searches over DEVICE-resident haystacks (roofline measurement; no H2D) -
 * d_hay points at haystack byte 0 in device memory.  Results are written to the
 * host array `out` (matches are sparse); *kernel_ms, if non-NULL, receives the
 * CUDA-event time of the scan kernels on the library's stream. */
int acg_find_overlapping_dev(const acg_dfa* dfa, const void* d_hay, uint64_t hay_len,
                             uint64_t span_start, uint64_t span_end,
                             acg_match* out, uint64_t cap, uint64_t* n_out, float* kernel_ms);
int acg_find_iter_dev(const acg_dfa* dfa, const void* d_hay, uint64_t hay_len,
                      uint64_t span_start, uint64_t span_end,
                      acg_match* out, uint64_t cap, uint64_t* n_out, float* kernel_ms);
/* Same as acg_find_overlapping_dev but the ordered matches stay on the device: d_out is a device
 * array of acg_match (cap entries); only matches with end > min_end are kept (shard ownership
 * by end offset) and `offset_add` is added to start/end (global offsets of
 * a sliced haystack).  *n_out receives the number written.  Feeds the NCCL gather directly. */
int acg_find_overlapping_devout(const acg_dfa* dfa, const void* d_hay, uint64_t hay_len,
                                uint64_t span_start, uint64_t span_end, uint64_t min_end,
                                uint64_t offset_add, void* d_out, uint64_t cap, uint64_t* n_out,
                                float* kernel_ms);
/* Count-only variants: scan + order on the device, return the number of matches
 * and an FNV-1a checksum of the ordered (pid,start,end) stream computed on the
 * device-ordered tuples (host side folds it).  `d_out`/cap may be 0/NULL. */
int acg_count_overlapping_dev(const acg_dfa* dfa, const void* d_hay, uint64_t hay_len,
                              uint64_t span_start, uint64_t span_end,
                              uint64_t* n_out, uint64_t* fnv, float* kernel_ms);

/* ---- multi-GPU: haystack slices + gather of match buffers to rank 0 ----
 * One process (or thread) per GPU.  The path shards naturally: rank g owns the matches whose END
 * lies in (own_lo, own_hi] (rank 0 also owns end == span_start: empty-pattern matches of the start
 * state, src/automaton.rs:1456-1464), scans [read_lo, own_hi) from a cold start with
 * read_lo = own_lo - (max_pattern_len - 1), and never exchanges haystack bytes.  The only
 * exchange is the gather of the per-rank match buffers to rank 0; concatenated in rank order they
 * are the list AhoCorasick::find_overlapping_iter yields on the whole haystack
 * (src/automaton.rs:954-970, 1423-1537).
 * Transport: the records are stored by each rank's expand kernel directly into rank 0's receive
 * buffer through a cudaIpc peer mapping (NVLink / NVSwitch); NCCL carries the 8-byte counts and the
 * closing barrier, and the payload too (ncclSend / ncclRecv) if the peer mapping is unavailable. */
enum { ACG_TRANSPORT_NONE = 0, ACG_TRANSPORT_PEER = 1, ACG_TRANSPORT_NCCL = 2 };
#define ACG_COMM_ID_BYTES 128
typedef struct acg_comm acg_comm;
/* Rank 0 creates the rendezvous token (an ncclUniqueId) and hands it to the other ranks by whatever
 * channel launched them (MPI, torch.distributed, a file). */
int acg_comm_unique_id(uint8_t id[ACG_COMM_ID_BYTES]);
/* Collective over all ranks; binds to the calling thread's current CUDA device. */
int acg_comm_init(const uint8_t id[ACG_COMM_ID_BYTES], int rank, int nranks, acg_comm** out);
void acg_comm_free(acg_comm* comm);
int acg_comm_rank(const acg_comm* comm);
int acg_comm_size(const acg_comm* comm);
int acg_comm_transport(const acg_comm* comm); /* ACG_TRANSPORT_* */
/* Slice of rank `rank` out of `nranks` for the span [span_start, span_end): the rank owns ends in
 * (own_lo, own_hi] and must hold the haystack bytes [read_lo, own_hi).  Pure arithmetic. */
int acg_shard_plan(uint64_t span_start, uint64_t span_end, int nranks, int rank, uint64_t max_pattern_len,
                   uint64_t* own_lo, uint64_t* own_hi, uint64_t* read_lo);
typedef struct {
  uint64_t local_matches;  /* records this rank contributed */
  uint64_t total_matches;  /* records in rank 0's buffer (known to every rank) */
  uint64_t candidates;
  float scan_ms, order_ms; /* this rank's scan kernel(s) / ordering */
  float gather_ms;         /* count exchange + expand into rank 0's buffer + closing barrier */
  int32_t transport;
  int32_t launches;
} acg_shard_stats;
/* Collective.  `hay` holds this rank's slice of the global haystack: its byte 0 is global offset
 * `hay_global_offset`, `hay_len` bytes are readable, and it must cover the rank's
 * [read_lo, own_hi) of acg_shard_plan for the span (ACG_E_INVALID_SPAN otherwise).  hay_on_device
 * != 0: `hay` is a device pointer on the communicator's device; 0: a host pointer (the copy is
 * pipelined with the scan, as in acg_find_overlapping).  On rank 0, *d_matches receives a device
 * pointer to *n_total acg_match records in GLOBAL offsets, in the reference's iteration order,
 * valid until the next call on this communicator; if h_out != NULL they are also copied to the host
 * (ACG_E_OVERFLOW with *n_total set if h_cap is too small; acg_comm_fetch then gets them without
 * another scan).  Other ranks receive NULL / the total. */
int acg_find_overlapping_sharded(const acg_dfa* dfa, acg_comm* comm, const void* hay, int hay_on_device,
                                 uint64_t hay_len, uint64_t hay_global_offset, uint64_t span_start,
                                 uint64_t span_end, const acg_match** d_matches, uint64_t* n_total,
                                 acg_match* h_out, uint64_t h_cap, acg_shard_stats* stats);
/* The same search in two halves, for callers that run one sharded search after another (a stream of
 * haystack batches): _begin returns once this rank's records are on their way into rank 0's buffer,
 * _wait completes the step.  Up to two steps may be in flight, so the scan of batch k + 1 overlaps
 * the NVLink transfer of batch k's records (rank 0's buffer has two halves; the records of a step
 * stay valid until the step after the next begins).  Both are collective; every rank must issue
 * begin / wait in the same order.  *ticket identifies the step for _wait. */
int acg_find_overlapping_sharded_begin(const acg_dfa* dfa, acg_comm* comm, const void* hay, int hay_on_device,
                                       uint64_t hay_len, uint64_t hay_global_offset, uint64_t span_start,
                                       uint64_t span_end, int* ticket);
int acg_find_overlapping_sharded_wait(acg_comm* comm, int ticket, const acg_match** d_matches, uint64_t* n_total,
                                      acg_match* h_out, uint64_t h_cap, acg_shard_stats* stats);
/* In the begin / wait form the records travel by copy engine: the expand kernel writes them into this
 * rank's own HBM and one device-to-device copy over NVLink puts them at their place in rank 0's
 * buffer, so the next step's scan has every SM while they are under way (the blocking call stores them
 * from the kernel itself, which is the shorter path for a single step).
 *
 * Device timestamps around a stream of steps: acg_comm_mark(comm, 0) before the first _begin and
 * acg_comm_mark(comm, 1) after the last _wait each wait for the device to drain and record a CUDA
 * event; acg_comm_mark_elapsed_ms gives the time between them on the device's clock. */
int acg_comm_mark(acg_comm* comm, int which);
int acg_comm_mark_elapsed_ms(const acg_comm* comm, float* ms);
/* Rank 0: copy the records of the most recent sharded search to the host; *n_out = their number. */
int acg_comm_fetch(const acg_comm* comm, acg_match* out, uint64_t cap, uint64_t* n_out);
/* Rank 0: the same records in page-locked host memory owned by the communicator (one full-speed
 * device-to-host copy, no staging through pageable memory); *view stays valid until the next
 * acg_comm_fetch_view or sharded search on this communicator. */
int acg_comm_fetch_view(acg_comm* comm, const acg_match** view, uint64_t* n_out);
/* Count + FNV-1a of the ordered (pid, start, end) stream of the most recent sharded search
 * (rank 0), comparable with acg_count_overlapping_dev on one GPU over the same haystack. */
int acg_comm_checksum(const acg_comm* comm, uint64_t* n_out, uint64_t* fnv);

/* ---- packed searcher: packed::Config / Builder / Searcher, src/packed/api.rs ----------------
 * The reference's standalone "packed" API is Teddy (or Rabin-Karp) over a small pattern set with
 * leftmost semantics.  On the device its role is played by the same K3/K3b kernel pair that serves
 * AhoCorasick (a fingerprint prefilter feeding the DFA verifier); what this API keeps from the
 * reference is the construction contract -- when Builder::build returns None -- and the search
 * results, which are the leftmost-first / leftmost-longest non-overlapping matches. */
enum { ACG_PACKED_FORCE_NONE = 0, ACG_PACKED_FORCE_TEDDY = 1, ACG_PACKED_FORCE_RABINKARP = 2 };
typedef struct {
  int32_t match_kind;               /* ACG_LEFTMOST_FIRST (default) | ACG_LEFTMOST_LONGEST, api.rs:28-46 */
  int32_t force;                    /* Config::only_teddy / only_rabin_karp, api.rs:143-190 */
  int32_t only_teddy_fat;           /* -1 = None, 0 / 1 = Some(false / true), api.rs:158 */
  int32_t only_teddy_256bit;        /* -1 = None, 0 / 1 = Some(false / true), api.rs:170 */
  int32_t heuristic_pattern_limits; /* default 1, api.rs:196 */
} acg_packed_config;
typedef struct acg_packed acg_packed;
void acg_packed_config_default(acg_packed_config* c);
/* Builder::extend + Builder::build (api.rs:253-345).  Returns ACG_OK with *out == NULL where the
 * reference returns None: no patterns, an empty pattern, more than 128 patterns, or a pattern set
 * Teddy declines (teddy/builder.rs:98-231, decided as on x86-64 with AVX2). */
int acg_packed_build(const uint8_t* const* patterns, const uint64_t* lens, uint64_t n,
                     const acg_packed_config* cfg, acg_packed** out);
/* Same decision and tables without CUDA; searches return ACG_E_NO_DEVICE. */
int acg_packed_build_host(const uint8_t* const* patterns, const uint64_t* lens, uint64_t n,
                          const acg_packed_config* cfg, acg_packed** out);
void acg_packed_free(acg_packed* s);
/* Searcher::find_iter (api.rs:580) -- and find_in (:529) applied repeatedly for a sub-span:
 * non-overlapping leftmost matches inside [span_start, span_end).  Two-call protocol on
 * ACG_E_OVERFLOW; ACG_E_INVALID_SPAN where the reference's slice indexing panics. */
int acg_packed_find_iter(const acg_packed* s, const uint8_t* hay, uint64_t hay_len,
                         uint64_t span_start, uint64_t span_end, acg_match* out, uint64_t cap,
                         uint64_t* n_out);
/* Searcher::find (api.rs:491) / find_in (:529). */
int acg_packed_find(const acg_packed* s, const uint8_t* hay, uint64_t hay_len, uint64_t span_start,
                    uint64_t span_end, acg_match* out, int* found);
int acg_packed_match_kind(const acg_packed* s);        /* api.rs:612 */
uint64_t acg_packed_minimum_len(const acg_packed* s);  /* api.rs:627: 0 for Rabin-Karp, else Teddy's */
uint64_t acg_packed_memory_usage(const acg_packed* s); /* api.rs:634 (heap of the device engine's tables) */
uint64_t acg_packed_patterns_len(const acg_packed* s);
/* Which searcher the reference would run: returns 1 for Teddy (and fills fat / mask_len /
 * vector_bytes), 0 for Rabin-Karp. */
int acg_packed_searcher_variant(const acg_packed* s, int* fat, int* mask_len, int* vector_bytes);

/* per-call statistics of the most recent search on this handle (bench glue) */
typedef struct {
  int32_t engine;
  int32_t launches;          /* kernels launched by the library in the call */
  uint64_t candidates;       /* prefilter survivors (ACG_ENGINE_PREFILTER) */
  uint64_t raw_matches;      /* tuples appended before ordering/stitching */
  float scan_ms;             /* dominant scan kernel(s) */
  float order_ms;            /* ordering / compaction */
  float h2d_ms, d2h_ms;
} acg_stats;
int acg_last_stats(const acg_dfa* dfa, acg_stats* out);

const char* acg_strerror(int code);
/* number of CUDA devices visible (0 if none / driver missing) */
int acg_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
