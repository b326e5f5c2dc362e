// C++ facade test (include/acb200.hpp): written to read like the reference's own doc tests.
// Built with g++ against libacb200.so and run by tests/test_gpu_cpp_facade.py on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <tuple>
#include <vector>

#include "acb200.hpp"

using namespace acb200;
using T3 = std::tuple<unsigned, unsigned long, unsigned long>;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

static std::vector<T3> triples(const MatchIter& it) {
  std::vector<T3> v;
  for (const Match& m : it) v.emplace_back(m.pattern(), m.start(), m.end());
  return v;
}

int main() {
  // README.md:34-50 (BASELINE config 1)
  {
    std::vector<std::string> patterns = {"apple", "maple", "Snapple"};
    std::string haystack = "Nobody likes maple in their apple flavored Snapple.";
    AhoCorasick ac = AhoCorasick::create(patterns);
    CHECK(triples(ac.find_iter(haystack)) == (std::vector<T3>{{1, 13, 18}, {0, 28, 33}, {2, 43, 50}}));
    CHECK(ac.kind() == AhoCorasickKind::DFA);  // auto selection for <= 100 patterns
    CHECK(ac.patterns_len() == 3 && ac.min_pattern_len() == 5 && ac.max_pattern_len() == 7);
  }
  // README.md:58-77
  {
    std::vector<std::string> patterns = {"apple", "maple", "snapple"};
    std::string haystack = "Nobody likes maple in their apple flavored Snapple.";
    AhoCorasick ac = AhoCorasick::builder().ascii_case_insensitive(true).build(patterns);
    CHECK(triples(ac.find_iter(haystack)) == (std::vector<T3>{{1, 13, 18}, {0, 28, 33}, {2, 43, 50}}));
  }
  // src/ahocorasick.rs:442-468 and :499-560
  {
    std::vector<std::string> patterns = {"append", "appendage", "app"};
    std::string haystack = "append the app to the appendage";
    AhoCorasick ac = AhoCorasick::create(patterns);
    CHECK(triples(ac.find_overlapping_iter(haystack)) ==
          (std::vector<T3>{{2, 0, 3}, {0, 0, 6}, {2, 11, 14}, {2, 22, 25}, {0, 22, 28}, {1, 22, 31}}));
    CHECK(triples(ac.find_iter(haystack)) == (std::vector<T3>{{2, 0, 3}, {2, 11, 14}, {2, 22, 25}}));
    AhoCorasick lf = AhoCorasick::builder().match_kind(MatchKind::LeftmostFirst).build(patterns);
    CHECK(triples(lf.find_iter(haystack)) == (std::vector<T3>{{0, 0, 6}, {2, 11, 14}, {0, 22, 28}}));
    AhoCorasick ll = AhoCorasick::builder().match_kind(MatchKind::LeftmostLongest).build(patterns);
    CHECK(triples(ll.find_iter(haystack)) == (std::vector<T3>{{0, 0, 6}, {2, 11, 14}, {1, 22, 31}}));
    Match m;
    CHECK(lf.find(haystack, &m) && m == Match(0, 0, 6));
    CHECK(lf.is_match(haystack) && !lf.is_match("xyz"));
    // Input::span: matches must lie inside the span, the automaton starts cold at span.start
    CHECK(triples(ll.find_iter(Input(haystack).span(1, 28))) == (std::vector<T3>{{2, 11, 14}, {0, 22, 28}}));
  }
  // error behaviour, src/tests.rs:1407-1511 and src/automaton.rs:397-423
  {
    std::vector<std::string> patterns = {"a"};
    AhoCorasick lf = AhoCorasick::builder().match_kind(MatchKind::LeftmostFirst).build(patterns);
    auto r = lf.try_find_overlapping_iter("a");
    CHECK(r.is_err() && MatchError(r.error).kind() == MatchErrorKind::UnsupportedOverlapping);
    AhoCorasick un = AhoCorasick::create(patterns);
    auto r2 = un.try_find_iter(Input("a").anchored(Anchored::Yes));
    CHECK(r2.is_err() && MatchError(r2.error).kind() == MatchErrorKind::InvalidInputAnchored);
    AhoCorasick an = AhoCorasick::builder().start_kind(StartKind::Anchored).build(patterns);
    auto r3 = an.try_find_iter("a");
    CHECK(r3.is_err() && MatchError(r3.error).kind() == MatchErrorKind::InvalidInputUnanchored);
    CHECK(triples(an.find_iter(Input("aab").anchored(Anchored::Yes))) == (std::vector<T3>{{0, 0, 1}, {0, 1, 2}}));
    bool threw = false;
    try { lf.find_overlapping_iter("a"); } catch (const MatchError&) { threw = true; }
    CHECK(threw);
  }
  // empty patterns (src/tests.rs:521-546) go through the walk / sequential engines
  {
    std::vector<std::string> patterns = {"", "a", ""};
    AhoCorasick ac = AhoCorasick::create(patterns);
    CHECK(triples(ac.find_overlapping_iter("a")) ==
          (std::vector<T3>{{0, 0, 0}, {2, 0, 0}, {1, 0, 1}, {0, 1, 1}, {2, 1, 1}}));
    CHECK(triples(ac.find_iter("a")) == (std::vector<T3>{{0, 0, 0}, {0, 1, 1}}));
  }
  if (failures == 0) std::printf("acb200.hpp facade: all checks passed\n");
  return failures == 0 ? 0 : 1;
}
