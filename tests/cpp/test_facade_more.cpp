// Second facade program: the parts of include/acb200.hpp added after the first one --
// OverlappingState, replace_all*, acb200::packed -- on a device (or on the dry-run library).
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
  // find_overlapping + OverlappingState, doc example of src/ahocorasick.rs:430-470
  {
    std::vector<std::string> patterns = {"append", "appendage", "app"};
    const std::string haystack = "append the app to the appendage";
    AhoCorasick ac = AhoCorasick::create(patterns);
    acb200::OverlappingState state = acb200::OverlappingState::start();
    std::vector<T3> got;
    for (;;) {
      ac.find_overlapping(haystack, state);
      auto m = state.get_match();
      if (!m) break;
      got.push_back(T3{m->pattern(), m->start(), m->end()});
    }
    CHECK(got == (std::vector<T3>{{2, 0, 3}, {0, 0, 6}, {2, 11, 14}, {2, 22, 25}, {0, 22, 28}, {1, 22, 31}}));
    // replace_all / replace_all_with, src/ahocorasick.rs:651-760
    AhoCorasick lf = AhoCorasick::builder().match_kind(MatchKind::LeftmostFirst).build(patterns);
    CHECK(lf.replace_all(haystack, std::vector<std::string>{"x", "y", "z"}) == "x the z to the xage");
    AhoCorasick ll = AhoCorasick::builder().match_kind(MatchKind::LeftmostLongest).build(patterns);
    CHECK(ll.replace_all_bytes(haystack, std::vector<std::string>{"x", "y", "z"}) == "x the z to the y");
    std::string dst;
    lf.replace_all_with(haystack, dst, [](const Match& m, std::string_view txt, std::string& out) {
      for (char c : txt) out.push_back(char(c - 32));
      return m.pattern() != 2;
    });
    CHECK(dst == "APPEND the APP to the appendage");
  }
  // aho_corasick::packed, doc example of src/packed/api.rs:365-395
  {
    namespace pk = acb200::packed;
    auto s = pk::Searcher::create(std::vector<std::string>{"foobar", "foo"});
    CHECK(s.has_value());
    CHECK(triples(s->find_iter("foobar")) == (std::vector<T3>{{0, 0, 6}}));
    Match m(9, 9, 9);
    CHECK(s->find("xxfooxx", &m) && m == Match(1, 2, 5));
    CHECK(!s->find_in(Input("xxfooxx").span(3, 7), &m));
    auto ll = pk::Config().match_kind(pk::MatchKind::LeftmostLongest).builder().add("ab").add("abcd").build();
    CHECK(ll.has_value() && triples(ll->find_iter("xabcdabx")) == (std::vector<T3>{{1, 1, 5}, {0, 5, 7}}));
  }
  if (failures == 0) std::printf("acb200.hpp facade (more): all checks passed\n");
  return failures == 0 ? 0 : 1;
}
