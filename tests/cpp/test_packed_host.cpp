// Host-only checks of acb200::packed (include/acb200.hpp): the construction contract of
// packed::Builder::build (src/packed/api.rs:253-322, teddy/builder.rs:98-231) and the error
// behaviour of a searcher that has no device.  Runs without a GPU.
#include <cstdio>
#include <string>
#include <vector>

#include "acb200.hpp"

namespace pk = acb200::packed;

static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

int main() {
  // no patterns / an empty pattern / a 129th pattern: Builder::build returns None
  CHECK(!pk::Config().host_only(true).builder().build().has_value());
  CHECK(!pk::Config().host_only(true).builder().add("a").add("").build().has_value());
  std::vector<std::string> many;
  for (int i = 0; i < 129; ++i) many.push_back(std::string(1, char('a' + i % 26)) + char('a' + i / 26));
  CHECK(!pk::Config().host_only(true).only_rabin_karp(true).builder().extend(many).build().has_value());
  many.pop_back();
  auto rk = pk::Config().host_only(true).only_rabin_karp(true).builder().extend(many).build();
  CHECK(rk.has_value() && rk->minimum_len() == 0 && rk->patterns_len() == 128);
  // Teddy's heuristic limits
  CHECK(!pk::Config().host_only(true).builder().extend(many).build().has_value());  // > 64 patterns
  CHECK(pk::Config().host_only(true).heuristic_pattern_limits(false).builder().extend(many).build().has_value());

  // the README-sized case: Slim AVX2 Teddy with 3-byte masks => minimum_len 32 + 3 - 1
  std::vector<std::string> pats = {"apple", "maple", "Snapple", "foo"};
  pk::Builder b = pk::Config().host_only(true).match_kind(pk::MatchKind::LeftmostLongest).builder();
  b.extend(pats);
  CHECK(b.len() == 4 && b.minimum_len() == 3);
  auto s = b.build();
  CHECK(s.has_value());
  CHECK(s->match_kind() == pk::MatchKind::LeftmostLongest);
  CHECK(s->minimum_len() == 34);
  CHECK(s->memory_usage() > 0);
  auto ssse3 = pk::Config().host_only(true).only_teddy_256bit(false).builder().extend(pats).build();
  CHECK(ssse3.has_value() && ssse3->minimum_len() == 18);
  CHECK(!pk::Config().host_only(true).only_teddy_fat(true).only_teddy_256bit(false).builder().extend(pats).build().has_value());
  auto fat = pk::Config().host_only(true).only_teddy_fat(true).builder().extend(pats).build();
  CHECK(fat.has_value() && fat->minimum_len() == 18);

  // searches: validation first, then "no device" -- there is no CPU search path
  bool threw = false;
  try { acb200::Match m; s->find_in(acb200::Input("abcd").span(3, 9), &m); } catch (const std::out_of_range&) { threw = true; }
  CHECK(threw);
  threw = false;
  try { s->find_iter("xx maple xx"); } catch (const acb200::DeviceError& e) { threw = e.code() == ACG_E_NO_DEVICE; }
  CHECK(threw);
  threw = false;
  try { acb200::Match m; s->find("xx maple xx", &m); } catch (const acb200::DeviceError& e) { threw = e.code() == ACG_E_NO_DEVICE; }
  CHECK(threw);

  // ---- replace glue (src/automaton.rs:498-550) over hand-made match lists -------------------
  {
    using acb200::Match;
    const std::string hay = "append the app to the appendage";
    // leftmost-first matches of {append, appendage, app} (src/ahocorasick.rs:651-690)
    const std::vector<Match> lf = {Match(0, 0, 6), Match(2, 11, 14), Match(0, 22, 28)};
    const std::vector<std::string> reps = {"x", "y", "z"};
    std::string dst;
    acb200::detail::splice(hay, lf, dst, [&](const Match& m, std::string_view, std::string& out) {
      out.append(reps[m.pattern()]);
      return true;
    }, true);
    CHECK(dst == "x the z to the xage");
    dst.clear();
    acb200::detail::splice(hay, lf, dst, [&](const Match& m, std::string_view txt, std::string& out) {
      for (char c : txt) out.push_back(char(c - 32));
      return m.pattern() != 2;
    }, false);
    CHECK(dst == "APPEND the APP to the appendage");  // stops after the first "app"
    // a match that splits a code point is skipped by the &str flavour only
    const std::string utf = "a\xc3\xa9" "b";
    const std::vector<Match> split = {Match(0, 1, 2), Match(1, 3, 4)};
    const std::vector<std::string> r2 = {"?", "B"};
    auto put = [&](const Match& m, std::string_view, std::string& out) { out.append(r2[m.pattern()]); return true; };
    dst.clear();
    acb200::detail::splice(utf, split, dst, put, true);
    CHECK(dst == "a\xc3\xa9" "B");
    dst.clear();
    acb200::detail::splice(utf, split, dst, put, false);
    CHECK(dst == "a?\xa9" "B");
    CHECK(acb200::detail::is_char_boundary(utf, 0) && acb200::detail::is_char_boundary(utf, 1) &&
          !acb200::detail::is_char_boundary(utf, 2) && acb200::detail::is_char_boundary(utf, 3) &&
          acb200::detail::is_char_boundary(utf, 4) && !acb200::detail::is_char_boundary(utf, 5));
    // replace_all on a host-only automaton: argument check first, then "no device"
    acb200::AhoCorasick ac;  // empty handle: patterns_len() == 0
    bool threw2 = false;
    try { ac.replace_all("abc", std::vector<std::string>{"x"}); } catch (const std::invalid_argument&) { threw2 = true; }
    CHECK(threw2);
  }

  // ---- OverlappingState on a host-only automaton: the error is reported on every call ----------
  {
    std::vector<std::string> p3 = {"append", "appendage", "app"};
    acb200::OverlappingState st = acb200::OverlappingState::start();
    CHECK(!st.get_match().has_value());
    acb200::AhoCorasick none;  // empty handle: acg_find_overlapping reports an invalid argument
    auto r1 = none.try_find_overlapping("append", st);
    CHECK(r1.is_err() && !st.get_match().has_value());
    auto r2 = none.try_find_overlapping("append", st);
    CHECK(r2.is_err());
  }

  if (failures == 0) std::printf("all checks passed\n");
  return failures == 0 ? 0 : 1;
}
