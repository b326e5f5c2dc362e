"""The search-related doc examples of the reference (src/ahocorasick.rs, src/lib.rs / README),
written against the Python mirror.  `run_all(ab)` is executed on the CPU dry-run library by
tests/test_doc_examples_dry_run.py and on the GPU by tests/test_gpu_zz_doc_examples.py."""
import io

APP = ["append", "appendage", "app"]
APP_HAY = "append the app to the appendage"


def pids(ms):
    return [m.pattern() for m in ms]


def lib_and_readme(ab):
    # src/ahocorasick.rs:140-176 (= README.md:34-77)
    ac = ab.AhoCorasick.builder().ascii_case_insensitive(True).build(["apple", "maple", "snapple"])
    hay = "Nobody likes maple in their apple flavored Snapple."
    assert [m.as_tuple() for m in ac.find_iter(hay)] == [(1, 13, 18), (0, 28, 33), (2, 43, 50)]
    ac = ab.AhoCorasick.new(["fox", "brown", "quick"])
    assert ac.replace_all("The quick brown fox.", ["sloth", "grey", "slow"]) == "The slow grey sloth."


def new_builder_is_match(ab):
    # :232-242, :258-267, :300-310
    assert ab.AhoCorasick.new(["foo", "bar", "baz"]).find("xxx bar xxx").pattern() == 1
    ac = ab.AhoCorasick.builder().match_kind(ab.MatchKind.LeftmostFirst).build(["samwise", "sam"])
    assert ac.find("samwise").as_tuple() == (0, 0, 7)
    ac = ab.AhoCorasick.new(["foo", "bar", "quux", "baz"])
    assert ac.is_match("xxx bar xxx") and not ac.is_match("xxx qux xxx")


def find_under_each_match_kind(ab):
    # :330-403
    pats, hay = ["b", "abc", "abcd"], "abcd"

    def found(kind, inp=hay):
        m = ab.AhoCorasick.builder().match_kind(kind).build(pats).find(inp)
        return hay[m.start():m.end()]
    assert found(ab.MatchKind.Standard) == "b"
    assert found(ab.MatchKind.LeftmostFirst) == "abc"
    assert found(ab.MatchKind.LeftmostLongest) == "abcd"
    assert found(ab.MatchKind.LeftmostLongest, ab.Input(hay).earliest(True)) == "b"


def overlapping_state(ab):
    # :430-470 and :1055-1140
    ac = ab.AhoCorasick.new(APP)
    state = ab.OverlappingState.start()
    got = []
    while True:
        ac.find_overlapping(APP_HAY, state)
        if state.get_match() is None:
            break
        got.append(state.get_match().as_tuple())
    assert got == [(2, 0, 3), (0, 0, 6), (2, 11, 14), (2, 22, 25), (0, 22, 28), (1, 22, 31)]


def find_iter_under_each_match_kind(ab):
    # :495-561, :588-608
    def it(kind):
        return pids(ab.AhoCorasick.builder().match_kind(kind).build(APP).find_iter(APP_HAY))
    assert it(ab.MatchKind.Standard) == [2, 2, 2]
    assert it(ab.MatchKind.LeftmostFirst) == [0, 2, 0]
    assert it(ab.MatchKind.LeftmostLongest) == [0, 2, 1]
    assert pids(ab.AhoCorasick.new(APP).find_overlapping_iter(APP_HAY)) == [2, 0, 2, 2, 0, 1]


def replace_family(ab):
    # :636-832
    lf = ab.AhoCorasick.builder().match_kind(ab.MatchKind.LeftmostFirst).build(APP)
    assert lf.replace_all(APP_HAY, ["x", "y", "z"]) == "x the z to the xage"
    assert lf.replace_all_bytes(APP_HAY.encode(), [b"x", b"y", b"z"]) == b"x the z to the xage"
    dst = bytearray()
    lf.replace_all_with(APP_HAY.encode(), dst, lambda m, _, out: (out.extend(str(m.pattern()).encode()), True)[1])
    assert bytes(dst) == b"0 the 2 to the 0age"
    dst = bytearray()
    lf.replace_all_with(APP_HAY.encode(), dst, lambda m, _, out: (out.extend(str(m.pattern()).encode()), m.pattern() != 2)[1])
    assert bytes(dst) == b"0 the 2 to the appendage"


def stream_family(ab):
    # :880-905, :1640-1850
    ac = ab.AhoCorasick.new(APP)
    assert pids(ac.stream_find_iter(io.BytesIO(APP_HAY.encode()))) == [2, 2, 2]
    out = io.BytesIO()
    ac.stream_replace_all(io.BytesIO(APP_HAY.encode()), out, ["x", "y", "z"])
    assert out.getvalue() == b"zend the z to the zendage"
    out = io.BytesIO()
    ac.stream_replace_all_with(io.BytesIO(APP_HAY.encode()), out, lambda m, _, w: w.write(str(m.pattern()).encode()))
    assert out.getvalue() == b"2end the 2 to the 2endage"


def try_find_configurations(ab):
    # :935-1020
    pats, hay = ["b", "abc", "abcd"], "foo abcd"
    lf = ab.AhoCorasick.builder().match_kind(ab.MatchKind.LeftmostFirst).build(pats)
    m = lf.try_find(hay)
    assert hay[m.start():m.end()] == "abc"
    anch = ab.AhoCorasick.builder().match_kind(ab.MatchKind.LeftmostFirst).start_kind(ab.StartKind.Anchored).build(pats)
    assert anch.try_find(ab.Input(hay).anchored(ab.Anchored.Yes)) is None
    m = anch.try_find(ab.Input(hay).range(slice(4, None)).anchored(ab.Anchored.Yes))
    assert hay[m.start():m.end()] == "abc"
    m = lf.try_find(ab.Input(hay).earliest(True))
    assert hay[m.start():m.end()] == "b"


def getters(ab):
    # :1860-2021; memory_usage pins the table dimensions of the DFA (states x stride, match states)
    ac = ab.AhoCorasick.new(["foo", "bar", "quux", "baz"])
    assert ac.kind() == ab.AhoCorasickKind.DFA and ac.start_kind() == ab.StartKind.Unanchored
    assert ac.match_kind() == ab.MatchKind.Standard
    assert (ac.min_pattern_len(), ac.max_pattern_len(), ac.patterns_len()) == (3, 4, 4)
    assert ab.AhoCorasick.new(["foo", "", "quux", "baz"]).min_pattern_len() == 0
    pats = ["foobar", "bruce", "triskaidekaphobia", "springsteen"]
    assert ab.AhoCorasick.builder().kind(None).build(pats).memory_usage() == 5632
    assert ab.AhoCorasick.builder().kind(None).ascii_case_insensitive(True).build(pats).memory_usage() == 11136
    assert ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).ascii_case_insensitive(True).build(pats).memory_usage() == 11136


def packed_examples(ab):
    # src/packed/mod.rs:30-45 and src/packed/api.rs:60-80, 205-225, 375-605
    from aho_corasick_b200 import packed
    s = packed.Searcher.new(["foobar", "foo"])
    assert pids(s.find_iter("foobar")) == [0]
    ll = packed.Config.new().match_kind(packed.MatchKind.LeftmostLongest).builder().add("foo").add("foobar").build()
    assert pids(ll.find_iter("foobar")) == [1]
    assert pids(packed.Builder.new().add("foobar").add("foo").build().find_iter("foobar")) == [0]
    assert s.find("foobar").as_tuple() == (0, 0, 6)
    hay = "foofoobar"
    assert s.find_in(hay, (3, len(hay))).as_tuple() == (0, 3, 9)
    assert pids(s.find_iter("foobar fooba foofoo")) == [0, 1, 1, 1]
    assert s.match_kind() == packed.MatchKind.LeftmostFirst and s.minimum_len() > 0 and s.memory_usage() > 0


ALL = [packed_examples, lib_and_readme, new_builder_is_match, find_under_each_match_kind, overlapping_state,
       find_iter_under_each_match_kind, replace_family, stream_family, try_find_configurations, getters]


def run_all(ab):
    for f in ALL:
        f(ab)
