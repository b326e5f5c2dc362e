"""Pin the CPU oracle against every golden vector the reference holds for this path
(src/tests.rs:96-642 through the matrix :653-1323; src/packed/tests.rs:129-504;
README.md:34-77 and the doc examples listed in SURVEY.md section 8c)."""
import pytest

import bruteforce
import golden_util as G
import oracle_py as O

AC = G.load("ac_vectors.json")
PK = G.load("packed_vectors.json")


@pytest.mark.parametrize("combo", list(G.COMBO))
@pytest.mark.parametrize("coll,kind", G.NON_OVERLAPPING_COLLECTIONS)
def test_find_iter_matrix(coll, kind, combo):
    for t in G.collection(AC, coll):
        ac = O.Oracle(t["patterns"], match_kind=kind, **G.COMBO[combo])
        assert ac.find_iter(t["haystack"]) == t["matches"], (t["name"], combo)


@pytest.mark.parametrize("combo", list(G.COMBO))
def test_find_overlapping_iter_matrix(combo):
    for t in G.collection(AC, "AC_STANDARD_OVERLAPPING"):
        ac = O.Oracle(t["patterns"], match_kind=G.STANDARD, **G.COMBO[combo])
        assert ac.find_overlapping_iter(t["haystack"]) == t["matches"], (t["name"], combo)


@pytest.mark.parametrize("combo", list(G.ANCHORED_COMBO))
@pytest.mark.parametrize("coll,kind", G.ANCHORED)
def test_anchored_matrix(coll, kind, combo):
    for t in G.collection(AC, coll):
        ac = O.Oracle(t["patterns"], match_kind=kind, **G.ANCHORED_COMBO[combo])
        assert ac.find_iter(t["haystack"], anchored=True) == t["matches"], (t["name"], combo)


# src/tests.rs:1182-1323
ACASEI = [
    ("standard_default", ["ASCII_CASE_INSENSITIVE"], G.STANDARD, {"prefilter": False}, False),
    ("standard_nfa", ["ASCII_CASE_INSENSITIVE"], G.STANDARD, {"kind": G.KIND_NFA, "prefilter": False}, False),
    ("standard_dfa", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_NON_OVERLAPPING"], G.STANDARD,
     {"kind": G.KIND_DFA}, False),
    ("overlapping_default", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_OVERLAPPING"], G.STANDARD, {}, True),
    ("overlapping_nfa", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_OVERLAPPING"], G.STANDARD,
     {"kind": G.KIND_NFA}, True),
    ("overlapping_dfa", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_OVERLAPPING"], G.STANDARD,
     {"kind": G.KIND_DFA}, True),
    ("leftmost_first_default", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_NON_OVERLAPPING"],
     G.LEFTMOST_FIRST, {}, False),
    ("leftmost_first_nfa", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_NON_OVERLAPPING"],
     G.LEFTMOST_FIRST, {"kind": G.KIND_NFA}, False),
    ("leftmost_first_dfa", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_NON_OVERLAPPING"],
     G.LEFTMOST_FIRST, {"kind": G.KIND_DFA}, False),
    ("leftmost_longest_default", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_NON_OVERLAPPING"],
     G.LEFTMOST_LONGEST, {}, False),
    ("leftmost_longest_nfa", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_NON_OVERLAPPING"],
     G.LEFTMOST_LONGEST, {"kind": G.KIND_NFA}, False),
    ("leftmost_longest_dfa", ["ASCII_CASE_INSENSITIVE", "ASCII_CASE_INSENSITIVE_NON_OVERLAPPING"],
     G.LEFTMOST_LONGEST, {"kind": G.KIND_DFA}, False),
]


@pytest.mark.parametrize("name,groups,kind,kw,overlapping", ACASEI, ids=[a[0] for a in ACASEI])
def test_ascii_case_insensitive(name, groups, kind, kw, overlapping):
    for g in groups:
        for t in AC["groups"][g]:
            ac = O.Oracle(t["patterns"], match_kind=kind, ascii_case_insensitive=True, **kw)
            got = ac.find_overlapping_iter(t["haystack"]) if overlapping else ac.find_iter(t["haystack"])
            assert got == t["matches"], (t["name"], name)


def test_readme_examples():
    # README.md:34-50 (= BASELINE config 1)
    ac = O.Oracle([b"apple", b"maple", b"Snapple"])
    hay = b"Nobody likes maple in their apple flavored Snapple."
    assert ac.find_iter(hay) == [(1, 13, 18), (0, 28, 33), (2, 43, 50)]
    assert ac.kind == O.KIND_DFA  # auto => DFA for <= 100 patterns
    # README.md:58-77
    ac = O.Oracle([b"apple", b"maple", b"snapple"], ascii_case_insensitive=True)
    assert ac.find_iter(hay) == [(1, 13, 18), (0, 28, 33), (2, 43, 50)]
    # src/ahocorasick.rs:442-468 / src/automaton.rs:756-779
    ac = O.Oracle([b"append", b"appendage", b"app"])
    assert ac.find_overlapping_iter(b"append the app to the appendage") == [
        (2, 0, 3), (0, 0, 6), (2, 11, 14), (2, 22, 25), (0, 22, 28), (1, 22, 31)]
    # src/ahocorasick.rs:499-560
    pats, hay = [b"append", b"appendage", b"app"], b"append the app to the appendage"
    assert O.Oracle(pats).find_iter(hay) == [(2, 0, 3), (2, 11, 14), (2, 22, 25)]
    assert O.Oracle(pats, match_kind=G.LEFTMOST_FIRST).find_iter(hay) == [(0, 0, 6), (2, 11, 14), (0, 22, 28)]
    assert O.Oracle(pats, match_kind=G.LEFTMOST_LONGEST).find_iter(hay) == [(0, 0, 6), (2, 11, 14), (1, 22, 31)]


def test_error_behaviour():
    # src/tests.rs:1407-1511: overlapping unsupported for leftmost kinds; anchored mismatch is an error
    for k in (G.LEFTMOST_FIRST, G.LEFTMOST_LONGEST):
        with pytest.raises(O.OracleError) as e:
            O.Oracle([b"a"], match_kind=k).find_overlapping_iter(b"a")
        assert e.value.code == -13
    with pytest.raises(O.OracleError) as e:
        O.Oracle([b"a"]).find_iter(b"a", anchored=True)
    assert e.value.code == -10
    with pytest.raises(O.OracleError) as e:
        O.Oracle([b"a"], start_kind=G.START_ANCHORED).find_iter(b"a")
    assert e.value.code == -11
    with pytest.raises(O.OracleError) as e:  # src/automaton.rs:415-417
        O.Oracle([b"a"], start_kind=G.START_BOTH).find_overlapping_iter(b"a", anchored=True)
    assert e.value.code == -10


def test_regression_prefilter_stays_in_bounds():
    # src/tests.rs:1523-1530: Teddy-as-prefilter must honour Input::range
    ac = O.Oracle([b"sam", b"frodo", b"pippin", b"merry", b"gandalf", b"sauron"], match_kind=G.LEFTMOST_FIRST)
    assert ac.prefilter_kind == O.PRE_PACKED
    hay = b"foo gandalf"
    assert ac.try_find(hay, span=(0, 10)) is None
    assert ac.try_find(hay, span=(0, 11)) == (4, 4, 11)


def test_prefilter_decision():
    # src/util/prefilter.rs:163-305
    assert O.Oracle([b"foo"]).prefilter_kind == O.PRE_MEMMEM
    assert O.Oracle([b"foo", b"bar"]).prefilter_kind == O.PRE_START_BYTES
    assert O.Oracle([b"foo", b""]).prefilter_kind == O.PRE_NONE
    assert O.Oracle([b"foo", b"bar"], prefilter=False).prefilter_kind == O.PRE_NONE
    many = [bytes([97 + i]) + b"xyz" for i in range(8)]
    assert O.Oracle(many).prefilter_kind in (O.PRE_RARE_BYTES, O.PRE_NONE, O.PRE_START_BYTES)  # Standard: no packed
    lf = O.Oracle(many, match_kind=G.LEFTMOST_FIRST)
    assert lf.prefilter_kind in (O.PRE_PACKED, O.PRE_RARE_BYTES)
    pats50 = [bytes([33 + (i * 7) % 90, 40 + (i * 11) % 80, 50 + (i * 13) % 70, 35 + (i * 17) % 85, 48 + i % 60])
              for i in range(50)]
    o = O.Oracle(pats50, match_kind=G.LEFTMOST_FIRST)
    assert o.prefilter_kind == O.PRE_PACKED
    assert o.packed_variant() == {"fat": True, "mask_len": 4, "vector_bytes": 32}
    assert O.Oracle(pats50, match_kind=G.LEFTMOST_FIRST, ascii_case_insensitive=True).prefilter_kind == O.PRE_NONE


# ---- packed (src/packed/tests.rs:380-504) --------------------------------------------------
PACKED_CONFIGS = {
    "default": {},
    "teddy": {"force": 1},
    "teddy_ssse3": {"force": 1, "only_teddy_256bit": 0},
    "teddy_avx2": {"force": 1, "only_teddy_256bit": 1},
    "teddy_fat": {"force": 1, "only_teddy_fat": 1},
    "rabinkarp": {"force": 2},
}


def _variations(t, count=261):
    # SearchTest::variations, src/packed/tests.rs:42-92
    for off in range(count):
        z = b"Z" * off
        sh = [(p, s + off, e + off) for p, s, e in t["matches"]]
        yield off, z + t["haystack"], sh
        yield off, t["haystack"] + z, list(t["matches"])
        yield off, z + t["haystack"] + z, sh


@pytest.mark.parametrize("cfg", list(PACKED_CONFIGS))
@pytest.mark.parametrize("coll,kind", [("PACKED_LEFTMOST_FIRST", 0), ("PACKED_LEFTMOST_LONGEST", 1)])
def test_packed_matrix(coll, kind, cfg):
    for t in G.collection(PK, coll):
        s = O.PackedOracle(t["patterns"], kind=kind, **PACKED_CONFIGS[cfg])
        assert s.built, (t["name"], cfg)  # on x86_64 the reference panics if None
        for off, hay, want in _variations(t):
            assert s.find_iter(hay) == want, (t["name"], cfg, off)


# ---- oracle vs declarative spec (two independent implementations agreeing) ------------------
def test_oracle_vs_bruteforce_on_golden():
    for coll, kind in G.NON_OVERLAPPING_COLLECTIONS:
        for t in G.collection(AC, coll):
            assert bruteforce.find_iter(t["patterns"], t["haystack"], kind) == t["matches"], t["name"]
    for t in G.collection(AC, "AC_STANDARD_OVERLAPPING"):
        assert bruteforce.find_overlapping(t["patterns"], t["haystack"]) == t["matches"], t["name"]


# ---- src/tests.rs:1537-1660: regressions around the memchr-class prefilters (results only) ------
def test_regression_ascii_case_insensitive_no_exponential():
    pat = "Tsubaki House-Triple Shot Vol01校花三姐妹".encode()
    assert O.Oracle([pat], ascii_case_insensitive=True).try_find(b"") is None


def test_regression_rare_byte_prefilter():
    # https://github.com/BurntSushi/aho-corasick/issues/53
    o = O.Oracle([b"ab/j/", b"x/"])
    assert o.try_find(b"ab/j/", earliest=True) is not None


def test_regression_case_insensitive_prefilter():
    for c in range(ord("a"), ord("z")):
        for c2 in range(ord("a"), ord("z")):
            needle = bytes([c, c2])
            o = O.Oracle([needle], ascii_case_insensitive=True, prefilter=True)
            assert len(o.find_iter(needle.upper())) == 1, needle
