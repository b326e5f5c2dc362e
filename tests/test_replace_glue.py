"""CPU checks of the replace_all* host glue (src/automaton.rs:433-550): the splice loop runs over a
materialised match list, so it can be driven by the oracle's matches without a device.  The
examples are the reference's doc examples (src/ahocorasick.rs:651-760)."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import aho_corasick_b200 as ab  # noqa: E402
import oracle_py as O  # noqa: E402


class OracleBacked(ab.AhoCorasick):
    """The product class with `try_find_iter` answered by the oracle (host-only handle inside)."""

    def __init__(self, pats, match_kind):
        inner = ab.AhoCorasick.builder().host_only().match_kind(match_kind).build(pats)
        super().__init__(inner._h)
        inner._h = None  # ownership moved
        self._o = O.Oracle(pats, match_kind=int(match_kind))

    def try_find_iter(self, hay, span=None, anchored=ab.Anchored.No):
        return [ab.Match(*m) for m in self._o.find_iter(bytes(hay))]


PATS = [b"append", b"appendage", b"app"]
HAY = b"append the app to the appendage"


def test_replace_all_doc_examples():
    # src/ahocorasick.rs:651-760
    assert OracleBacked(PATS, ab.MatchKind.Standard).replace_all(HAY.decode(), ["x", "y", "z"]) == "zend the z to the zendage"
    assert OracleBacked(PATS, ab.MatchKind.LeftmostFirst).replace_all(HAY.decode(), ["x", "y", "z"]) == "x the z to the xage"
    assert OracleBacked(PATS, ab.MatchKind.LeftmostLongest).replace_all_bytes(HAY, [b"x", b"y", b"z"]) == b"x the z to the y"


def test_replace_all_with_closure_can_stop():
    ac = OracleBacked(PATS, ab.MatchKind.LeftmostFirst)
    dst = bytearray()
    ac.replace_all_with(HAY, dst, lambda m, txt, out: (out.extend(txt.upper()), m.pattern() != 2)[1])
    assert bytes(dst) == b"APPEND the APP to the appendage"  # stops after the first "app"


def test_replace_all_requires_one_replacement_per_pattern():
    ac = OracleBacked(PATS, ab.MatchKind.LeftmostFirst)
    with pytest.raises(ValueError):
        ac.replace_all_bytes(HAY, [b"x"])
    with pytest.raises(ValueError):
        ac.replace_all(HAY.decode(), ["x", "y"])


def test_str_flavour_skips_matches_that_split_a_code_point():
    # src/automaton.rs:514-518: a pattern may be a partial code point; such matches are skipped by
    # the &str flavour and replaced by the bytes flavour
    pats = [b"\xc3", b"b"]
    hay = "aéb"  # "a" C3 A9 "b"
    ac = OracleBacked(pats, ab.MatchKind.LeftmostFirst)
    assert ac.replace_all(hay, ["?", "B"]) == "aéB"
    assert ac.replace_all_bytes(hay.encode(), [b"?", b"B"]) == b"a?\xa9B"
    assert ac.replace_all("", ["?", "B"]) == ""


def test_char_boundary_rule():
    b = "aé\U0001F600z".encode()
    n = len(b)
    want = [i == 0 or i == n or (i < n and (b[i] & 0xC0) != 0x80) for i in range(n + 2)]
    got = [ab.AhoCorasick._is_char_boundary(b, n, i) for i in range(n + 2)]
    assert got == want and got[:4] == [True, True, False, True] and got[n + 1] is False


# ---- stream search / replace (src/automaton.rs:1059-1256, 567-636) ----------------------------
import io  # noqa: E402

import numpy as np  # noqa: E402

import golden_util as G  # noqa: E402
from aho_corasick_b200 import workload as W  # noqa: E402

AC = G.load("ac_vectors.json")


class OracleBackedNp(OracleBacked):
    def try_find_iter_np(self, hay, span=None, anchored=ab.Anchored.No):
        return self._o.find_iter_np(np.ascontiguousarray(hay), span=span)


def _apply(hay, matches, reps):
    out, last = bytearray(), 0
    for pid, s, e in matches:
        out += hay[last:s] + reps[pid]
        last = e
    return bytes(out + hay[last:])


def test_stream_find_iter_equals_find_iter_on_reference_vectors():
    # the stream rows of src/tests.rs:999-1036: every Standard vector without empty patterns
    n = 0
    for t in G.collection(AC, "AC_STANDARD_NON_OVERLAPPING"):
        if any(len(p) == 0 for p in t["patterns"]):
            continue
        ac = OracleBackedNp(t["patterns"], ab.MatchKind.Standard)
        for chunk in (1, 2, 3, 7, 64 << 20):
            got = [m.as_tuple() for m in ac.stream_find_iter(io.BytesIO(t["haystack"]), chunk_bytes=chunk)]
            assert got == t["matches"], (t["name"], chunk)
            out = io.BytesIO()
            reps = [b"<%d>" % i for i in range(len(t["patterns"]))]
            ac.stream_replace_all(io.BytesIO(t["haystack"]), out, reps, chunk_bytes=chunk)
            assert out.getvalue() == _apply(t["haystack"], t["matches"], reps), (t["name"], chunk)
        n += 1
    assert n > 20


def test_stream_unsupported_configurations():
    # src/automaton.rs:1087-1103
    with pytest.raises(ab.MatchError) as e:
        list(OracleBackedNp([b"a"], ab.MatchKind.LeftmostFirst).stream_find_iter(io.BytesIO(b"a")))
    assert e.value.kind == "UnsupportedStream"
    with pytest.raises(ab.MatchError) as e:
        list(OracleBackedNp([b"a", b""], ab.MatchKind.Standard).stream_find_iter(io.BytesIO(b"a")))
    assert e.value.kind == "UnsupportedEmpty"
    with pytest.raises(ValueError):
        OracleBackedNp([b"a", b"b"], ab.MatchKind.Standard).stream_replace_all(io.BytesIO(b"a"), io.BytesIO(), [b"x"])


def test_stream_with_straddling_matches_and_chunked_output():
    pats = W.make_patterns(200, 77)
    t = np.empty(1 << 20, dtype=np.uint8)
    W.fill_haystack(t, 99)
    W.plant(t, pats, 5, period=512, window=256)
    ac = OracleBackedNp(pats, ab.MatchKind.Standard)
    want = ac._o.find_iter(t)
    assert len(want) > 1500
    data = t.tobytes()
    for chunk in ((1 << 16) + 13, 4099, 1 << 20, 5 << 20):
        got = [m.as_tuple() for m in ac.stream_find_iter(io.BytesIO(data), chunk_bytes=chunk)]
        assert got == want, chunk
    # the chunk stream is a partition of the input: text and matches in stream order
    pos, pieces = 0, []
    for c in ac._stream_chunks(io.BytesIO(data), 70001):
        if c[0] == "bytes":
            pieces.append(c[1])
            pos += len(c[1])
        else:
            assert c[1].start() == pos and data[c[1].start():c[1].end()] == c[2]
            pieces.append(c[2])
            pos = c[1].end()
    assert b"".join(pieces) == data
    reps = [b"<%d>" % i for i in range(len(pats))]
    out = io.BytesIO()
    ac.try_stream_replace_all(io.BytesIO(data), out, reps, chunk_bytes=1 << 15)
    assert out.getvalue() == _apply(data, want, reps)
    # closure flavour: sees the matched bytes and the writer
    out = io.BytesIO()
    ac.stream_replace_all_with(io.BytesIO(data), out, lambda m, txt, w: w.write(txt.upper()), chunk_bytes=1 << 17)
    assert out.getvalue().lower() == data.lower() and len(out.getvalue()) == len(data)


# ---- OverlappingState / find_overlapping (src/ahocorasick.rs:470, 1184; automaton.rs:756-840) ---
class OracleBackedOverlapping(OracleBacked):
    def try_find_overlapping_iter(self, hay, span=None, anchored=ab.Anchored.No):
        if self.match_kind() != ab.MatchKind.Standard:
            raise ab.MatchError(-13)
        return [ab.Match(*m) for m in self._o.find_overlapping_iter(bytes(hay), span=span)]


def test_find_overlapping_state_doc_example():
    ac = OracleBackedOverlapping(PATS, ab.MatchKind.Standard)
    state = ab.OverlappingState.start()
    assert state.get_match() is None
    got = []
    while True:
        ac.find_overlapping(HAY, state)
        m = state.get_match()
        if m is None:
            break
        got.append(m.as_tuple())
    assert got == [(2, 0, 3), (0, 0, 6), (2, 11, 14), (2, 22, 25), (0, 22, 28), (1, 22, 31)]
    ac.find_overlapping(HAY, state)      # stays exhausted
    assert state.get_match() is None
    # a sub-span, fresh state
    state = ab.OverlappingState.start()
    ac.try_find_overlapping(HAY, state, span=(11, 25))
    assert state.get_match().as_tuple() == (2, 11, 14)
    ac.try_find_overlapping(HAY, state, span=(11, 25))
    assert state.get_match().as_tuple() == (2, 22, 25)
    ac.try_find_overlapping(HAY, state, span=(11, 25))
    assert state.get_match() is None


def test_find_overlapping_errors_are_reported_each_call():
    ac = OracleBackedOverlapping(PATS, ab.MatchKind.LeftmostFirst)
    state = ab.OverlappingState.start()
    for _ in range(2):
        with pytest.raises(ab.MatchError) as e:
            ac.try_find_overlapping(HAY, state)
        assert e.value.kind == "UnsupportedOverlapping" and state.get_match() is None
