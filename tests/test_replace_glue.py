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
