"""GPU parity of the packed searcher mirror (src/packed/api.rs): the reference's own packed test
matrix (src/packed/tests.rs:380-504: 6 configurations x 2 match kinds x its vectors, with a subset
of the "Z" padding variations, :42-92) and a randomized differential against the oracle's
restatement of Teddy / Rabin-Karp.  Searches run through acg_packed_find_iter / acg_packed_find."""
import random

import pytest

import golden_util as G
import oracle_py as O
from aho_corasick_b200 import packed

pytestmark = pytest.mark.gpu

PK = G.load("packed_vectors.json")

CONFIGS = {
    "default": ({}, lambda c: c),
    "teddy": ({"force": 1}, lambda c: c.only_teddy(True)),
    "teddy_ssse3": ({"force": 1, "only_teddy_256bit": 0}, lambda c: c.only_teddy(True).only_teddy_256bit(False)),
    "teddy_avx2": ({"force": 1, "only_teddy_256bit": 1}, lambda c: c.only_teddy(True).only_teddy_256bit(True)),
    "teddy_fat": ({"force": 1, "only_teddy_fat": 1}, lambda c: c.only_teddy(True).only_teddy_fat(True)),
    "rabinkarp": ({"force": 2}, lambda c: c.only_rabin_karp(True)),
}
KINDS = [("PACKED_LEFTMOST_FIRST", 0, packed.MatchKind.LeftmostFirst),
         ("PACKED_LEFTMOST_LONGEST", 1, packed.MatchKind.LeftmostLongest)]


def tuples(ms):
    return [m.as_tuple() for m in ms]


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("coll,okind,kind", KINDS)
def test_packed_searcher_matrix(coll, okind, kind, cfg):
    for t in G.collection(PK, coll):
        s = CONFIGS[cfg][1](packed.Config().match_kind(kind)).builder().extend(t["patterns"]).build()
        assert s is not None, (t["name"], cfg)  # on x86-64 the reference panics if None
        want = [tuple(m) for m in t["matches"]]
        for off in (0, 1, 15, 16, 17, 33, 64, 260):
            z = b"Z" * off
            sh = [(p, a + off, b + off) for p, a, b in want]
            assert tuples(s.find_iter(z + t["haystack"])) == sh, (t["name"], cfg, off, "prefix")
            assert tuples(s.find_iter(t["haystack"] + z)) == want, (t["name"], cfg, off, "suffix")
            assert tuples(s.find_iter(z + t["haystack"] + z)) == sh, (t["name"], cfg, off, "both")
        first = s.find(t["haystack"])
        assert (first.as_tuple() if first is not None else None) == (want[0] if want else None), (t["name"], cfg)


@pytest.mark.parametrize("okind,kind", [(0, packed.MatchKind.LeftmostFirst), (1, packed.MatchKind.LeftmostLongest)])
def test_packed_random_differential(okind, kind):
    rng = random.Random(11 + okind)
    for it in range(60):
        alphabet = [b"ab", b"abcd", b"aAbBcC ", b"abcdefghijklmnopqrstuvwxyz"][it % 4]
        n = rng.choice([1, 2, 5, 20, 60])
        pats = [bytes(rng.choice(alphabet) for _ in range(rng.randint(1, rng.choice([3, 8, 20])))) for _ in range(n)]
        hay = bytes(rng.choice(alphabet) for _ in range(rng.choice([0, 1, 7, 40, 1000, 20000])))
        o = O.PackedOracle(pats, kind=okind)
        s = packed.Config().match_kind(kind).builder().extend(pats).build()
        assert (s is not None) == o.built, (it, n)
        if s is None:
            continue
        assert tuples(s.find_iter(hay)) == o.find_iter(hay), (it, pats[:4], len(hay))
        sub = (min(3, len(hay)), max(min(3, len(hay)), len(hay) - 2))
        got = s.find_in(hay, sub)
        # find_in(span) == first match of the leftmost iteration restricted to the span: compare with
        # the oracle run on the slice (matches never look outside the span, api.rs:529-560)
        ref = o.find_iter(hay[sub[0]:sub[1]])
        want = (ref[0][0], ref[0][1] + sub[0], ref[0][2] + sub[0]) if ref else None
        assert (got.as_tuple() if got is not None else None) == want, (it, sub)


def test_concurrent_searches_on_one_handle():
    """The reference's automata are Send + Sync and searched through &self from many threads
    (src/lib.rs:274-326); an acg_dfa handle leases a workspace of its pool to every search.  Several
    threads share one AhoCorasick and one packed Searcher (ctypes releases the GIL during a call)."""
    from concurrent.futures import ThreadPoolExecutor

    import aho_corasick_b200 as ab
    rng = random.Random(5)
    pats = [bytes(rng.choice(b"abcd") for _ in range(rng.randint(2, 6))) for _ in range(40)]
    hays = [bytes(rng.choice(b"abcd") for _ in range(rng.choice([10, 500, 20000]))) for _ in range(24)]
    ac = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).build(pats)
    lf = packed.Config().builder().extend(pats).build()
    o = O.Oracle(pats, kind=O.KIND_DFA)
    olf = O.Oracle(pats, match_kind=1, kind=O.KIND_DFA)
    want = [(o.find_overlapping_iter(h), o.find_iter(h), olf.find_iter(h)) for h in hays]

    def work(i):
        h = hays[i % len(hays)]
        return (i % len(hays), tuples(ac.find_overlapping_iter(h)), tuples(ac.find_iter(h)), tuples(lf.find_iter(h)))

    with ThreadPoolExecutor(4) as ex:
        for i, a, b, c in ex.map(work, range(96)):
            assert (a, b, c) == want[i], i
