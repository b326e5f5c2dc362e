"""CPU checks of the packed searcher mirror (src/packed/api.rs): the construction contract --
when `Builder::build` returns None, which Teddy the reference would pick, `minimum_len` -- must
agree with the oracle's restatement for the reference's own test configurations
(src/packed/tests.rs:380-504) and for random pattern sets; searches need the device."""
import random
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import aho_corasick_b200 as ab  # noqa: E402
from aho_corasick_b200 import packed  # noqa: E402
import golden_util as G  # noqa: E402
import oracle_py as O  # noqa: E402

PK = G.load("packed_vectors.json")

# name -> (oracle kwargs, how to set the same thing on packed.Config)
CONFIGS = {
    "default": ({}, lambda c: c),
    "teddy": ({"force": 1}, lambda c: c.only_teddy(True)),
    "teddy_ssse3": ({"force": 1, "only_teddy_256bit": 0}, lambda c: c.only_teddy(True).only_teddy_256bit(False)),
    "teddy_avx2": ({"force": 1, "only_teddy_256bit": 1}, lambda c: c.only_teddy(True).only_teddy_256bit(True)),
    "teddy_fat": ({"force": 1, "only_teddy_fat": 1}, lambda c: c.only_teddy(True).only_teddy_fat(True)),
    "teddy_slim": ({"force": 1, "only_teddy_fat": 0}, lambda c: c.only_teddy(True).only_teddy_fat(False)),
    "fat_without_avx2": ({"only_teddy_fat": 1, "only_teddy_256bit": 0},
                         lambda c: c.only_teddy_fat(True).only_teddy_256bit(False)),
    "no_limits": ({"heuristic_pattern_limits": False}, lambda c: c.heuristic_pattern_limits(False)),
    "rabinkarp": ({"force": 2}, lambda c: c.only_rabin_karp(True)),
}


def build_both(pats, kind, cfg):
    okw, setter = CONFIGS[cfg]
    o = O.PackedOracle(pats, kind=kind, **okw)
    c = setter(packed.Config().match_kind(packed.MatchKind.LeftmostFirst if kind == 0
                                          else packed.MatchKind.LeftmostLongest)).host_only()
    s = c.builder().extend(pats).build()
    return o, s


def agree(pats, kind, cfg):
    o, s = build_both(pats, kind, cfg)
    assert (s is not None) == o.built, (cfg, len(pats))
    if s is None:
        return None
    assert s.minimum_len() == o.minimum_len, (cfg, len(pats))
    assert s.patterns_len() == len(pats)
    assert s.match_kind() == (packed.MatchKind.LeftmostFirst if kind == 0 else packed.MatchKind.LeftmostLongest)
    v = s.variant()
    if CONFIGS[cfg][0].get("force") == 2:
        assert v is None and s.minimum_len() == 0
    else:
        assert v is not None and v["mask_len"] == min(4, min(len(p) for p in pats))
        assert s.minimum_len() == (32 if (v["vector_bytes"] == 32 and not v["fat"]) else 16) + v["mask_len"] - 1
    return s


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("coll,kind", [("PACKED_LEFTMOST_FIRST", 0), ("PACKED_LEFTMOST_LONGEST", 1)])
def test_build_decision_on_reference_vectors(coll, kind, cfg):
    for t in G.collection(PK, coll):
        agree(t["patterns"], kind, cfg)


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_build_decision_random_sets(cfg):
    rng = random.Random(hash(cfg) & 0xFFFF)
    for it in range(120):
        n = rng.choice([0, 1, 2, 7, 16, 17, 32, 33, 64, 65, 128, 129, 200])
        lo = rng.choice([0, 1, 1, 2, 3, 4, 5])
        pats = [bytes(rng.randrange(97, 123) for _ in range(rng.randint(lo, lo + 6))) for _ in range(n)]
        agree(pats, it % 2, cfg)


def test_none_cases_follow_builder_add():
    # api.rs:303-322: an empty pattern or a 129th pattern makes the builder inert; :254 no patterns
    assert packed.Config().host_only().builder().build() is None
    assert packed.Config().host_only().builder().extend([b"a", b"", b"b"]).build() is None
    many = [bytes([97 + i % 26, 97 + i // 26]) for i in range(129)]
    assert packed.Config().host_only().only_rabin_karp(True).builder().extend(many).build() is None
    assert packed.Config().host_only().only_rabin_karp(True).builder().extend(many[:128]).build() is not None
    # teddy/builder.rs:113-116, 166-171: heuristic limits
    assert packed.Config().host_only().builder().extend(many[:65]).build() is None
    assert packed.Config().host_only().heuristic_pattern_limits(False).builder().extend(many[:65]).build() is not None
    one_byte = [bytes([97 + i]) for i in range(17)]
    assert packed.Config().host_only().builder().extend(one_byte).build() is None
    assert packed.Config().host_only().builder().extend(one_byte[:16]).build() is not None


def test_default_variant_matches_prefilter_choice():
    # The Teddy that AhoCorasick would attach as a prefilter is the one packed::Builder builds by
    # default (src/util/prefilter.rs:296-303 -> packed::Config::new().builder()).
    from aho_corasick_b200 import workload as W
    pats = W.make_patterns(50, 0xAC0050)
    s = packed.Config().host_only().builder().extend(pats).build()
    ac = ab.AhoCorasick.builder().host_only().match_kind(ab.MatchKind.LeftmostFirst).build(pats)
    assert ac.packed_variant() == {"fat": True, "mask_len": 4}
    assert s.variant() == {"fat": True, "mask_len": 4, "vector_bytes": 32} and s.minimum_len() == 19


def test_host_only_searcher_validates_then_refuses():
    s = packed.Config().host_only().builder().extend([b"abc", b"bcd"]).build()
    with pytest.raises(ValueError):
        s.find_in(b"abcd", (3, 9))          # the reference's slice indexing panics
    with pytest.raises(ab.DeviceError):
        s.find_iter(b"xxabcdxx")            # no CPU search path in the product
    with pytest.raises(ab.DeviceError):
        s.find(b"xxabcdxx")
