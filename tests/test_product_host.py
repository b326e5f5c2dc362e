"""CPU-side checks of the product (no GPU needed): the host builder's tables must be bit-identical
to the oracle's restatement of the reference build, the C-ABI library must load and export every
symbol include/acb200.h declares, and input validation must mirror the reference's errors."""
import ctypes as C
import random
import re
from pathlib import Path

import numpy as np
import pytest

import aho_corasick_b200 as ab
import golden_util as G
import oracle_py as O

ROOT = Path(__file__).resolve().parent.parent
AC = G.load("ac_vectors.json")


def assert_tables_equal(pt, ot, ctx):
    for k in ("stride2", "alphabet_len", "max_special_id", "max_match_id", "start_unanchored_id",
              "start_anchored_id", "match_kind", "min_pattern_len", "max_pattern_len", "state_len"):
        assert int(pt[k]) == int(ot[k]), (k, ctx)
    for k in ("trans", "byte_classes", "match_offsets", "match_pids", "pattern_lens"):
        assert np.array_equal(pt[k], ot[k]), (k, ctx)


def build_pair(pats, match_kind=0, **kw):
    okw = dict(kw)
    okw["kind"] = O.KIND_DFA
    o = O.Oracle(pats, match_kind=match_kind, **okw)
    b = ab.AhoCorasick.builder().host_only().match_kind(match_kind).kind(ab.AhoCorasickKind.DFA)
    for k, v in kw.items():
        getattr(b, k)(v)
    return b.build(pats), o


KNOBS = [{}, {"start_kind": 2}, {"start_kind": 1}, {"byte_classes": False},
         {"start_kind": 2, "byte_classes": False}, {"prefilter": False},
         {"ascii_case_insensitive": True}, {"ascii_case_insensitive": True, "start_kind": 2}]


@pytest.mark.parametrize("knob", range(len(KNOBS)))
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_table_parity_golden_patterns(kind, knob):
    seen = set()
    for g in AC["groups"].values():
        for t in g:
            key = tuple(t["patterns"])
            if key in seen:
                continue
            seen.add(key)
            p, o = build_pair(t["patterns"], match_kind=kind, **KNOBS[knob])
            assert_tables_equal(p.tables(), o.dfa(), (t["name"], kind, KNOBS[knob]))
            assert p.prefilter_kind() == o.prefilter_kind, t["name"]


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_table_parity_random(kind):
    rng = random.Random(0x7AB1E + kind)
    for it in range(150):
        alphabet = [b"ab", b"abcdefgh", bytes(range(256)), b"aAbBcC.-"][it % 4]
        npat = rng.choice([1, 2, 3, 8, 40, 130])
        pats = [bytes(rng.choice(alphabet) for _ in range(rng.randint(0 if it % 7 == 0 else 1, 9)))
                for _ in range(npat)]
        knob = dict(KNOBS[it % len(KNOBS)])
        p, o = build_pair(pats, match_kind=kind, **knob)
        assert_tables_equal(p.tables(), o.dfa(), (it, kind, knob))
        assert p.prefilter_kind() == o.prefilter_kind, (it, pats)
        pv, ov = p.packed_variant(), o.packed_variant()
        assert (pv is None) == (ov is None)
        if pv:
            assert pv["fat"] == ov["fat"] and pv["mask_len"] == ov["mask_len"]


def test_auto_kind_report():
    # build_auto, src/ahocorasick.rs:2213-2261
    assert ab.AhoCorasick.builder().host_only().build([b"a"]).kind() == ab.AhoCorasickKind.DFA
    many = [bytes([65 + i % 26, 65 + i // 26, 66]) for i in range(101)]
    assert ab.AhoCorasick.builder().host_only().build(many).kind() == ab.AhoCorasickKind.ContiguousNFA
    assert ab.AhoCorasick.builder().host_only().start_kind(ab.StartKind.Both).build([b"a"]).kind() == \
        ab.AhoCorasickKind.ContiguousNFA
    assert O.Oracle(many).kind == O.KIND_CONTIGUOUS


def test_library_exports_every_declared_symbol():
    names = set()
    for h in sorted((ROOT / "include").glob("*.h")):
        hdr = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)  # declarations only, not prose in comments
        names |= set(re.findall(r"\b(acg_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 26 and "acg_debug_prefilter_plan" in names
    lib = C.CDLL(str(ROOT / "aho-corasick_b200" / "libacb200.so"))
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_validation_errors_mirror_reference():
    # src/tests.rs:1407-1511 + src/automaton.rs:397-423 + src/ahocorasick.rs:2778-2789.
    # Validation happens before any device work, so it is observable on a host-only handle.
    def mk(**kw):
        b = ab.AhoCorasick.builder().host_only()
        for k, v in kw.items():
            getattr(b, k)(v)
        return b.build([b"a"])
    for k in (ab.MatchKind.LeftmostFirst, ab.MatchKind.LeftmostLongest):
        with pytest.raises(ab.MatchError) as e:
            mk(match_kind=k).try_find_overlapping_iter(b"a")
        assert e.value.kind == "UnsupportedOverlapping"
    with pytest.raises(ab.MatchError) as e:
        mk().try_find_iter(b"a", anchored=ab.Anchored.Yes)
    assert e.value.kind == "InvalidInputAnchored"
    with pytest.raises(ab.MatchError) as e:
        mk(start_kind=ab.StartKind.Anchored).try_find_iter(b"a")
    assert e.value.kind == "InvalidInputUnanchored"
    with pytest.raises(ab.MatchError) as e:
        mk(start_kind=ab.StartKind.Both).try_find_overlapping_iter(b"a", anchored=ab.Anchored.Yes)
    assert e.value.kind == "InvalidInputAnchored"
    with pytest.raises(ValueError):
        mk().try_find_iter(b"abc", span=(0, 4))
    with pytest.raises(ValueError):
        mk().try_find_iter(b"abc", span=(3, 1))
    # without a device the search itself must fail loudly -- there is no CPU fallback
    if ab.device_count() == 0:
        with pytest.raises(ab.DeviceError):
            mk().try_find_iter(b"abc")


def test_build_errors():
    # pattern longer than SmallIndex::MAX cannot be materialised here; the overflow guards are
    # exercised through the state-id limit with byte_classes(false): 2^31 / 256 rows.
    assert ab.BuildError(-1).code == -1
    t = ab.AhoCorasick.builder().host_only().build([]).tables()
    assert t["state_len"] == 4 and int(t["max_match_id"]) >> t["stride2"] == 1


def test_byte_class_corner_cases_from_the_reference_unit_tests():
    # src/util/alphabet.rs:337-408 (full_byte_classes, elements_singletons, elements_empty), reached
    # the way the automaton builders reach ByteClassSet: one set_range(b, b) per pattern byte
    # (src/nfa/noncontiguous.rs:1119-1123).  Product and oracle must agree with those expectations.
    def classes(pats, **kw):
        b = ab.AhoCorasick.builder().host_only().kind(ab.AhoCorasickKind.DFA)
        for k, v in kw.items():
            getattr(b, k)(v)
        t = b.build(pats).tables()
        o = O.Oracle(pats, kind=O.KIND_DFA, **kw).dfa()
        assert np.array_equal(t["byte_classes"], o["byte_classes"]) and t["alphabet_len"] == o["alphabet_len"]
        return t
    t = classes([])                                   # ByteClasses::empty(): one class for everything
    assert t["alphabet_len"] == 1 and not t["byte_classes"].any()
    t = classes([bytes(range(256))])                  # every byte its own class
    assert t["alphabet_len"] == 256 and np.array_equal(t["byte_classes"], np.arange(256, dtype=np.uint8))
    t = classes([b"a"], byte_classes=False)           # ByteClasses::singletons()
    assert t["alphabet_len"] == 256 and np.array_equal(t["byte_classes"], np.arange(256, dtype=np.uint8))
    t = classes([b"bd", b"z"])                        # classes: \\x00-a | b | c | d | e-y | z | {-\\xff
    bc = t["byte_classes"]
    assert t["alphabet_len"] == 7
    assert (bc[0], bc[ord("a")], bc[ord("b")], bc[ord("c")], bc[ord("d")], bc[ord("e")], bc[ord("y")],
            bc[ord("z")], bc[ord("{")], bc[255]) == (0, 0, 1, 2, 3, 4, 4, 5, 6, 6)
