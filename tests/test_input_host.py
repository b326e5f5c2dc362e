"""CPU checks of the `Input` mirror (src/util/search.rs:60-720) and of how searches consume it."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import aho_corasick_b200 as ab  # noqa: E402


def test_input_builder_and_getters():
    i = ab.Input(b"foobar")
    assert i.get_span() == (0, 6) and i.get_anchored() == ab.Anchored.No and not i.get_earliest()
    assert not i.is_done() and bytes(i.haystack()) == b"foobar"
    i = ab.Input.new("foobar").span((2, 4)).anchored(ab.Anchored.Yes).earliest(True)
    assert (i.start(), i.end(), i.get_anchored(), i.get_earliest()) == (2, 4, ab.Anchored.Yes, True)
    assert i.get_range() == range(2, 4)
    assert ab.Input(b"foobar").range(range(1, 3)).get_span() == (1, 3)
    assert ab.Input(b"foobar").range(slice(None, 3)).get_span() == (0, 3)
    assert ab.Input(b"foobar").range((4, 6)).get_span() == (4, 6)
    j = i.clone()
    j.set_start(3)
    j.set_end(3)
    assert j.get_span() == (3, 3) and i.get_span() == (2, 4)
    arr = np.frombuffer(b"xyz", dtype=np.uint8)
    assert ab.Input(arr).haystack() is arr


def test_input_span_rules():
    # src/util/search.rs:332-343: valid iff end <= len and start <= end + 1 (the reference panics)
    i = ab.Input(b"foobar")
    i.set_span((6, 6))
    i.set_span((4, 3))           # start == end + 1: a "done" input, allowed
    assert i.is_done()
    for bad in ((0, 7), (5, 3), (8, 6)):
        with pytest.raises(ValueError):
            ab.Input(b"foobar").span(bad)
    with pytest.raises(ValueError):
        ab.Input(b"foobar").set_end(7)


def test_searches_consume_an_input():
    # validation happens before any device work: observable on host-only handles
    un = ab.AhoCorasick.builder().host_only().build([b"a"])
    with pytest.raises(ab.MatchError) as e:
        un.try_find_iter(ab.Input(b"a").anchored(ab.Anchored.Yes))
    assert e.value.kind == "InvalidInputAnchored"
    with pytest.raises(ab.MatchError) as e:
        un.try_find(ab.Input(b"a").anchored(ab.Anchored.Yes))
    assert e.value.kind == "InvalidInputAnchored"
    anch = ab.AhoCorasick.builder().host_only().start_kind(ab.StartKind.Anchored).build([b"a"])
    with pytest.raises(ab.MatchError) as e:
        anch.try_find_overlapping_iter(ab.Input(b"a"))
    assert e.value.kind == "InvalidInputUnanchored"
    with pytest.raises(ab.MatchError) as e:
        anch.is_match(ab.Input(b"a"))
    assert e.value.kind == "InvalidInputUnanchored"
    # a valid input reaches the device layer, which a host-only handle does not have
    with pytest.raises(ab.DeviceError):
        un.try_find_iter(ab.Input(b"banana").span((1, 5)))
    with pytest.raises(ab.DeviceError):
        un.try_find_overlapping(ab.Input(b"banana"), ab.OverlappingState.start())
