"""Loader for the committed golden fixtures + the reference's test matrix (src/tests.rs:653-1323)."""
import json
from pathlib import Path

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    d = json.loads((GOLDEN / name).read_text())
    for g in d["groups"].values():
        for t in g:
            t["patterns"] = [bytes.fromhex(p) for p in t["patterns"]]
            t["haystack"] = bytes.fromhex(t["haystack"])
            t["matches"] = [tuple(m) for m in t["matches"]]
    return d


def collection(d, name):
    out = []
    for g in d["collections"][name]:
        out.extend(d["groups"][g])
    return out


STANDARD, LEFTMOST_FIRST, LEFTMOST_LONGEST = 0, 1, 2
KIND_AUTO, KIND_NFA, KIND_CONTIGUOUS, KIND_DFA = 0, 1, 2, 3
START_UNANCHORED, START_ANCHORED, START_BOTH = 0, 1, 2

# The `testcombo!` knob matrix of src/tests.rs:723-863, minus the contiguous-NFA rows
# (out of scope).  Values are builder kwargs.
COMBO = {
    "default": {},
    "nfa_default": {"kind": KIND_NFA},
    "nfa_noncontig_no_prefilter": {"kind": KIND_NFA, "prefilter": False},
    "nfa_noncontig_all_sparse": {"kind": KIND_NFA, "dense_depth": 0},
    "nfa_noncontig_all_dense": {"kind": KIND_NFA, "dense_depth": -1},
    "dfa_default": {"kind": KIND_DFA},
    "dfa_start_both": {"kind": KIND_DFA, "start_kind": START_BOTH},
    "dfa_no_prefilter": {"kind": KIND_DFA, "prefilter": False},
    "dfa_start_both_no_prefilter": {"kind": KIND_DFA, "start_kind": START_BOTH, "prefilter": False},
    "dfa_no_byte_class": {"kind": KIND_DFA, "byte_classes": False},
    "dfa_start_both_no_byte_class": {"kind": KIND_DFA, "start_kind": START_BOTH, "byte_classes": False},
}
# DFA-only rows: what the device path must cover.
COMBO_DFA = {k: v for k, v in COMBO.items() if k.startswith("dfa")}

NON_OVERLAPPING_COLLECTIONS = [
    ("AC_LEFTMOST_LONGEST", LEFTMOST_LONGEST),
    ("AC_LEFTMOST_FIRST", LEFTMOST_FIRST),
    ("AC_STANDARD_NON_OVERLAPPING", STANDARD),
]

# anchored rows, src/tests.rs:1039-1179
ANCHORED = [
    ("AC_STANDARD_ANCHORED_NON_OVERLAPPING", STANDARD),
    ("AC_LEFTMOST_FIRST_ANCHORED", LEFTMOST_FIRST),
    ("AC_LEFTMOST_LONGEST_ANCHORED", LEFTMOST_LONGEST),
]
ANCHORED_COMBO = {
    "default": {"start_kind": START_ANCHORED},
    "nfa_noncontig_default": {"start_kind": START_ANCHORED, "kind": KIND_NFA},
    "dfa_default": {"start_kind": START_ANCHORED, "kind": KIND_DFA},
    "dfa_start_both": {"start_kind": START_BOTH, "kind": KIND_DFA},
}
