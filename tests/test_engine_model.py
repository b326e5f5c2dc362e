"""CPU model of the prefilter engine's *algorithm* (not of the CUDA code): every start offset is
tested independently -- Bloom probes, anchor-map lookup, then a walk of the shipped DFA while the
state stays on the trie path anchored at that offset (depth == bytes consumed), reporting the
node's own patterns -- and the tuples are ordered by the packed key and, for find_iter, resolved
by the greedy chain.  The model runs on the very tables the library derives for the device
(include/acb200_debug.h) and must reproduce the oracle's results, order included.  This pins the
exactness argument of DESIGN.md section 3 (per-start verification, tie-break layout, "last own
match" for leftmost kinds, chain resolution) without a GPU; the kernels are checked against the
same oracle by the `-m gpu` suite."""
import random
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import aho_corasick_b200 as ab  # noqa: E402
import oracle_py as O  # noqa: E402
from test_prefilter_plan import (anchor_lookup, first_stage_hit, le32, plan_of, second_stage_hit)  # noqa: E402

TIE_BITS = 24


class Model:
    def __init__(self, pats, match_kind=0, ci=False):
        self.ac = (ab.AhoCorasick.builder().host_only().kind(ab.AhoCorasickKind.DFA).match_kind(match_kind)
                   .ascii_case_insensitive(ci).build(pats))
        self.p = plan_of(self.ac)
        self.t = self.ac.tables()
        self.mode = 0 if match_kind == 0 else 1
        self.s2 = int(self.t["stride2"])

    def candidates(self, hay, lo, hi):
        """Start offsets that survive the fingerprint stages (a superset of pattern beginnings)."""
        p = self.p
        pad = bytes(hay) + b"\0" * 8
        out = []
        for s in range(lo, hi):
            if p.brute:
                out.append(s)
                continue
            w = le32(pad[s:s + 4])
            if p.stride == 2:
                # the probe at the even offset covering s: s itself, or s + 1 with the window shifted by a byte
                hit = first_stage_hit(p, w) if s % 2 == 0 else first_stage_hit(p, le32(pad[s + 1:s + 5]))
            else:
                hit = first_stage_hit(p, w)
            if hit and second_stage_hit(p, w):
                out.append(s)
        return out

    def verify(self, hay, s, span_start, span_end, emit):
        """verify_at / verify_from of csrc/acb_prefilter.cu."""
        p, t, s2 = self.p, self.t, self.s2
        trans, cls = t["trans"], t["byte_classes"]
        if p.amap_log:
            if s + p.k > span_end:
                return
            sid = anchor_lookup(p, le32(hay[s:s + 4]) & p.kmask)
            if sid == 0:
                return
            j, entered = p.k, True
        else:
            sid, j, entered = int(t["start_unanchored_id"]), 0, False
        pos = s + j
        best = None
        max_len = int(t["max_pattern_len"])
        while True:
            if not entered:
                if pos >= span_end:
                    break
                sid = int(trans[sid + int(cls[hay[pos]])])
                j += 1
                pos += 1
                if sid == 0 or p.depth16[sid >> s2] != j:
                    break
            entered = False
            if sid <= int(t["max_match_id"]):
                row = sid >> s2
                lo, hi = int(t["match_offsets"][row - 2]), int(t["match_offsets"][row - 1])
                if self.mode == 0:
                    for i in range(lo, hi):
                        pid = int(t["match_pids"][i])
                        if int(t["pattern_lens"][pid]) != j:
                            break
                        tie = ((max_len - j) << p.dup_shift) | (i - lo)
                        emit((((pos - span_start) << TIE_BITS) | tie, pid))
                else:
                    pid = int(t["match_pids"][lo])
                    if int(t["pattern_lens"][pid]) == j:
                        best = (pid, j)
        if self.mode == 1 and best:
            emit((((s - span_start) << TIE_BITS) | best[1], best[0]))

    def tuples(self, hay, span):
        lo, hi = span
        out = []
        for s in self.candidates(hay, lo, hi):
            self.verify(hay, s, lo, hi, out.append)
        out.sort()
        lens = self.t["pattern_lens"]
        res = []
        for key, pid in out:
            if self.mode == 1:
                start = lo + (key >> TIE_BITS)
                res.append((pid, start, start + (key & ((1 << TIE_BITS) - 1))))
            else:
                end = lo + (key >> TIE_BITS)
                res.append((pid, end - int(lens[pid]), end))
        return res

    def find_overlapping(self, hay, span):
        return self.tuples(hay, span)

    def find_iter(self, hay, span):
        """chain_select_kernel: greedy FindIter over the ordered tuples (src/automaton.rs:927-935)."""
        cur, out = span[0], []
        for pid, s, e in self.tuples(hay, span):
            if s >= cur:
                out.append((pid, s, e))
                cur = e
        return out


def rand_case(rng, it):
    alphabet = [b"ab", b"abcd", b"aAbBcC ", b"abcdefghij"][it % 4]
    n = rng.choice([1, 2, 5, 20, 120])
    pats = [bytes(rng.choice(alphabet) for _ in range(rng.randint(1, rng.choice([3, 6, 12])))) for _ in range(n)]
    if it % 3 == 0:
        pats += [pats[0], pats[-1]]  # duplicates
    hay = bytes(rng.choice(alphabet) for _ in range(rng.choice([0, 1, 5, 60, 700, 2500])))
    lo = rng.randrange(0, len(hay) + 1) if it % 5 == 0 else 0
    hi = rng.randrange(lo, len(hay) + 1) if it % 5 == 0 else len(hay)
    return pats, hay, (lo, hi)


@pytest.mark.parametrize("ci", [False, True])
def test_model_overlapping_equals_oracle(ci):
    rng = random.Random(101 + ci)
    for it in range(60):
        pats, hay, span = rand_case(rng, it)
        m = Model(pats, 0, ci)
        if not m.p.supported:
            continue
        want = O.Oracle(pats, ascii_case_insensitive=ci, kind=O.KIND_DFA).find_overlapping_iter(hay, span=span)
        assert m.find_overlapping(hay, span) == want, (it, pats[:4], len(hay), span)


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("ci", [False, True])
def test_model_find_iter_equals_oracle(kind, ci):
    rng = random.Random(7 + kind * 2 + ci)
    for it in range(50):
        pats, hay, span = rand_case(rng, it)
        m = Model(pats, kind, ci)
        if not m.p.supported:
            continue
        want = O.Oracle(pats, match_kind=kind, ascii_case_insensitive=ci, kind=O.KIND_DFA).find_iter(hay, span=span)
        assert m.find_iter(hay, span) == want, (it, kind, pats[:4], len(hay), span)


def test_model_on_workload_like_sets():
    from aho_corasick_b200 import workload as W
    for n, seed, kind, ci in [(5000, 0xAC5000, 0, False), (5000, 0xAC5000, 1, True), (50, 0xAC0050, 1, False)]:
        pats = W.make_patterns(n, seed)
        hay = np.empty(24 << 10, dtype=np.uint8)
        W.fill_haystack(hay, 5)
        W.plant(hay, pats, 6, period=256, window=128)
        if ci:
            W.flip_case(hay, 7)
        hb = hay.tobytes()
        m = Model(pats, kind, ci)
        assert m.p.supported and m.p.stride == 2
        o = O.Oracle(pats, match_kind=kind, ascii_case_insensitive=ci, kind=O.KIND_DFA)
        if kind == 0:
            want = o.find_overlapping_iter(hb)
            assert len(want) > 80 and m.find_overlapping(hb, (0, len(hb))) == want
        assert m.find_iter(hb, (0, len(hb))) == o.find_iter(hb)
