#!/usr/bin/env python3
"""Randomized differential fuzzing of the product on the CPU dry-run library (tests/emu) against
the oracle.  usage: python tools/fuzz_dry_run.py [--minutes M] [--seed S] [--asan]
(with --asan start the interpreter with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0)."""
import argparse
import ctypes
import random
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "emu"):
    sys.path.insert(0, str(p))
import aho_corasick_b200 as ab  # noqa: E402
import build_emu  # noqa: E402
import oracle_py as O  # noqa: E402
from aho_corasick_b200 import packed, workload as W  # noqa: E402


def eq(got, want, ctx):
    ok = len(got) == len(want) and all(np.array_equal(got[k], want[k]) for k in ("pid", "start", "end"))
    if not ok:
        raise AssertionError(ctx)


def gen_case(rng):
    alphabet = rng.choice([b"ab", b"abc", b"abcd", b"aAbB", b"aAbBcC xyz", bytes(range(256)),
                           bytes(range(0x20, 0x7F)), b"abcdefghijklmnopqrstuvwxyz"])
    npat = rng.choice([1, 2, 3, 8, 40, 300, 3000, 12000])
    if len(alphabet) > 64:
        npat = min(npat, 300)   # keep the dense tables (states x 256 columns) and the runs small
    lo = rng.choice([1, 1, 2, 3, 4, 4, 5])
    hi = lo + rng.choice([0, 1, 3, 8, 16])
    pats = [bytes(rng.choice(alphabet) for _ in range(rng.randint(lo, hi))) for _ in range(npat)]
    if rng.random() < 0.15:
        pats.append(b"")
    if rng.random() < 0.3:
        pats += [pats[0], pats[len(pats) // 2]]
    size = rng.choice([0, 1, 2, 5, 16, 17, 100, 1000, 4096, 5000, 33000, 200000])
    style = rng.random()
    if style < 0.5:
        hay = np.frombuffer(bytes(rng.choice(alphabet) for _ in range(min(size, 4000))), dtype=np.uint8)
        hay = np.resize(hay, size) if size else hay[:0]
    elif style < 0.8:
        hay = np.frombuffer((b"".join(rng.choice(pats) or b"x" for _ in range(64)) * (size // 64 + 1))[:size], dtype=np.uint8)
    else:
        hay = np.empty(size, dtype=np.uint8)
        if size:
            W.fill_haystack(hay, rng.randrange(1 << 30))
    exact = np.empty(size, dtype=np.uint8)   # its own exact-size allocation
    exact[:] = hay
    return pats, exact


def one(rng, it):
    pats, hay = gen_case(rng)
    n = hay.size
    kind = rng.choice([0, 0, 1, 2])
    ci = rng.random() < 0.25
    start_kind = rng.choice([0, 0, 0, 2, 1])
    bc = rng.random() > 0.15
    b = (ab.AhoCorasick.builder().match_kind(kind).ascii_case_insensitive(ci).start_kind(start_kind)
         .byte_classes(bc).kind(ab.AhoCorasickKind.DFA))
    device_fill = rng.random() < 0.4   # dense table filled by dfa_fill_level_kernel (unanchored start kind)
    if device_fill:
        b.device_fill(True)
    ac = b.build(pats)
    # ACG_EXP_* kernel / plan variants (acb200_debug.h): KEY24 8, GLOBAL_TILES 16, STATIC_TILES 32, NO_BYTESCAN 64
    experiment = rng.choice([0, 0, 0] + [a | b | c for a in (0, 8) for b in (0, 16, 32) for c in (0, 64)])
    assert ab._lib.acg_debug_set_experiment(ac._h, experiment) == 0
    o = O.Oracle(pats, match_kind=kind, ascii_case_insensitive=ci, start_kind=start_kind, byte_classes=bc, kind=O.KIND_DFA)
    if rng.random() < 0.3:
        ab._lib.acg_debug_set_pipeline_chunk(ac._h, rng.choice([4096, 8192, 64 << 10]))
    lo = rng.randrange(0, n + 1)
    hi = rng.randrange(lo, n + 1)
    spans = [(0, n), (lo, hi)]
    ptr = hay.ctypes.data if n else 0
    ctx = (it, len(pats), n, kind, ci, start_kind, bc, device_fill, experiment)
    for span in spans:
        for anchored in ([False] if start_kind == 0 else [True] if start_kind == 1 else [False, True]):
            a = ab.Anchored.Yes if anchored else ab.Anchored.No
            want = o.find_iter_np(hay, span=span, anchored=anchored)
            eq(ac.try_find_iter_np(hay, span=span, anchored=a), want, ctx + (span, anchored, "find_iter host"))
            for earliest in (False, True):
                m = ac.try_find(hay, span=span, anchored=a, earliest=earliest)
                assert (m.as_tuple() if m else None) == o.try_find(hay, span=span, anchored=anchored, earliest=earliest), \
                    ctx + (span, anchored, earliest, "find")
        if start_kind != 1:
            want = o.find_iter_np(hay, span=span)
            for eng in (ab.Engine.Auto, ab.Engine.Sequential):
                ac.set_engine(eng)
                eq(ac.find_iter_dev_np(ptr, n, span=span)[0], want, ctx + (span, int(eng), "find_iter dev"))
            ac.set_engine(ab.Engine.Auto)
            # all occurrences: skip when the expected output is enormous (duplicate-heavy sets)
            expect = n * sum(len(set(hay[:4000].tolist()) or {0}) ** -float(len(p)) for p in pats if p) if n else 0
            if kind == 0 and expect < 2e6:
                want = o.find_overlapping_iter_np(hay, span=span)
                for eng in (ab.Engine.Auto, ab.Engine.Walk):
                    ac.set_engine(eng)
                    eq(ac.find_overlapping_iter_dev_np(ptr, n, span=span)[0], want, ctx + (span, int(eng), "overlapping dev"))
                ac.set_engine(ab.Engine.Auto)
                eq(ac.try_find_overlapping_iter_np(hay, span=span), want, ctx + (span, "overlapping host"))
                cnt, fnv, _ = ac.count_overlapping_dev(ptr, n, span=span)
                assert (cnt, fnv) == o.scan_overlapping_count(hay, span=span), ctx + (span, "count")
    if kind != 0 and start_kind == 0 and not ci and all(pats) and len(pats) <= 128:
        s = packed.Config().match_kind(kind).builder().extend(pats).build()
        po = O.PackedOracle(pats, kind=kind - 1)
        assert (s is not None) == po.built, ctx + ("packed built",)
        if s is not None:
            assert [m.as_tuple() for m in s.find_iter(hay)] == po.find_iter(hay), ctx + ("packed",)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=5)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--asan", action="store_true")
    args = ap.parse_args()
    lib = ctypes.CDLL(str(build_emu.build(asan=args.asan)))
    ab._declare(lib)
    packed._declare(lib)
    lib.acg_debug_set_pipeline_chunk.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    lib.acg_debug_set_experiment.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    ab._lib = packed._lib = lib
    rng = random.Random(args.seed)
    t0, it = time.time(), 0
    while time.time() - t0 < args.minutes * 60:
        t1 = time.time()
        one(rng, it)
        if time.time() - t1 > 20:
            print(f"slow case {it}: {time.time() - t1:.0f} s", flush=True)
        it += 1
        if it % 50 == 0:
            print(f"{it} cases, {time.time() - t0:.0f} s", flush=True)
    print(f"done: {it} cases without a discrepancy (seed {args.seed})")


if __name__ == "__main__":
    main()
