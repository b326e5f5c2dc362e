"""acg_find_overlapping_sharded (include/acb200.h) on the CPU dry run: the
"ranks" are threads of this process, each with its own automaton handle and communicator, the
library is the g++ build of the product sources (tests/emu) whose fabric replaces NCCL / cudaIpc
with in-process rendezvous.  Checks the slice plan, ownership by end offset at slice boundaries
that cut planted matches, the offsets of the global list, growth of rank 0's receive buffer, the
host-buffer (pipelined) input path, sub-spans, and more ranks than 64-byte blocks -- against the
oracle over the whole haystack.  The NCCL / peer-memory transport itself is covered by
tests/test_gpu_zz_multirank.py on hardware."""
import ctypes
import sys
import threading
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "emu"))
import aho_corasick_b200 as ab  # noqa: E402
import oracle_py as O  # noqa: E402
from aho_corasick_b200 import packed, sharded as S, workload as W  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def emulated_library():
    import build_emu
    lib = ctypes.CDLL(str(build_emu.build()))
    ab._declare(lib)
    packed._declare(lib)
    saved = ab._lib, packed._lib
    ab._lib = packed._lib = lib
    try:
        yield lib
    finally:
        ab._lib, packed._lib = saved


def run_ranks(world, pats, hay, span, on_device=True, engine=ab.Engine.Auto, whole_buffer=False):
    """One thread per rank; returns (rank-0 matches, per-rank stats)."""
    uid = S.unique_id()
    res, errs = [None] * world, []
    maxlen = max((len(p) for p in pats), default=0)
    plan = S.slice_plan(span[0], span[1], world, maxlen)

    def work(rank):
        try:
            ac = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).build(pats).set_engine(engine)
            comm = S.Comm(uid, rank, world)
            lo, hi, rd = plan[rank]
            if whole_buffer:   # the rank holds the whole haystack
                g0, local = 0, hay
            else:              # the rank holds exactly its slice (16-byte phase of the global offsets kept)
                g0 = rd - rd % 16
                local = np.ascontiguousarray(hay[g0:hi])
            n, dptr, st, out = comm.find_overlapping(ac, local.ctypes.data, local.size, g0, span,
                                                     on_device=on_device, host_out=True)
            chk = comm.checksum() if rank == 0 else None
            if out is not None:
                assert np.array_equal(out, comm.fetch())   # the page-locked view and the copying fetch agree
                out = out.copy()                           # (the view dies with the communicator)
            res[rank] = (n, out, st, chk)
            comm.close()
        except Exception as e:  # noqa: BLE001
            errs.append((rank, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=600) for t in ts]
    assert not errs, errs
    assert all(r is not None for r in res)
    return res


def eq(got, want):
    assert len(got) == len(want), (len(got), len(want))
    for k in ("pid", "start", "end"):
        assert np.array_equal(got[k], want[k]), k


def fnv(m):
    h = 0xcbf29ce484222325
    for pid, s, e in zip(m["pid"].tolist(), m["start"].tolist(), m["end"].tolist()):
        for v in (pid, s, e):
            for k in range(8):
                h ^= (v >> (8 * k)) & 0xFF
                h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_sharded_equals_single_list(world):
    pats = W.make_patterns(400, 77)
    hay = np.empty(96 << 10, dtype=np.uint8)
    W.fill_haystack(hay, 5)
    W.plant(hay, pats, 6, period=512, window=256)
    # make sure a match straddles every interior boundary: plant one across it
    for lo, hi, rd in S.slice_plan(0, hay.size, world, 16)[1:]:
        p = pats[3]
        hay[lo - len(p) // 2: lo - len(p) // 2 + len(p)] = np.frombuffer(p, np.uint8)
        q = pats[5]
        hay[lo - len(q) + 1: lo + 1] = np.frombuffer(q, np.uint8)   # ends exactly at lo + 1 (first owned end)
        r = pats[7]
        hay[lo - len(r): lo] = np.frombuffer(r, np.uint8)           # ends exactly at lo (previous rank's)
    want = O.Oracle(pats, kind=O.KIND_DFA).find_overlapping_iter_np(hay)
    res = run_ranks(world, pats, hay, (0, hay.size))
    n, out, st, chk = res[0]
    assert n == len(want) and all(r[0] == n for r in res)
    eq(out, want)
    assert chk == (len(want), fnv(want))
    assert sum(r[2]["local_matches"] for r in res) == n
    assert all(r[2]["transport"] == 1 for r in res)


def test_sharded_host_input_subspan_and_walk_engine():
    pats = W.make_patterns(200, 3) + [b"ab", b"b"]
    hay = np.empty(40 << 10, dtype=np.uint8)
    W.fill_haystack(hay, 9, alphabet=(0x61, 0x64))
    o = O.Oracle(pats, kind=O.KIND_DFA)
    span = (1003, 39000)
    want = o.find_overlapping_iter_np(hay, span=span)
    assert len(want) > 500
    for kw in (dict(on_device=False), dict(engine=ab.Engine.Walk), dict(whole_buffer=True)):
        res = run_ranks(3, pats, hay, span, **kw)
        eq(res[0][1], want)


def test_more_ranks_than_blocks_and_empty_results():
    pats = [b"needle", b"hay"]
    hay = np.frombuffer(b"a haystack with a needle in the hay, and hay again........" * 2, dtype=np.uint8).copy()
    want = O.Oracle(pats, kind=O.KIND_DFA).find_overlapping_iter_np(hay)
    res = run_ranks(4, pats, hay, (0, hay.size))
    eq(res[0][1], want)
    res = run_ranks(2, [b"zzzzzz"], hay, (0, hay.size))
    assert res[0][0] == 0 and len(res[0][1]) == 0


def test_receive_buffer_grows():
    """More matches than the initial capacity of rank 0's receive buffer (65 536 records)."""
    pats = [b"a", b"aa", b"aaa"]
    hay = np.full(40000, ord("a"), dtype=np.uint8)
    want = O.Oracle(pats, kind=O.KIND_DFA).find_overlapping_iter_np(hay)
    assert len(want) > 100000
    res = run_ranks(2, pats, hay, (0, hay.size))
    eq(res[0][1], want)


def test_slice_does_not_cover_plan_is_an_error_on_every_rank():
    pats = [b"abc"]
    hay = np.zeros(4096, dtype=np.uint8)
    uid = S.unique_id()
    out = [None, None]

    def work(rank):
        ac = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).build(pats)
        comm = S.Comm(uid, rank, 2)
        try:
            # rank 1 hands over too few bytes; rank 0 is fine -- both must return an error, nobody hangs
            ln = hay.size if rank == 0 else 100
            comm.find_overlapping(ac, hay.ctypes.data, ln, 0, (0, hay.size))
            out[rank] = "ok"
        except Exception as e:  # noqa: BLE001
            out[rank] = type(e).__name__
        comm.close()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert out[1] == "ValueError" and out[0] in ("DeviceError", "ValueError"), out


def test_pipelined_steps_two_in_flight():
    """acg_find_overlapping_sharded_begin / _wait over a stream of haystack batches: step k + 1 begins
    before step k is waited for (two halves of rank 0's buffer, one leased workspace per step); every
    step's gathered list is the oracle's for its batch; a third begin without a wait is refused."""
    world = 3
    pats = W.make_patterns(300, 5)
    o = O.Oracle(pats, kind=O.KIND_DFA)
    batches = []
    for b in range(5):
        h = np.empty((24 << 10) + 4096 * b, dtype=np.uint8)
        W.fill_haystack(h, 100 + b)
        W.plant(h, pats, 200 + b, period=256, window=128)
        batches.append(h)
    wants = [o.find_overlapping_iter_np(h) for h in batches]
    uid = S.unique_id()
    got, errs = [None] * len(batches), []

    def work(rank):
        try:
            ac = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).build(pats)
            comm = S.Comm(uid, rank, world)
            # the buffer must not have to grow with steps in flight: one blocking step sizes it
            h = batches[-1]
            comm.find_overlapping(ac, h.ctypes.data, h.size, 0, (0, h.size))
            tickets = []
            comm.mark(0)   # device timestamps around the loop (acg_comm_mark), as bench.py takes them
            for k, h in enumerate(batches):
                tickets.append(comm.begin(ac, h.ctypes.data, h.size, 0, (0, h.size)))
                if k >= 1:
                    n, dptr, st = comm.wait(tickets[k - 1])
                    if rank == 0:
                        got[k - 1] = comm.fetch()
                        assert n == len(got[k - 1])
            n, dptr, st = comm.wait(tickets[-1])
            comm.mark(1)
            assert comm.mark_elapsed_ms() >= 0.0
            if rank == 0:
                got[-1] = comm.fetch()
            comm.close()
        except Exception as e:  # noqa: BLE001
            errs.append((rank, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=600) for t in ts]
    assert not errs, errs
    for g, w in zip(got, wants):
        eq(g, w)
    # protocol errors: a third step without a wait, waiting twice
    ac = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).build(pats)
    comm = S.Comm(S.unique_id(), 0, 1)
    h = batches[0]
    comm.find_overlapping(ac, h.ctypes.data, h.size, 0, (0, h.size))
    t0 = comm.begin(ac, h.ctypes.data, h.size, 0, (0, h.size))
    t1 = comm.begin(ac, h.ctypes.data, h.size, 0, (0, h.size))
    with pytest.raises(ab.DeviceError):
        comm.begin(ac, h.ctypes.data, h.size, 0, (0, h.size))
    assert comm.wait(t0)[0] == len(wants[0]) and comm.wait(t1)[0] == len(wants[0])
    with pytest.raises(ab.DeviceError):
        comm.wait(t1)
    with pytest.raises(ab.DeviceError):
        S.Comm(S.unique_id(), 0, 1).mark_elapsed_ms()   # no marks recorded yet
    comm.close()
