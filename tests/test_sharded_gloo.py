"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: slicing with
max_pattern_len-1 overlap, ownership by end offset, gather to rank 0.  The per-slice match lists
come from the oracle (this is a test of the sharding logic, not of the kernels)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_py as O
        from aho_corasick_b200 import sharded as S
        pats, hay, span = case
        o = O.Oracle(pats, kind=O.KIND_DFA)
        s0, s1 = span
        lo, hi, read_lo = S.slice_plan(s0, s1, world, o.max_pattern_len)[rank]
        local = o.find_overlapping_iter_np(hay, span=(read_lo, hi))
        local = S.owned(local, rank, lo)
        full = S.gather_to_rank0(local, dist)
        if rank == 0:
            want = o.find_overlapping_iter_np(hay, span=span)
            ok = len(full) == len(want) and all(np.array_equal(full[k], want[k]) for k in ("pid", "start", "end"))
            q.put((ok, len(full), len(want)))
    finally:
        dist.destroy_process_group()


def _cases():
    rng = np.random.default_rng(7)
    hay = rng.integers(97, 101, size=20000, dtype=np.uint8)
    pats = [b"ab", b"abc", b"bcda", b"a", b"ddddddd", b"cab", b"ab"]
    yield pats, hay, (0, hay.size)
    yield pats, hay, (123, 19001)
    yield [b"", b"ab", b"b"], hay[:3000].copy(), (5, 2999)  # empty pattern: every position matches
    yield [b"abcdabcdabcdabcdabcd"], np.tile(np.frombuffer(b"abcd", dtype=np.uint8), 2000), (0, 8000)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_overlapping_gloo(world):
    ctx = mp.get_context("spawn")
    for case in _cases():
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(600)
            assert p.exitcode == 0
        ok, n, want = q.get(timeout=60)
        assert ok, (n, want, case[2], world)


def test_slice_plan_properties():
    from aho_corasick_b200 import sharded as S
    for world in (1, 2, 4, 8):
        plan = S.slice_plan(100, 100 + (1 << 20) + 37, world, 16)
        assert plan[0][0] == 100 and plan[-1][1] == 100 + (1 << 20) + 37
        for g in range(world):
            lo, hi, read_lo = plan[g]
            assert lo <= hi and read_lo == max(100, lo - 15)
            if g:
                assert plan[g - 1][1] == lo
