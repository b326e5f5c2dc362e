"""One rank of the multi-GPU parity check (tests/test_gpu_zz_multirank.py): sharded search through the
C ABI (acg_comm_init + acg_find_overlapping_sharded) on this rank's slice of a global synthetic
haystack; rank 0 compares count + FNV of the gathered stream with acg_count_overlapping_dev over the
whole haystack on its own GPU, and a window of the gathered records with the oracle.

    python tests/multirank_worker.py <rank> <world> <uid_file> <workload> <total_bytes> [host]
"""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def plant_boundary_matches(t, g0, pats, plan, shift=0):
    """A match cut by every interior slice boundary b, of a kind that rotates with the boundary:
    0 -- straddles b (seen by the next rank only through its cold-start overlap);
    1 -- ends exactly at b (owned by the previous rank, must not be reported twice);
    2 -- ends at b + 1 (the first end the next rank owns).
    Every rank applies the same edits to the part of them its slice holds.  Returns the expected
    (kind, boundary, end offset) list."""
    n = t.numel()
    expect = []
    for j, (lo, hi, rd) in enumerate(plan[1:], start=1):
        kind = (j + shift) % 3
        p = pats[3 + j]
        end = lo + len(p) // 2 if kind == 0 else (lo if kind == 1 else lo + 1)
        if kind == 0 and end <= lo:
            end = lo + 1
        for i, byte in enumerate(p):
            k = end - len(p) + i - g0
            if 0 <= k < n:
                t[k] = byte
        expect.append((kind, lo, end))
    return expect


def main():
    rank, world, uid_file, workload, total = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5])
    host_input = len(sys.argv) > 6 and sys.argv[6] == "host"
    shift = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    import numpy as np
    import torch
    import aho_corasick_b200 as ab
    from aho_corasick_b200 import sharded as S, workload as W
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    cfg = W.CONFIGS[workload]
    pats = W.make_patterns(cfg["n_patterns"], cfg["pattern_seed"], alphabet=cfg["alphabet"])
    ac = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).build(pats)
    if rank == 0:
        Path(uid_file + ".tmp").write_bytes(S.unique_id())
        os.replace(uid_file + ".tmp", uid_file)
    t0 = time.time()
    while not Path(uid_file).exists():
        if time.time() - t0 > 120:
            raise SystemExit("no rendezvous token")
        time.sleep(0.05)
    uid = Path(uid_file).read_bytes()
    comm = S.Comm(uid, rank, world)
    span = (0, total)
    plan = S.slice_plan(span[0], span[1], world, ac.max_pattern_len())
    lo, hi, rd = plan[rank]
    g0 = rd - rd % 4096
    n_local = hi - g0
    n_local += (-n_local) % 8
    d_hay = torch.empty(n_local, dtype=torch.uint8, device=dev)
    W.torch_fill_config(workload, d_hay, pats, global_offset=g0)
    plant_boundary_matches(d_hay, g0, pats, plan, shift)
    torch.cuda.synchronize()
    if host_input:
        h = d_hay.cpu().numpy()
        n, dptr, st, out = comm.find_overlapping(ac, h.ctypes.data, min(h.size, hi - g0), g0, span, on_device=False)
    else:
        n, dptr, st, out = comm.find_overlapping(ac, d_hay.data_ptr(), min(n_local, hi - g0), g0, span)
    # a second call on the same communicator (buffers reused, counts differ per rank) must agree
    n2, _, st2, _ = comm.find_overlapping(ac, d_hay.data_ptr(), min(n_local, hi - g0), g0, span)
    assert n2 == n, (n, n2)
    blocking_chk = comm.checksum() if rank == 0 else None
    # pipelined form: three steps, two in flight (the scan of step k + 1 overlaps the transfer of step k)
    args_b = (ac, d_hay.data_ptr(), min(n_local, hi - g0), g0, span)
    t0 = comm.begin(*args_b)
    t1 = comm.begin(*args_b)
    r0 = comm.wait(t0)
    c0 = comm.checksum() if rank == 0 else None
    t2 = comm.begin(*args_b)
    r1 = comm.wait(t1)
    c1 = comm.checksum() if rank == 0 else None
    r2 = comm.wait(t2)
    c2 = comm.checksum() if rank == 0 else None
    assert r0[0] == r1[0] == r2[0] == n, (r0[0], r1[0], r2[0], n)
    if rank == 0:
        assert c0 == c1 == c2 == blocking_chk, (c0, c1, c2, blocking_chk)
    pipe_gather_ms = r2[2]["gather_ms"]
    if rank == 0:
        got_n, got_fnv = comm.checksum()
        assert got_n == n
        del d_hay
        whole = torch.empty(total, dtype=torch.uint8, device=dev)
        W.torch_fill_config(workload, whole, pats, global_offset=0)
        expect = plant_boundary_matches(whole, 0, pats, plan, shift)
        torch.cuda.synchronize()
        want_n, want_fnv, _ = ac.count_overlapping_dev(whole.data_ptr(), total)
        assert (got_n, got_fnv) == (want_n, want_fnv), ((got_n, got_fnv), (want_n, want_fnv))
        # the records around the first interior boundary against the oracle
        import oracle_py as O
        rec = comm.fetch()
        assert len(rec) == n and bool(np.all(np.diff(rec["end"].astype(np.int64)) >= 0))
        if world > 1:
            b = plan[1][0]
            w0, w1 = max(0, b - (2 << 20)), min(total, b + (2 << 20))
            win = whole[w0:w1].cpu().numpy()
            o = O.Oracle(pats, kind=O.KIND_DFA)
            want = o.find_overlapping_iter_np(win)
            sel = rec[(rec["start"] >= w0) & (rec["end"] <= w1)]
            assert len(sel) == len(want) and np.array_equal(sel["end"] - w0, want["end"]) \
                and np.array_equal(sel["start"] - w0, want["start"]) and np.array_equal(sel["pid"], want["pid"])
        # every planted boundary match is in the stream exactly once
        for kind, b, end in expect:
            hits = rec[rec["end"] == end]
            assert len(hits) >= 1 and len(np.unique(hits[["pid", "start", "end"]])) == len(hits), (kind, b, end)
        print(f"MULTIRANK OK world={world} workload={workload} total={total} matches={n} transport={comm.transport()} "
              f"scan_ms={st['scan_ms']:.3f} gather_ms={st['gather_ms']:.3f} gather_ms_2nd={st2['gather_ms']:.3f} pipelined_ok gather_ms_in_pipeline={pipe_gather_ms:.3f}", flush=True)
    comm.close()


if __name__ == "__main__":
    main()
