"""Haystack slicing across GPUs and the gather of match buffers to rank 0.

The path shards naturally: rank g owns the matches whose END lies in (a_g, b_g] (rank 0 also owns
end == span start, i.e. empty-pattern matches of the start state, src/automaton.rs:1456-1464), reads
max_pattern_len-1 bytes before a_g so that every pattern ending in its range is seen from a cold
start, and never exchanges haystack data with another rank.  The only exchange is the gather of
the per-rank match buffers to rank 0; because the slices are ordered and every end offset has
exactly one owner, concatenating the per-rank buffers in rank order reproduces the single-GPU
iteration order.

The product path is behind the C ABI (include/acb200.h: acg_comm_init, acg_shard_plan,
acg_find_overlapping_sharded): `Comm` below is its ctypes mirror.  `gather_to_rank0` is the same
gather written with torch.distributed for host-side arrays (gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MATCH_DTYPE = np.dtype([("pid", "<u4"), ("_pad", "<u4"), ("start", "<u8"), ("end", "<u8")])
COMM_ID_BYTES = 128
TRANSPORT = {0: "none", 1: "peer", 2: "nccl"}


def _ab():
    import aho_corasick_b200 as ab
    return ab


class ShardStats(C.Structure):
    _fields_ = [("local_matches", C.c_uint64), ("total_matches", C.c_uint64), ("candidates", C.c_uint64),
                ("scan_ms", C.c_float), ("order_ms", C.c_float), ("gather_ms", C.c_float),
                ("transport", C.c_int32), ("launches", C.c_int32)]


def unique_id() -> bytes:
    """Rendezvous token (ncclUniqueId) created by rank 0; hand it to the other ranks."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = _ab()._lib.acg_comm_unique_id(buf)
    if rc:
        raise _ab().DeviceError(rc)
    return bytes(buf)


class Comm:
    """acg_comm: one per rank, bound to the current CUDA device (collective constructor)."""

    def __init__(self, uid: bytes, rank: int, nranks: int):
        assert len(uid) == COMM_ID_BYTES
        self._lib = _ab()._lib
        h = C.c_void_p()
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(uid)
        rc = self._lib.acg_comm_init(buf, rank, nranks, C.byref(h))
        if rc:
            raise _ab().DeviceError(rc)
        self._h, self.rank, self.nranks = h, rank, nranks

    def close(self):
        if getattr(self, "_h", None):
            self._lib.acg_comm_free(self._h)
            self._h = None

    __del__ = close

    def transport(self) -> str:
        return TRANSPORT[self._lib.acg_comm_transport(self._h)]

    def find_overlapping(self, ac, hay_ptr, hay_len, hay_global_offset, span, on_device=True, host_out=False):
        """Collective.  Returns (n_total, device pointer of the records on rank 0 (0 elsewhere), stats,
        host array on rank 0 if host_out)."""
        ab = _ab()
        dptr, n, st = C.c_void_p(), C.c_uint64(), ShardStats()
        out, cap = None, 0
        rc = self._lib.acg_find_overlapping_sharded(ac._h, self._h, hay_ptr, 1 if on_device else 0, hay_len,
                                                    hay_global_offset, span[0], span[1], C.byref(dptr),
                                                    C.byref(n), None, 0, C.byref(st))
        if rc:
            ab.AhoCorasick._raise(rc)
        if host_out and self.rank == 0:
            out = self.fetch_view()
        stats = {k: getattr(st, k) for k, _ in ShardStats._fields_}
        return n.value, (dptr.value or 0), stats, out

    def begin(self, ac, hay_ptr, hay_len, hay_global_offset, span, on_device=True) -> int:
        """Collective, first half of find_overlapping (acg_find_overlapping_sharded_begin): returns a
        ticket once this rank's records are on their way; at most two steps in flight."""
        t = C.c_int()
        rc = self._lib.acg_find_overlapping_sharded_begin(ac._h, self._h, hay_ptr, 1 if on_device else 0, hay_len,
                                                          hay_global_offset, span[0], span[1], C.byref(t))
        if rc:
            _ab().AhoCorasick._raise(rc)
        return t.value

    def wait(self, ticket: int):
        """Collective, second half: (n_total, device pointer on rank 0, stats)."""
        dptr, n, st = C.c_void_p(), C.c_uint64(), ShardStats()
        rc = self._lib.acg_find_overlapping_sharded_wait(self._h, ticket, C.byref(dptr), C.byref(n), None, 0, C.byref(st))
        if rc:
            _ab().AhoCorasick._raise(rc)
        return n.value, (dptr.value or 0), {k: getattr(st, k) for k, _ in ShardStats._fields_}

    def mark(self, which: int) -> None:
        """Device timestamp (acg_comm_mark): 0 before a stream of steps, 1 after it."""
        rc = self._lib.acg_comm_mark(self._h, which)
        if rc:
            raise _ab().DeviceError(rc)

    def mark_elapsed_ms(self) -> float:
        ms = C.c_float()
        rc = self._lib.acg_comm_mark_elapsed_ms(self._h, C.byref(ms))
        if rc:
            raise _ab().DeviceError(rc)
        return ms.value

    def fetch(self) -> np.ndarray:
        n = C.c_uint64()
        self._lib.acg_comm_fetch(self._h, None, 0, C.byref(n))
        out = np.empty(n.value, MATCH_DTYPE)
        rc = self._lib.acg_comm_fetch(self._h, out.ctypes.data, n.value, C.byref(n))
        if rc:
            raise _ab().DeviceError(rc)
        return out

    def fetch_view(self) -> np.ndarray:
        """The gathered records as a view of the communicator's page-locked host buffer (valid until the
        next fetch_view / search on this communicator)."""
        n, ptr = C.c_uint64(), C.c_void_p()
        rc = self._lib.acg_comm_fetch_view(self._h, C.byref(ptr), C.byref(n))
        if rc:
            raise _ab().DeviceError(rc)
        if n.value == 0:
            return np.zeros(0, MATCH_DTYPE)
        buf = (C.c_uint8 * (n.value * MATCH_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=MATCH_DTYPE)

    def checksum(self):
        n, f = C.c_uint64(), C.c_uint64()
        rc = self._lib.acg_comm_checksum(self._h, C.byref(n), C.byref(f))
        if rc:
            raise _ab().DeviceError(rc)
        return n.value, f.value


def slice_plan(span_start: int, span_end: int, world: int, max_pattern_len: int):
    """Per-rank (own_lo, own_hi, read_lo) from acg_shard_plan: rank g owns ends in (own_lo, own_hi]
    and scans the bytes [read_lo, own_hi)."""
    lib = _ab()._lib
    plan = []
    for g in range(world):
        lo, hi, rd = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = lib.acg_shard_plan(span_start, span_end, world, g, max_pattern_len, C.byref(lo), C.byref(hi), C.byref(rd))
        if rc:
            raise ValueError("acg_shard_plan rejected the arguments")
        plan.append((lo.value, hi.value, rd.value))
    return plan


def owned(matches: np.ndarray, rank: int, own_lo: int) -> np.ndarray:
    """Drop the matches that belong to the previous rank (end <= own_lo); rank 0 keeps everything."""
    if rank == 0:
        return matches
    return matches[matches["end"] > own_lo]


def gather_to_rank0(local, dist, device=None):
    """Gather variable-length match buffers (torch uint8 tensors holding acg_match records, or numpy
    structured arrays) to rank 0 in rank order.  Counts travel by all_gather, payloads by a padded
    gather.  Returns the concatenated numpy array on rank 0, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    if isinstance(local, np.ndarray):
        t = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
        if device is not None:
            t = t.to(device)
    else:
        t = local.reshape(-1)
    dev = t.device
    n = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 24)
    padded = torch.zeros(width, dtype=torch.uint8, device=dev)
    padded[: t.numel()] = t
    bufs = [torch.empty(width, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0)
    if rank != 0:
        return None
    parts = [b[:c].cpu().numpy().view(MATCH_DTYPE) for b, c in zip(bufs, counts)]
    return np.concatenate(parts) if parts else np.zeros(0, MATCH_DTYPE)
