"""Haystack slicing across GPUs and the gather of match buffers to rank 0.

The path shards naturally: rank g owns the matches whose END lies in (a_g, b_g] (rank 0 also owns
end == span start, i.e. empty-pattern matches of the start state, src/automaton.rs:1456-1464), reads
max_pattern_len-1 bytes before a_g so that every pattern ending in its range is seen from a cold
start, and never exchanges haystack data with another rank.  The only collective is the gather of
the per-rank match buffers to rank 0 (torch.distributed: NCCL on GPUs, gloo in the CPU tests);
because the slices are ordered and every end offset has exactly one owner, concatenating the
per-rank buffers in rank order reproduces the single-GPU iteration order.
"""
from __future__ import annotations

import numpy as np

MATCH_DTYPE = np.dtype([("pid", "<u4"), ("_pad", "<u4"), ("start", "<u8"), ("end", "<u8")])


def slice_plan(span_start: int, span_end: int, world: int, max_pattern_len: int, align: int = 64):
    """Per-rank (own_lo, own_hi, read_lo): rank g owns ends in (own_lo, own_hi] and scans the bytes
    [read_lo, own_hi).  Boundaries are aligned so device loads stay vectorisable."""
    n = span_end - span_start
    back = max(max_pattern_len - 1, 0)
    plan = []
    for g in range(world):
        lo = span_start + (n * g // world) // align * align if g else span_start
        hi = span_start + (n * (g + 1) // world) // align * align if g + 1 < world else span_end
        plan.append((lo, hi, max(span_start, lo - back)))
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


class MatchGatherer:
    """Persistent-buffer version of gather_to_rank0 for the bench loop: match buffers stay on the
    device (rank 0 ends up with every rank's acg_match records in rank order), one small
    all_gather for the counts and one gather for the payload per call, no host round trips other
    than reading the counts."""

    def __init__(self, dist, device, cap_bytes):
        import torch
        self.dist, self.device = dist, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.cap = int(cap_bytes)
        self._torch = torch
        self.count = torch.zeros(1, dtype=torch.int64, device=device)
        self.counts = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.recv = ([torch.empty(self.cap, dtype=torch.uint8, device=device) for _ in range(self.world)]
                     if self.rank == 0 else None)

    def gather(self, local_buf, n_bytes):
        """local_buf: uint8 CUDA tensor of capacity >= cap (only the first n_bytes are meaningful)."""
        assert local_buf.numel() >= self.cap and n_bytes <= self.cap
        self.count.fill_(n_bytes)
        self.dist.all_gather_into_tensor(self.counts, self.count)
        # ship only what the fullest rank needs (one small host read of the counts), not the capacity
        width = int(self.counts.max().item())
        width = min(self.cap, (width + 4095) // 4096 * 4096)
        if width:
            recv = [b[:width] for b in self.recv] if self.rank == 0 else None
            self.dist.gather(local_buf[:width], recv, dst=0)
        return self.counts

    def result_numpy(self):
        """(rank 0) concatenated matches as a host array -- outside any timed region."""
        counts = self.counts.tolist()
        parts = [b[:c].cpu().numpy().view(MATCH_DTYPE) for b, c in zip(self.recv, counts)]
        return np.concatenate(parts)
