#!/usr/bin/env python3
"""bench.py -- BASELINE metric: GiB/s of haystack scanned (config 2: 5000 patterns, 4 GiB, DFA,
MatchKind::Standard overlapping) on N B200s, with roofline / cpu_baseline / e2e objects.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--hay-gib G] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GIB = float(1 << 30)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.samples, self.reasons, self._stop, self.index = [], set(), threading.Event(), index
        self.max_mhz = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def reference_arm(args):
    """The reference's own CPU path: the scalar DFA loop (src/automaton.rs:1491-1534 over
    src/dfa.rs:218-226) as restated in oracle/ (rustc is unavailable, so kind = "port"), run on
    all host cores by slicing the sample with max_pattern_len-1 overlap."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor

    import oracle_py as O
    from aho_corasick_b200 import workload as W
    import torch
    cores = os.cpu_count() or 1
    sample = min(int(args.hay_gib * GIB), cores * (24 << 20))
    sample -= sample % 8
    pats = W.make_patterns(5000, W.CONFIGS["cfg2"]["pattern_seed"])
    hay_t = torch.empty(sample, dtype=torch.uint8)
    W.torch_fill_config("cfg2", hay_t, pats, chunk=1 << 24)
    hay = hay_t.numpy()
    o = O.Oracle(pats, kind=O.KIND_DFA)
    back = o.max_pattern_len - 1
    bounds = [sample * i // cores for i in range(cores + 1)]

    def work(i):
        s = max(0, bounds[i] - back)
        return o.scan_overlapping_count(hay, span=(s, bounds[i + 1]))[0]

    def step():
        with ThreadPoolExecutor(cores) as ex:
            return sum(ex.map(work, range(cores)))
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / GIB / dt
    line = {"impl": "reference", "metric": "haystack_scan_throughput", "value": val, "unit": "GiB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "cfg2: 5000 random 4-16B printable-ASCII patterns, DFA, "
                                   "MatchKind::Standard overlapping", "sample_bytes": sample},
            "cpu_baseline": {"value": val, "unit": "GiB/s", "cores": cores, "kind": "port",
                             "sample": f"{sample >> 20} MiB of the cfg2 haystack, {cores} threads"},
            "e2e": {"value": val, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--hay-gib", type=float, default=4.0, help="haystack GiB per GPU (weak scaling)")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 walk, 2 prefilter")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import aho_corasick_b200 as ab
    from aho_corasick_b200 import workload as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_bytes = int(args.hay_gib * GIB)
    n_bytes -= n_bytes % 4096
    goff = rank * n_bytes  # weak scaling: every rank owns its own slice of the global stream
    pats = W.make_patterns(5000, W.CONFIGS["cfg2"]["pattern_seed"])
    ac = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA).build(pats).set_engine(args.engine)
    d_hay = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    planted = W.torch_fill_config("cfg2", d_hay, pats, global_offset=goff)
    torch.cuda.synchronize()
    # rank r additionally sees max_pattern_len-1 bytes before its slice? The slices are independent
    # haystack slices here (each rank scans its own 4 GiB cold), matches carry global offsets.

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_dev():
        return ac.count_overlapping_dev(d_hay.data_ptr(), n_bytes)

    # ---- device-resident throughput (inputs already in HBM) ----
    for _ in range(args.warmup):
        step_dev()
    barrier()
    kernel_ms, scan_ms = [], []
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cnt, fnv, ms = step_dev()
            st = ac.last_stats()
            kernel_ms.append(ms)
            scan_ms.append(st["scan_ms"])
        barrier()
        wall = time.perf_counter() - t0
    stats = ac.last_stats()
    # device time of the K steps = sum of per-step CUDA-event times of the library's kernels
    dev_s = sum(kernel_ms) / 1e3
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall = t.tolist()
    value = world * n_bytes * args.steps / GIB / dev_s
    matches = torch.tensor([cnt], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(matches)
    total_matches = int(matches.item())

    # ---- end to end through the host-buffer C-ABI call (pinned host haystack, H2D inside) ----
    e2e_bytes = min(n_bytes, 1 << 30)
    h_hay = torch.empty(e2e_bytes, dtype=torch.uint8, pin_memory=True)
    h_hay.copy_(d_hay[:e2e_bytes])
    h_np = h_hay.numpy()
    e2e_steps = 0 if args.no_e2e else max(2, min(args.steps, 5))
    r = []
    for _ in range(0 if args.no_e2e else 2):
        r = ac.try_find_overlapping_iter_np(h_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = ac.try_find_overlapping_iter_np(h_np)
    barrier()
    e2e_s = (time.perf_counter() - t0) / max(e2e_steps, 1)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_bytes / GIB / te.item()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, which = peaks()
    scan_s = sum(scan_ms) / len(scan_ms) / 1e3
    achieved = n_bytes / scan_s / 1e9
    line = {
        "metric": "haystack_scan_throughput", "value": value, "unit": "GiB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "cfg2: 5000 random 4-16B printable-ASCII patterns, "
                               f"{n_bytes / GIB:g} GiB synthetic ASCII haystack per GPU with ~1 planted pattern/4 KiB, "
                               "DFA, MatchKind::Standard, find_overlapping_iter",
                   "haystack_bytes_per_gpu": n_bytes, "l2": "input (>=1 GiB) is much larger than the 126 MB L2",
                   "engine": int(stats["engine"]), "table_bytes": ac.memory_usage(),
                   "states": ac.state_len()},
        "matches": total_matches, "matches_per_s": total_matches * args.steps / dev_s,
        "candidates": int(stats["candidates"]), "scan_ms": sum(scan_ms) / len(scan_ms),
        "order_ms": float(stats["order_ms"]),
        "wall_ms_per_step": wall / args.steps * 1e3,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": which,
                     "kernel": "walk_overlapping_kernel" if stats["engine"] == 1 else "prefilter",
                     "algorithmic_bytes_per_launch": n_bytes},
        "e2e": {"value": e2e_val, "unit": "GiB/s", "h2d_bytes_per_step": e2e_bytes,
                "d2h_bytes_per_step": int(len(r) * 12 + 8)},
        "gpu_launches": int(stats["launches"]) * args.steps,
        "clocks": clocks.summary(),
    }
    if not args.no_cpu_baseline and world == 1:
        import oracle_py as O
        sample = 64 << 20
        o = O.Oracle(pats, kind=O.KIND_DFA)
        h = h_np[:sample]
        t0 = time.perf_counter()
        c1 = o.scan_overlapping_count(h)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample / GIB / dt, "unit": "GiB/s", "cores": 1, "kind": "port",
                                "sample": "first 64 MiB of the same haystack, scalar DFA loop, 1 thread"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
