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


def usable_cores():
    """Host threads this process can really run at once: the CPU count, cut by the affinity mask
    and by a cgroup CPU quota (a container on a 128-thread host may be limited to far fewer)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = Path(path).read_text().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]) + 0.5)))
            else:
                quota = int(txt[0])
                period = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
                if quota > 0:
                    n = min(n, max(1, int(quota / period + 0.5)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)


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
    cores = usable_cores()
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
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--device-fill", action="store_true",
                    help="build the dense table on the GPU (acg_build_on_device); see build_s in the output")
    ap.add_argument("--experiment", type=int, default=0,
                    help="ACG_EXP_* flags (include/acb200_debug.h): kernel variants awaiting measurement; 0 = default kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import aho_corasick_b200 as ab
    from aho_corasick_b200 import sharded as S
    from aho_corasick_b200 import workload as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    cfg = W.CONFIGS[args.workload]
    per_gpu = int(args.hay_gib * GIB)
    per_gpu -= per_gpu % 4096
    total = per_gpu * world  # weak scaling: the global haystack grows with the number of GPUs
    pats = W.make_patterns(cfg["n_patterns"], cfg["pattern_seed"], alphabet=cfg["alphabet"])
    overlapping = args.workload in ("cfg2", "cfg5")
    b = ab.AhoCorasick.builder().kind(ab.AhoCorasickKind.DFA)
    if args.workload == "cfg3":
        b.ascii_case_insensitive(True).match_kind(ab.MatchKind.LeftmostFirst)
    if args.workload == "cfg4":
        b.match_kind(ab.MatchKind.LeftmostFirst)
    if args.device_fill:
        b.device_fill(True)
    t0 = time.perf_counter()
    ac = b.build(pats).set_engine(args.engine)
    build_s = time.perf_counter() - t0
    if args.experiment:
        import ctypes
        ab._lib.acg_debug_set_experiment.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        assert ab._lib.acg_debug_set_experiment(ac._h, args.experiment) == 0
    # haystack slicing: this rank owns ends in (own_lo, own_hi] and reads from read_lo
    own_lo, own_hi, read_lo = S.slice_plan(0, total, world, ac.max_pattern_len())[rank]
    gen_lo = read_lo - read_lo % 4096
    n_local = own_hi - gen_lo
    n_local += (-n_local) % 8
    d_hay = torch.empty(n_local, dtype=torch.uint8, device=dev)
    W.torch_fill_config(args.workload, d_hay, pats, global_offset=gen_lo)
    torch.cuda.synchronize()
    span = (read_lo - gen_lo, own_hi - gen_lo)
    n_bytes = own_hi - own_lo  # bytes this rank is credited with (overlap re-reads are not)
    cap = max(1 << 20, n_bytes // 512)
    d_out = torch.empty(cap * 24, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gatherer = None

    def step():
        """One pass of the hot path over this rank's slice, matches left on the device in global
        offsets; for N > 1 followed by the NCCL gather of the match buffers to rank 0."""
        nonlocal cap, d_out
        while True:
            try:
                if overlapping:
                    n, ms = ac.find_overlapping_devout(d_hay.data_ptr(), n_local, span, own_lo - gen_lo, gen_lo,
                                                       d_out.data_ptr(), cap)
                else:
                    r, ms = ac.find_iter_dev_np(d_hay.data_ptr(), n_local, span)
                    n = len(r)
                break
            except OverflowError as e:
                cap = int(e.args[0]) * 9 // 8 + 1024
                d_out = torch.empty(cap * 24, dtype=torch.uint8, device=dev)
        nonlocal gatherer
        gms = 0.0
        if world > 1 and overlapping:
            # every rank must use the same payload width: agree on the capacity once (and again
            # only if some rank had to grow its buffer)
            capt = torch.tensor([cap], dtype=torch.int64, device=dev)
            if gatherer is None or gatherer.cap < cap * 24:
                dist.all_reduce(capt, op=dist.ReduceOp.MAX)
                gcap = int(capt.item())
                if gcap > cap:
                    cap = gcap
                    grown = torch.empty(cap * 24, dtype=torch.uint8, device=dev)
                    grown[: n * 24] = d_out[: n * 24]
                    d_out = grown
                gatherer = S.MatchGatherer(dist, dev, cap * 24)
            ev0.record()
            gatherer.gather(d_out, n * 24)
            ev1.record()
            ev1.synchronize()
            gms = ev0.elapsed_time(ev1)
        return n, ms, gms, None

    # ---- device-resident throughput (inputs already in HBM) ----
    for _ in range(args.warmup):
        step()
    barrier()
    kernel_ms, scan_ms, gather_ms = [], [], []
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cnt, ms, gms, gathered = step()
            st = ac.last_stats()
            kernel_ms.append(ms)
            scan_ms.append(st["scan_ms"])
            gather_ms.append(gms)
        barrier()
        wall = time.perf_counter() - t0
    stats = ac.last_stats()
    dev_s = (sum(kernel_ms) + sum(gather_ms)) / 1e3
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall = t.tolist()
    value = total * args.steps / GIB / dev_s
    matches = torch.tensor([cnt], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(matches)
    total_matches = int(matches.item())
    if world > 1 and rank == 0 and gatherer is not None:
        gathered = gatherer.result_numpy()  # outside the timed region: check the gathered stream
        assert len(gathered) == total_matches and bool(np.all(np.diff(gathered["end"].astype(np.int64)) >= 0))

    # ---- end to end through the host-buffer C-ABI call (pinned host haystack, H2D inside) ----
    e2e_val, e2e_bytes, d2h = None, 0, 0
    if not args.no_e2e:
        e2e_bytes = min(span[1] - span[0], 1 << 30)
        h_hay = torch.empty(e2e_bytes, dtype=torch.uint8, pin_memory=True)
        h_hay.copy_(d_hay[span[0]: span[0] + e2e_bytes])
        h_np = h_hay.numpy()
        call = ac.try_find_overlapping_iter_np if overlapping else ac.try_find_iter_np
        for _ in range(2):
            r = call(h_np)
        barrier()
        e2e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r = call(h_np)
            if world > 1:
                S.gather_to_rank0(r, dist, device=dev)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_val = world * e2e_bytes / GIB / te.item()
        d2h = int(len(r) * 12 + 16)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, which = peaks()
    scan_s = sum(scan_ms) / len(scan_ms) / 1e3
    achieved = n_bytes / scan_s / 1e9
    kname = {1: "walk_overlapping_kernel", 2: "prefilter_kernel", 3: "seq_find_kernel"}[int(stats["engine"])]
    traffic = None
    tf = ROOT / "profiles" / "r01_dram_traffic.json"
    if tf.exists():
        rec = json.loads(tf.read_text()).get(f"{args.workload}:{kname}")
        if rec:  # measured with `ncu --set full` on this kernel; scaled to this launch's bytes
            traffic = rec["dram_bytes_per_haystack_byte"] * n_bytes
    desc = {"cfg2": "5000 random 4-16B printable-ASCII patterns, DFA, MatchKind::Standard, find_overlapping_iter",
            "cfg3": "5000 patterns, ascii_case_insensitive, DFA, MatchKind::LeftmostFirst, find_iter",
            "cfg4": "50 literals (Teddy-active set), MatchKind::LeftmostFirst, find_iter",
            "cfg5": "100000 patterns, DFA, MatchKind::Standard, find_overlapping_iter"}[args.workload]
    line = {
        "metric": "haystack_scan_throughput", "value": value, "unit": "GiB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}; {per_gpu / GIB:g} GiB synthetic haystack per GPU, "
                               "~1 planted pattern per 4 KiB",
                   "haystack_bytes_per_gpu": per_gpu, "global_haystack_bytes": total,
                   "l2": "input per launch is far larger than the 126 MB L2",
                   "engine": kname, "experiment": args.experiment, "device_fill": bool(args.device_fill), "table_bytes": ac.memory_usage(), "states": ac.state_len(),
                   "sharding": "haystack slices, max_pattern_len-1 overlap, NCCL gather of match buffers to rank 0"
                               if world > 1 else "single GPU"},
        "matches": total_matches, "matches_per_s": total_matches * args.steps / dev_s,
        "candidates": int(stats["candidates"]), "scan_ms": sum(scan_ms) / len(scan_ms),
        "order_ms": float(stats["order_ms"]), "gather_ms": sum(gather_ms) / len(gather_ms),
        "build_s": build_s, "wall_ms_per_step": wall / args.steps * 1e3,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                     "kernel": kname, "algorithmic_bytes_per_launch": n_bytes},
        "gpu_launches": int(stats["launches"]) * args.steps,
        "clocks": clocks.summary(),
    }
    if e2e_val is not None:
        line["e2e"] = {"value": e2e_val, "unit": "GiB/s", "h2d_bytes_per_step": e2e_bytes,
                       "d2h_bytes_per_step": d2h}
    if not args.no_cpu_baseline and world == 1 and overlapping:
        import oracle_py as O
        sample = min(64 << 20, e2e_bytes or (64 << 20))
        o = O.Oracle(pats, kind=O.KIND_DFA)
        h = (h_np if not args.no_e2e else d_hay[span[0]: span[0] + sample].cpu().numpy())[:sample]
        t0 = time.perf_counter()
        o.scan_overlapping_count(h)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample / GIB / dt, "unit": "GiB/s", "cores": 1, "kind": "port",
                                "sample": f"first {sample >> 20} MiB of the same haystack, scalar DFA loop "
                                          "(src/automaton.rs:1491-1534 restated in oracle/), 1 thread"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
