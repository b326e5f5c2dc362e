"""Multi-GPU parity on hardware: N processes, one per GPU, run the sharded
search through the C ABI; the stream gathered on rank 0 must have the count + FNV of the single-GPU
search over the same global haystack, including matches cut by the slice boundaries.  Skipped when
the box has fewer GPUs than ranks."""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def run_world(world, workload, total, mode="dev", shift=0, timeout=900):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with tempfile.TemporaryDirectory() as tmp:
        uid = str(Path(tmp) / "uid")
        procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "multirank_worker.py"), str(r), str(world), uid,
                                   workload, str(total), mode, str(shift)],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ))
                 for r in range(world)]
        outs = []
        for p in procs:
            try:
                outs.append(p.communicate(timeout=timeout)[0])
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
        assert all(p.returncode == 0 for p in procs), "\n----\n".join(outs)
        assert "MULTIRANK OK" in outs[0], outs[0]
        print(outs[0].strip().splitlines()[-1])


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("world", [2, 8])
def test_sharded_stream_equals_single_gpu_stream_cfg2(world):
    run_world(world, "cfg2", (1 << 30) + 4096 * 37 + 24)


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("shift", [0, 1, 2])
def test_two_ranks_every_kind_of_boundary_match(shift):
    """A match straddling the boundary / ending exactly at it / ending one byte behind it; the last
    run feeds host buffers (pipelined H2D inside the call)."""
    run_world(2, "cfg2", (64 << 20) + 1000 * shift, mode="host" if shift == 2 else "dev", shift=shift)


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("world", [2, 8])
def test_sharded_stream_equals_single_gpu_stream_cfg5(world):
    """BASELINE config 5's automaton (100 000 patterns) on a 1 GiB global haystack."""
    run_world(world, "cfg5", 1 << 30)
