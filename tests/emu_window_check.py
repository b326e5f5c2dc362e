"""Run by tests/test_emulated_kernels.py::test_queue_windows_on_small_inputs in a subprocess with
ACB_EMU_WINSHIFT set: the prefilter kernel keeps 32-bit queued offsets inside windows of
2^kWinShift bytes (2 GiB on the device); the dry-run build reads the shift from the environment so
that a sub-MiB input crosses dozens of windows, in every kernel variant and tile distribution."""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "emu"))
import numpy as np  # noqa: E402
import aho_corasick_b200 as ab  # noqa: E402
import oracle_py as O  # noqa: E402
from aho_corasick_b200 import packed, workload as W  # noqa: E402
import build_emu  # noqa: E402

lib = ctypes.CDLL(str(build_emu.build()))
ab._declare(lib)
packed._declare(lib)
ab._lib = packed._lib = lib
lib.acg_debug_set_experiment.argtypes = [ctypes.c_void_p, ctypes.c_uint32]


def same(got, want, ctx):
    assert len(got) == len(want), (len(got), len(want), ctx)
    for k in ("pid", "start", "end"):
        assert np.array_equal(got[k], want[k]), (k, ctx)


for n_pat, seed, kind, ci, nbytes in ((5000, 0xAC5000, 0, False, 512 << 10), (5000, 0xAC5000, 1, True, 256 << 10),
                                      (50, 0xAC0050, 1, False, 256 << 10), (20000, 0xAC1000, 0, False, 192 << 10),
                                      (300, 31, 2, False, 128 << 10)):
    pats = W.make_patterns(n_pat, seed)
    hay = np.empty(nbytes, dtype=np.uint8)
    W.fill_haystack(hay, 5)
    W.plant(hay, pats, 6, period=512, window=256)
    o = O.Oracle(pats, match_kind=kind, ascii_case_insensitive=ci, kind=O.KIND_DFA)
    for flags in (0, 16, 32):
        ac = (ab.AhoCorasick.builder().match_kind(kind).ascii_case_insensitive(ci).kind(ab.AhoCorasickKind.DFA).build(pats))
        assert lib.acg_debug_set_experiment(ac._h, flags) == 0
        for phase in (0, 1, 15):
            view = np.zeros(hay.size + 32, dtype=np.uint8)[phase:phase + hay.size]
            view[:] = hay
            same(ac.find_iter_dev_np(view.ctypes.data, view.size)[0], o.find_iter_np(view), (n_pat, kind, flags, phase))
            if kind == 0:
                same(ac.find_overlapping_iter_dev_np(view.ctypes.data, view.size)[0], o.find_overlapping_iter_np(view), (n_pat, flags, phase))
print("WINDOWS OK")
