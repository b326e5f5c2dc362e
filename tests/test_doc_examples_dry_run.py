"""The reference's doc examples (tests/doc_examples.py) on the CPU dry-run library of tests/emu/."""
import ctypes
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "emu"))
import aho_corasick_b200 as ab  # noqa: E402
import doc_examples  # noqa: E402
from aho_corasick_b200 import packed  # noqa: E402


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


@pytest.mark.parametrize("example", doc_examples.ALL, ids=lambda f: f.__name__)
def test_doc_example(example):
    example(ab)
