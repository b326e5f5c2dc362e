"""Builds tests/cpp/test_facade.cpp (the C++ mirror of the reference API, include/acb200.hpp) with
g++ against libacb200.so and runs it on the GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpp" / "test_facade.cpp"
EXE = ROOT / "tests" / "cpp" / "test_facade"


def _build():
    libdir = ROOT / "aho-corasick_b200"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", str(ROOT / "include"), str(SRC), "-o", str(EXE),
           "-L", str(libdir), "-lacb200", f"-Wl,-rpath,{libdir}"]
    subprocess.check_call(cmd)


def test_cpp_packed_host_checks():
    """-m "not gpu": acb200::packed on host-only searchers (construction contract, error behaviour)."""
    src = ROOT / "tests" / "cpp" / "test_packed_host.cpp"
    exe = ROOT / "tests" / "cpp" / "test_packed_host"
    libdir = ROOT / "aho-corasick_b200"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o",
                           str(exe), "-L", str(libdir), "-lacb200", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr


def test_cpp_facade_compiles():
    """-m "not gpu": the header and the test program must at least build and link."""
    _build()
    assert EXE.exists()


@pytest.mark.gpu
def test_cpp_facade_runs():
    _build()
    r = subprocess.run([str(EXE)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
