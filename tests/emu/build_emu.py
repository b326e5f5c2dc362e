#!/usr/bin/env python3
"""Builds tests/emu/libacb200_emu.so: the product's C++/CUDA sources compiled with g++ against the
dry-run runtime in this directory (see cuda_runtime.h).  Test infrastructure only."""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "aho-corasick_b200" / "csrc"
OUT = HERE / "libacb200_emu.so"
SOURCES = ["acb_build.cpp", "acb_kernels.cu", "acb_prefilter.cu", "acb_comm.cu", "acb_api.cu"]


def build(force=False, asan=False):
    """asan=True builds libacb200_emu_asan.so (-fsanitize=address); load it in a process started with
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0."""
    global OUT
    san = ["-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []
    OUT = HERE / ("libacb200_emu_asan.so" if asan else "libacb200_emu.so")
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + \
        list(HERE.rglob("*.h")) + list(HERE.rglob("*.cuh")) + list((ROOT / "include").glob("*.h"))
    if not force and OUT.exists() and all(d.stat().st_mtime <= OUT.stat().st_mtime for d in deps):
        return OUT
    objs = []
    for s in SOURCES:
        o = HERE / (s.replace(".", "_") + ("_asan" if asan else "") + ".o")
        cmd = ["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fPIC", "-Wall", "-Wno-unused-function",
               "-Wno-unknown-pragmas", "-Wno-unused-variable", "-Wno-unused-but-set-variable",
               "-I", str(HERE), "-I", str(CSRC), "-DACB_EMULATE=1", '-DACB_PTX_HEADER="acb_ptx_emu.cuh"',
               "-c", str(CSRC / s), "-o", str(o)] + san
        subprocess.check_call(cmd)
        objs.append(str(o))
    subprocess.check_call(["g++", "-shared", "-o", str(OUT)] + san + objs + ["-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
