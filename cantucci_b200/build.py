"""Build recipe for libcantucci_b200.so (nvcc, sm_100a only, in-tree)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcantucci_b200.so")
SOURCES = [os.path.join(CSRC, "cantucci_b200.cu")]
DEPS = SOURCES + [
    os.path.join(CSRC, "kernels.cuh"),
    os.path.join(CSRC, "de_device.cuh"),
    os.path.join(CSRC, "multi.inc"),
    os.path.join(CSRC, "interop.inc"),
    os.path.join(HERE, "..", "include", "cantucci_b200.h"),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    # host arithmetic for the span geometry must not be contracted either
    "-Xcompiler", "-ffp-contract=off",
    "-Xcompiler", "-fvisibility=hidden",
    "--cudart", "static",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build_variant(out: str, defines: list[str]) -> str:
    """An A/B build of the same library with extra -D flags (select it with CANTUCCI_B200_LIB)."""
    subprocess.run([nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, *SOURCES], check=True)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
