"""The C++ host mirror (cantucci_b200/host/cantucci.hpp) compiles against the C ABI and behaves like
the reference's Rust interface: CPU part = build + error paths, GPU part = result parity."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cantucci_b200")


def _build(tmp_path):
    exe = str(tmp_path / "host_mirror_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host_mirror_check.cpp"),
                    f"-L{PKG}", "-lcantucci_b200", f"-Wl,-rpath,{PKG}"], check=True)
    return exe


def fnv(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_cpp_host_mirror_builds_and_mirrors_asserts(tmp_path):
    out = subprocess.run([_build(tmp_path)], check=True, capture_output=True, text=True).stdout
    assert out.startswith("ok")


@pytest.mark.gpu
def test_cpp_host_mirror_matches_oracle(tmp_path, oracle):
    out = subprocess.run([_build(tmp_path), "gpu"], check=True, capture_output=True, text=True).stdout
    sh = oracle.mandelbulb(8, 6, 2.5)
    v, i, _ = oracle.generate_for_box(sh, oracle.make_span((0.0, 0.0, 0.0), (0.6, 0.6, 0.6)), 16)
    first = out.splitlines()[0].split()
    assert int(first[1]) == len(v) and int(first[3]) == len(i)
    assert int(first[5], 16) == fnv(v.tobytes()) and int(first[7], 16) == fnv(i.tobytes())
    d = np.float32(oracle.min_distance_from(sh, (0.3, 0.2, 0.1)))
    assert out.splitlines()[1].split()[1] == f"{d.view(np.uint32):08x}"
    # N2 mirror: the span meshed into interop buffers equals the host-buffer mesh; the empty span has an empty view;
    # order_spans puts the surface span (index 1) before the empty corner span (index 0)
    assert out.splitlines()[2] == "interop fd 1 empty 0 same 1 order 1 0"
