"""Static checks of the built sm_100a binary (no GPU): the hot kernels must stay spill-free and within
the register budget their occupancy assumes, and the library must carry sm_100a SASS only."""
import re
import shutil
import subprocess

import pytest

from cantucci_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def res_usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    fn, table = None, {}
    for line in out.stdout.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            table[fn] = tuple(int(x) for x in m.groups())
    return out.stdout, table


def test_hot_kernels_are_spill_free_and_within_register_budget():
    text, table = res_usage()
    assert "sm_100a" in text and "sm_90" not in text and "sm_80" not in text

    def find(*needles):
        hits = [v for k, v in table.items() if all(n in k for n in needles)]
        assert len(hits) == 1, (needles, len(hits))
        return hits[0]
    # K1 fast / power 8 (two samples per thread, packed FP32): launch bounds (256, 3) -> at most 85 registers.
    # The only stack use is the call frame of the out-of-line exact re-evaluation (suspects, rare).
    reg, stack, _, local = find("sample_grids_kernelILb1ELi0")
    assert reg <= 85 and stack <= 64 and local == 0
    # E3 fast / power 8: launch bounds (256, 4)
    reg, stack, _, local = find("vertex_kernelILb1ELi0")
    assert reg <= 64 and stack <= 128 and local == 0
    # the classification / compaction kernels and the per-quad kernel: small, no stack, no local memory
    for name in ("classify_count_kernel", "span_scan_kernel", "emit_lists_kernel", "quad_kernelILb0", "quad_kernelILb1", "expand_quads_kernel"):
        reg, stack, _, local = find(name)
        assert reg <= 64 and stack == 0 and local == 0, name


def test_fast_k1_uses_the_mufu_and_fma_paths_it_was_designed_around():
    out = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    body, grab = [], False
    for line in out.stdout.splitlines():
        if "Function :" in line:
            grab = "sample_grids_kernelILb1ELi0" in line
        elif grab:
            body.append(line)
    sass = "\n".join(body)
    # packed FP32 (sm_100's FFMA2 / FMUL2 / FADD2) carries the iteration; MUFU the roots, log and reciprocal
    for op in ("MUFU.RSQ", "MUFU.SQRT", "MUFU.LG2", "MUFU.RCP", "FFMA2", "FMUL2", "FADD2", "SHF.L.W", "SHFL.BFLY", "RED"):
        assert op in sass, op
    # the iteration itself is packed: far more packed than scalar FMA-pipe instructions would be a
    # regression back to the issue-bound scalar form
    assert sass.count("FFMA2") + sass.count("FMUL2") + sass.count("FADD2") >= 60
