"""N > 1 on real GPUs: both gather paths reproduce the single-GPU meshes (needs >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gathered_meshes_equal_single_gpu_run():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "multigpu_parity.py")],
                         capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_PARITY_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
