"""Multi-GPU parity of the genome-wide pass (BASELINE.md section 4, C4 gate): chromosomes sharded over
ranks and pooled by dfb_merge_results over the library's NCCL communicator must give exactly the
p-values and FWER of a single-GPU run over all chromosomes (R/mergeResults.R:216-235, 300-307, 380-386).
Needs >= 2 GPUs; spawns torchrun (rendezvous on 127.0.0.1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_pooled_equals_single_gpu(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "_merge_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MERGE-PARITY-OK" in r.stdout
