"""Multi-GPU parity (needs >= 2 GPUs on the box): launches tests/mgpu_worker.py under torchrun, one
process per GPU, NCCL inside libheat_b200 for the halo exchange and the dot-product all-reduce."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    n = _gpu_count()
    if n < world:
        pytest.skip(f"{world} GPUs needed, {n} visible (single-GPU box; the host-side plan logic is covered by "
                    "tests/test_dist_cpu.py with gloo)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + 11 * world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-4000:] + "\n" + p.stderr[-4000:]
    assert '"mgpu_ok": true' in p.stdout
