"""Multi-GPU path (y-slab partition, NCCL halo exchange, gathered PCG scalars): launched as one
process per GPU with torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world, mode):
    """mode p2p: halo rows and PCG scalars travel through CUDA-IPC mapped peer memory inside the kernels;
    mode nccl: the ncclSend/Recv + ncclAllGather fallback (SEMB_NO_P2P=1).  Same parity bar for both."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = ["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + (20 if mode == "nccl" else 0)),
           os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    env = dict(os.environ)
    if mode == "nccl":
        env["SEMB_NO_P2P"] = "1"
    else:
        env.pop("SEMB_NO_P2P", None)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=env)
    assert p.returncode == 0 and "DIST_CHECK OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
