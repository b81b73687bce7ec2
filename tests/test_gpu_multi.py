"""Multi-GPU path (y-slab partition, NCCL halo exchange, gathered PCG scalars): launched as one
process per GPU with torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    """GPUs on this box, asked of a child process: importing torch HERE, after the session's libsemb context exists, is
    what an in-process count would depend on (on a 2-GPU box the in-process form skipped these tests when they ran after
    the other GPU tests)."""
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("mode", ["p2p", "nccl", "p2p-seams"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world, mode):
    """mode p2p: one launch per apply -- the strip kernel's edge CTAs store the boundary rows into the neighbours' memory
    (CUDA IPC) and its last CTAs finish the interfaces; PCG scalars are all-gathered the same way (graph-replayed loop);
    mode p2p-seams: peer memory with the separate push + seam kernels (SEMB_NO_TAIL=1);
    mode nccl: the ncclSend/Recv + ncclAllGather fallback (SEMB_NO_P2P=1).  Same parity bar for all three."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = ["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + {"p2p": 0, "nccl": 20, "p2p-seams": 40}[mode]),
           os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    env = dict(os.environ)
    env.pop("SEMB_NO_P2P", None)
    env.pop("SEMB_NO_TAIL", None)
    if mode == "nccl":
        env["SEMB_NO_P2P"] = "1"
    if mode == "p2p-seams":
        env["SEMB_NO_TAIL"] = "1"
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=env)
    assert p.returncode == 0 and "DIST_CHECK OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
