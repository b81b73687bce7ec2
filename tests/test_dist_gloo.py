"""N>1 path on CPU: world_size 2 and 3 over gloo (tests/dist_cpu_protocol.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_protocol_over_gloo(world, sem):  # `sem` makes sure libsemb.so exists for the spawned ranks
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "dist_cpu_protocol.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "DIST_CPU OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_halo_plan(sem):
    assert sem.halo_plan(1, 0, False) == (0, 0, -1, -1)
    assert sem.halo_plan(1, 0, True) == (0, 0, -1, -1)      # single rank: periodic wrap is a local seam
    assert sem.halo_plan(2, 0, False) == (0, 1, -1, 1)
    assert sem.halo_plan(2, 1, False) == (1, 0, 0, -1)
    assert sem.halo_plan(2, 0, True) == (1, 1, 1, 1)         # both neighbours are the same peer
    assert sem.halo_plan(8, 0, True) == (1, 1, 7, 1)
    assert sem.halo_plan(8, 7, True) == (1, 1, 6, 0)
    assert sem.halo_plan(8, 3, False) == (1, 1, 2, 4)
