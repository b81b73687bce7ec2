"""Chunk-count sweep of the fused apply on one mesh (strip kernel plan: strips x chunks), L2 flushed per apply.
    python tools/sweep_chunks.py nr Ex Ey k  n1 n2 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralelements_jl_b200 as sem
from bench import time_steps

nr, E, Ey, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
ctx = sem.init(0)
msh = sem.Mesh(nr, nr, E, Ey, (False, False), "wavy", ctx=ctx)
u, out = msh.field().fill_random(1), msh.field()
n = msh.shape[0] * msh.shape[1]
fn = lambda: msh.oplhs_device(u, out, nu=1.0, k=k, bc="DDDD")
bpd = 48.0 if k else 40.0
print("mesh nr=%d %dx%d k=%g default plan" % (nr, E, Ey, k), msh.plan())
ms = time_steps(ctx, None, fn, 100, 5, flush=True) / 100
print("default      %.2f us  %.1f GDOF/s  %.0f GB/s" % (ms * 1e3, n / ms / 1e6, bpd * n / ms / 1e6))
for c in [int(a) for a in sys.argv[5:]]:
    msh.set_chunks(c)
    ms = time_steps(ctx, None, fn, 100, 5, flush=True) / 100
    print("chunks %4d  %.2f us  %.1f GDOF/s  %.0f GB/s" % (c, ms * 1e3, n / ms / 1e6, bpd * n / ms / 1e6))
