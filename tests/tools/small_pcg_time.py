"""Wall time of complete small PCG solves (launch-bound regime): CUDA-graph replay vs plain launches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import sem_oracle as so
import spectralelements_jl_b200 as sem
ctx = sem.init(0)
for nr, E in ((9, 8), (9, 32), (9, 64)):
    msh = sem.Mesh(nr, nr, E, E, (False, False), sem.wavy if E <= 64 else "wavy", ctx=ctx)
    f, t1, rhs, x = msh.field().fill(1.0), msh.field(), msh.field(), msh.field()
    msh.mass_device(f, t1); msh.mask_bc_device(t1, "DDDD", f); msh.gs_device(f, rhs)
    for tag, env in (("graph", None), ("plain", "1")):
        if env: os.environ["SEMB_NO_GRAPH"] = env
        else: os.environ.pop("SEMB_NO_GRAPH", None)
        msh.pcg_device(rhs, x, bc="DDDD")
        ctx.sync(); t0 = time.perf_counter()
        it, res, ok = msh.pcg_device(rhs, x, bc="DDDD")
        ctx.sync(); dt = time.perf_counter() - t0
        print("nr=%d E=%d dof=%d %s: %d iterations in %.2f ms (%.1f us/iter) res %.2e" % (nr, E, msh.shape[0] * msh.shape[1], tag, it, dt * 1e3, dt * 1e6 / it, res))
    msh.free()
