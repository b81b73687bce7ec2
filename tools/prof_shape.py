"""Profiling driver for a rectangular slab shape (what one rank sees under strong scaling)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralelements_jl_b200 as sem
nr, Ex, Ey = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = sem.init(0)
msh = sem.Mesh(nr, nr, Ex, Ey, (False, False), "wavy", ctx=ctx)
u, out, x = msh.field().fill_random(1), msh.field(), msh.field()
for _ in range(5):
    msh.oplhs_device(u, out, nu=1.0, k=0.0, bc="DDDD")
msh.pcg_begin(u, x, nu=1.0, k=0.0, bc="DDDD", tol=0.0, maxiter=10 ** 9)
msh.pcg_iterate(5)
ctx.sync()
ctx.timer_start(); msh.pcg_iterate(50); ms = ctx.timer_stop() / 50
ctx.timer_start()
for _ in range(50): msh.oplhs_device(u, out, nu=1.0, k=0.0, bc="DDDD")
ma = ctx.timer_stop() / 50
n = msh.shape[0] * msh.shape[1]
print("plan", msh.plan(), "dof", n, "apply %.1f us (%.1f GDOF/s)  pcg iter %.1f us" % (ma * 1e3, n / ma / 1e6, ms * 1e3))
