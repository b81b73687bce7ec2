"""Short profiling driver for ncu: a few FDM preconditioner applications (+ operator applies) on one mesh."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralelements_jl_b200 as sem
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 9
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1112
ctx = sem.init(0)
m = sem.Mesh(nr, nr, E, E, (False, False), "wavy", ctx=ctx)
u, out = m.field().fill_random(1), m.field()
P = sem.FdmPrecond(m, "DDDD", 1.0, 0.0)
for _ in range(3):
    P.apply_device(u, out)
    m.oplhs_device(u, out, nu=1.0, k=0.0, bc="DDDD")
ctx.sync()
print("done")
