"""Short profiling driver: a few fused applies (and PCG iterations) on the headline mesh, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralelements_jl_b200 as sem

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 9
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1112
napply = int(sys.argv[3]) if len(sys.argv) > 3 else 4
npcg = int(sys.argv[4]) if len(sys.argv) > 4 else 2
k = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
precond = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False   # opM = u ./ B ./ b0 (convectionDiffusion.jl:87-91)
ctx = sem.init(0)
msh = sem.Mesh(nr, nr, E, E, (False, False), "wavy", ctx=ctx)
u, out = msh.field().fill_random(1), msh.field()
for _ in range(napply):
    msh.oplhs_device(u, out, nu=1.0, k=k, bc="DDDD")
if npcg:
    x = msh.field()
    msh.pcg_begin(u, x, nu=1.0, k=k, bc="DDDD", tol=0.0, maxiter=10 ** 9, precond=precond, prec_b0=max(k, 1.0))
    msh.pcg_iterate(3)
    ctx.sync()
    ctx.timer_start()
    msh.pcg_iterate(npcg)
    print("pcg: %.4f ms per iteration" % (ctx.timer_stop() / npcg))
ctx.sync()
print("done", msh.plan())
