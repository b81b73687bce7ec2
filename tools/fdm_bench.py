"""FDM preconditioner on the device: cost of one application beside one operator apply, cost of a
preconditioned PCG iteration, and Poisson solves with / without it (iterations and time to tol).
    python tools/fdm_bench.py [quick]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spectralelements_jl_b200 as sem

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = sem.init(0)


def timed(fn, n):
    for _ in range(3):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(n):
        fn()
    return ctx.timer_stop() / n


print("== cost per application (ms), device resident")
for nr, E in ((9, 1112), (13, 776)) if not quick else ((9, 512),):
    m = sem.Mesh(nr, nr, E, E, (False, False), "wavy", ctx=ctx)
    n = m.shape[0] * m.shape[1]
    u, out, x = m.field().fill_random(1), m.field(), m.field()
    P = sem.FdmPrecond(m, "DDDD", 1.0, 0.0)
    t_op = timed(lambda: m.oplhs_device(u, out, nu=1.0, k=0.0, bc="DDDD"), 50)
    t_fdm = timed(lambda: P.apply_device(u, out), 50)
    res = []
    for precond in (0, 2):
        m.pcg_begin(u, x, nu=1.0, k=0.0, bc="DDDD", tol=0.0, maxiter=10 ** 9, precond=precond)
        m.pcg_iterate(3)
        ctx.sync(); ctx.timer_start(); m.pcg_iterate(30); res.append(ctx.timer_stop() / 30)
    print("nr=%2d %dx%d (%.3e DOF): opLHS %.3f  fdm %.3f (%.1f GB/s of r+h)  pcg iter %.3f  pcg+fdm iter %.3f"
          % (nr, E, E, n, t_op, t_fdm, 16 * n / t_fdm / 1e6, res[0], res[1]), flush=True)
    m.free()

print("== Poisson solves, f = 1, bc DDDD, wavy box, tol 1e-8 * norm(b, Inf)")
print("%-18s %10s %10s %10s %10s %8s" % ("mesh", "its none", "s none", "its fdm", "s fdm", "speed-up"))
for nr, E in ((9, 64), (9, 128), (9, 256), (13, 64), (13, 128), (13, 256)) if not quick else ((9, 64),):
    m = sem.Mesh(nr, nr, E, E, (False, False), "wavy", ctx=ctx)
    b, rhs, x = m.field().fill(1.0), m.field(), m.field()
    m.mass_device(b, rhs); m.mask_bc_device(rhs, "DDDD", b); m.gs_device(b, rhs)
    tol = 1e-8 * m.norm_inf(rhs)
    P = sem.FdmPrecond(m, "DDDD", 1.0, 0.0)
    row = []
    for precond in (0, 2):
        o, keep = sem._pcg_opts(1.0, 0.0, "DDDD", None, precond, 1.0, tol, 200000, 0)
        it, res = sem.C.c_longlong(), sem.C.c_double()
        ctx.sync(); t0 = time.perf_counter()
        sem._lib.check(ctx.lib.semb_pcg(m.h, sem.C.byref(o), rhs.h, x.h, sem.C.byref(it), sem.C.byref(res)))
        ctx.sync(); row += [it.value, time.perf_counter() - t0]
    print("%-18s %10d %10.3f %10d %10.3f %8.2f" % ("nr=%d %dx%d" % (nr, E, E), row[0], row[1], row[2], row[3], row[1] / row[3]), flush=True)
    m.free()
