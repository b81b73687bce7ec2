"""Where the fixed cost of one pressureProject call goes: wall-clock (with stream syncs) of its stages through the C ABI.
    python tools/stokes_fixed_cost.py [nr E]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralelements_jl_b200 as sem
from spectralelements_jl_b200._lib import check

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 11
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = sem.init(0)
mV = sem.Mesh(nr, nr, E, E, (False, False), "wavy", ctx=ctx)
mP = sem.Mesh(nr - 2, nr - 2, E, E, (False, False), "wavy", ctx=ctx)
sks = sem.Stokes("DDDD", "DDDD", mV, mP, 1.0)
lib = sks.lib
vx, vy, v1, v2 = mV.field().fill_random(5), mV.field().fill_random(6), mV.field(), mV.field()
rhs, dp, q = mP.field(), mP.field(), mP.field().fill_random(3)


def timed(name, fn, reps=1):
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    print("%-44s %8.3f ms" % (name, (time.perf_counter() - t0) * 1e3 / reps))


it, res = C.c_longlong(), C.c_double()
for rep in range(2):
    print("-- pass", rep)
    timed("semb_stokes_rhs", lambda: check(lib.semb_stokes_rhs(sks.h, vx.h, vy.h, rhs.h)))
    for n in (1, 16, 17, 48):
        timed("semb_stokes_solve maxiter=%d" % n,
              lambda: lib.semb_stokes_solve(sks.h, rhs.h, dp.h, 0.0, n, C.byref(it), C.byref(res)))
    timed("semb_diverT", lambda: check(lib.semb_diverT(sks.h, dp.h, v1.h, v2.h)))
    timed("semb_approx_hlmz_inv", lambda: check(lib.semb_approx_hlmz_inv(mV.h, v1.h, 1.0, b"DDDD", v2.h)))
    timed("semb_stokes_op x10", lambda: check(lib.semb_stokes_op(sks.h, q.h, dp.h)), reps=10)
    timed("field create+destroy (pressure mesh)", lambda: mP.field().free())
    timed("project maxiter=1", lambda: sks.project_device(vx, vy, None, tol=0.0, maxiter=1))
    timed("project maxiter=48", lambda: sks.project_device(vx, vy, None, tol=0.0, maxiter=48))
