"""Timing driver of the Stokes split at BASELINE configs[4] size (order 10 velocity / order 8 pressure):
    python tools/prof_stokes.py [nr E napply iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spectralelements_jl_b200 as sem

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 11
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
napply = int(sys.argv[3]) if len(sys.argv) > 3 else 10
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 50
ctx = sem.init(0)
mV = sem.Mesh(nr, nr, E, E, (False, False), "wavy", ctx=ctx)
mP = sem.Mesh(nr - 2, nr - 2, E, E, (False, False), "wavy", ctx=ctx)
sks = sem.Stokes("DDDD", "DDDD", mV, mP, 1.0)
nV, nP = mV.shape[0] * mV.shape[1], mP.shape[0] * mP.shape[1]
q, out = mP.field().fill_random(3), mP.field()
sks.op_device(q, out)
ctx.sync()
l0 = ctx.launch_count()
ctx.timer_start()
for _ in range(napply):
    sks.op_device(q, out)
ms = ctx.timer_stop() / napply
print("opStokesLHS: %.3f ms per apply, %d launches (%d velocity DOF, %d pressure DOF) -> %.2f GDOF/s (velocity nodes)"
      % (ms, (ctx.launch_count() - l0) // napply, nV, nP, nV / ms / 1e6))
vx, vy, pr = mV.field().fill_random(5), mV.field().fill_random(6), mP.field()
ctx.sync()
for its in (1, iters):   # two calls: fixed cost (first-use allocations) vs per-iteration cost
    t0 = time.perf_counter()
    sks.project_device(vx, vy, pr, tol=0.0, maxiter=its)
    ctx.sync()
    dt = time.perf_counter() - t0
    print("pressureProject: %d PCG iterations in %.1f ms -> %.1f it/s (resinf %.3e)"
          % (sks.pcg_iters[-1], dt * 1e3, sks.pcg_iters[-1] / dt, sks.resinf))
