"""Profiling / timing driver of the dealiased advection (advect.jl:45-64) and one ConvectionDiffusion step at BASELINE
configs[3] size:  python tools/prof_advect.py [nr nrd E reps steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spectralelements_jl_b200 as sem

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 9
nrd = int(sys.argv[2]) if len(sys.argv) > 2 else 14
E = int(sys.argv[3]) if len(sys.argv) > 3 else 512
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ctx = sem.init(0)
mV = sem.Mesh(nr, nr, E, E, (True, False), "wavy", ctx=ctx)
mD = sem.Mesh(nrd, nrd, E, E, (True, False), "wavy", ctx=ctx)
nV = mV.shape[0] * mV.shape[1]
T, vx, vy, out = mV.field().fill_random(5), mV.field().fill_random(6), mV.field().fill_random(7), mV.field()
lib = ctx.lib
for tag, env in (("tiled", None), ("generic fused", "SEMB_NO_TILED_ADVECT")):
    if env:
        os.environ[env] = "1"
    sem._lib.check(lib.semb_advect(mV.h, mD.h, T.h, vx.h, vy.h, out.h))
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        sem._lib.check(lib.semb_advect(mV.h, mD.h, T.h, vx.h, vy.h, out.h))
    ms = ctx.timer_stop() / reps
    byt = (8 * (8 + (nrd / nr) ** 2)) * nV
    print("advect %s: %.3f ms  (%d DOF, %.1f B/DOF algorithmic -> %.0f GB/s)" % (tag, ms, nV, byt / nV, byt / ms / 1e6))
    if env:
        del os.environ[env]
if steps:
    cdn = sem.ConvectionDiffusion("ps", list("NNDD"), mV, mD, None, None, Tf=1.0, dt=5e-3)
    for name, val in (("vx", 1.0), ("vy", 0.0), ("nu", 1e-3)):
        cdn._field(sem._DFN_FIELDS[name]).fill(val)
    cdn.u = np.sin(np.pi * mV.x) * np.sin(np.pi * mV.y)
    for _ in range(3):
        sem.step_b(cdn)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        sem.step_b(cdn)
    ctx.sync()
    print("cd2d step: %.2f ms (pcg iters %s)" % ((time.perf_counter() - t0) / steps * 1e3, cdn.pcg_iters[-1]))
