"""Per-CTA timeline of the one-launch apply (needs a library built with SEMB_EXTRA_FLAGS=-DSEMB_TAIL_TIMING):
when the CTAs start, finish their rows, have announced, have run their interface tasks -- relative to the first start.
    SEMB_EXTRA_FLAGS=-DSEMB_TAIL_TIMING python spectralelements.jl_b200/build.py --force; python tools/tail_timing.py 9 1112 139"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spectralelements_jl_b200 as sem

nr, Ex, Ey = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pcg = len(sys.argv) > 4 and sys.argv[4] == "pcg"
ctx = sem.init(0)
m = sem.Mesh(nr, nr, Ex, Ey, (False, False), "wavy", ctx=ctx)
u, out, x = m.field().fill_random(1), m.field(), m.field()
pl = m.plan()
ncta = pl["nstrips"] * pl["ngroups"]
if pcg:
    m.pcg_begin(u, x, nu=1.0, k=0.0, bc="DDDD", tol=0.0, maxiter=10 ** 9)
for rep in range(4):
    ctx.flush_l2()
    if pcg:
        m.pcg_iterate(1)
    else:
        m.oplhs_device(u, out, nu=1.0, k=0.0, bc="DDDD")
    buf = np.zeros(8 * ncta, dtype=np.int64)
    sem._lib.check(ctx.lib.semb_mesh_debug_read(m.h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), ncta))
d = buf.reshape(ncta, 8).astype(np.float64)
t0 = d[:, 0].min()
q = lambda v: "min %7.1f  med %7.1f  p90 %7.1f  max %7.1f" % (v.min(), np.median(v), np.percentile(v, 90), v.max())
us = lambda c: (d[:, c] - t0) / 1e3
print("mesh nr=%d %dx%d %s plan %s, %d CTAs (times in us after the first CTA start)" % (nr, Ex, Ey, "pcg" if pcg else "apply", pl, ncta))
print("start            ", q(us(0)))
print("rows done        ", q(us(1)))
print("announced        ", q(us(2)))
print("tail prologue    ", q(us(5)))
print("tasks done       ", q(us(6)))
print("rows duration    ", q(us(1) - us(0)))
print("fence+announce   ", q(us(2) - us(1)))
print("prologue (waits) ", q(us(5) - us(2)))
print("tasks            ", q(us(6) - us(5)))
print("micro-tasks/CTA  ", q(d[:, 7]))
last = int(np.argmax(d[:, 6]))
print("last CTA: rows done %.1f announced %.1f prologue %.1f tasks done %.1f, micro-tasks %d" % (us(1)[last], us(2)[last], us(5)[last], us(6)[last], d[last, 7]))
