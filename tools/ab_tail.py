"""A/B on one GPU: one-launch apply (fused tail, semb_tail.cuh) against the separate seam kernels (SEMB_NO_TAIL=1), same run.
    python tools/ab_tail.py [quick]
Per mesh: ms per fused Poisson apply (L2 flushed before every apply on the small meshes, cost subtracted) and ms per PCG
iteration; the headline mesh also gets a 20-apply burst and a 300-apply back-to-back figure."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralelements_jl_b200 as sem

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = sem.init(0)


def timed(fn, n, flush=False):
    for _ in range(3):
        if flush:
            ctx.flush_l2()
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(n):
        if flush:
            ctx.flush_l2()
        fn()
    ms = ctx.timer_stop()
    if flush:
        ctx.timer_start()
        for _ in range(n):
            ctx.flush_l2()
        ms -= ctx.timer_stop()
    return ms / n


MESHES = [(9, 1112, 1112, 0.0), (9, 256, 256, 1.0), (9, 1112, 139, 0.0), (13, 776, 776, 0.0), (13, 776, 97, 0.0)]
if quick:
    MESHES = MESHES[:3]
print("%-22s %-8s %10s %10s %10s %10s  plan" % ("mesh", "mode", "apply us", "burst20", "sust300", "pcg us"))
for nr, Ex, Ey, k in MESHES:
    for mode in ("tail", "seams"):
        if mode == "seams":
            os.environ["SEMB_NO_TAIL"] = "1"
        else:
            os.environ.pop("SEMB_NO_TAIL", None)
            os.environ["SEMB_FORCE_TAIL"] = "1"
        m = sem.Mesh(nr, nr, Ex, Ey, (False, False), "wavy", ctx=ctx)
        u, out = m.field().fill_random(1), m.field()
        fn = lambda: m.oplhs_device(u, out, nu=1.0, k=k, bc="DDDD")
        small = Ex * Ey < 400000
        ap = timed(fn, 100, flush=small)
        burst = sust = float("nan")
        if not small:
            time.sleep(1.0)
            burst = timed(fn, 20)
            sust = timed(fn, 300)
        x = m.field()
        m.pcg_begin(u, x, nu=1.0, k=k, bc="DDDD", tol=0.0, maxiter=10 ** 9)
        m.pcg_iterate(5)
        ctx.sync()
        ctx.timer_start()
        m.pcg_iterate(100)
        pcg = ctx.timer_stop() / 100
        print("%-22s %-8s %10.2f %10.2f %10.2f %10.2f  %s" % ("nr=%d %dx%d k=%g" % (nr, Ex, Ey, k), mode, ap * 1e3, burst * 1e3,
                                                            sust * 1e3, pcg * 1e3, m.plan()), flush=True)
        m.free()
os.environ.pop("SEMB_NO_TAIL", None)
os.environ.pop("SEMB_FORCE_TAIL", None)
