"""Throughput of the fused apply (and PCG iteration) for several polynomial sizes at ~target DOFs."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import spectralelements_jl_b200 as sem

ctx = sem.init(0)
target = float(sys.argv[1]) if len(sys.argv) > 1 else 5e7
orders = [int(a) for a in sys.argv[2:]] or [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17]
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
for nr in orders:
    E = max(2, int(round(target ** 0.5 / nr)))
    msh = sem.Mesh(nr, nr, E, E, (False, False), "wavy", ctx=ctx)
    n = msh.shape[0] * msh.shape[1]
    u, out, x = msh.field().fill_random(1), msh.field(), msh.field()
    res = {}
    for name, k in (("poisson", 0.0), ("helmholtz", 1.0)):
        f = lambda: msh.oplhs_device(u, out, nu=1.0, k=k, bc="DDDD")
        for _ in range(3): f()
        ctx.sync(); ctx.timer_start()
        for _ in range(20): f()
        ms = ctx.timer_stop() / 20
        res[name] = (n / ms / 1e6, (40 + 8 * (k != 0)) * n / ms / 1e6 / peak)
    msh.pcg_begin(u, x, nu=1.0, k=0.0, bc="DDDD", tol=0.0, maxiter=10 ** 9)
    msh.pcg_iterate(3); ctx.sync(); ctx.timer_start(); msh.pcg_iterate(20); pms = ctx.timer_stop() / 20
    r, s, o = C.c_int(), C.c_int(), C.c_int()
    ctx.lib.semb_strip_kernel_info(nr, 0, 0, C.byref(r), C.byref(s), C.byref(o))
    print("nr=%2d E=%4d dof=%.3e plan=%s regs=%d smem=%d occ=%d | apply %.1f GDOF/s (%.0f%% roof) | helm %.1f (%.0f%%) | pcg %.1f it/s (%.0f%% of 112B roof)"
          % (nr, E, n, msh.plan(), r.value, s.value, o.value, res["poisson"][0], 100 * res["poisson"][1], res["helmholtz"][0],
             100 * res["helmholtz"][1], 1e3 / pms, 100 * 112 * n / pms / 1e6 / peak), flush=True)
    for f_ in (u, out, x): f_.free()
    msh.free()
