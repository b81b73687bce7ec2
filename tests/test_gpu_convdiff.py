"""Next row 8f-2: grad, dealiased advection and the ConvectionDiffusion stepper (examples/cd2d.jl) against the
oracle (oracle/sem_oracle.py: grad.jl:15-34, advect.jl:27-78, convectionDiffusion.jl:76-157)."""
import numpy as np
import pytest

import sem_oracle as so

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# (nr, nrd) pairs served by the register-tiled kernel (semb_advect_tile.cu) with full and ragged element batches
# (EB = 128 // nrd elements per CTA), and pairs only the generic fused / multi-pass kernels serve ((6, 8), (4, 9))
@pytest.mark.parametrize("nr,nrd,Ex,Ey,per,deform", [(8, 12, 5, 5, (True, False), "box"), (6, 9, 3, 4, (False, False), "wavy"),
                                                     (9, 14, 4, 3, (False, True), "annulus"),
                                                     (9, 14, 11, 2, (False, False), "wavy"), (9, 13, 19, 2, (True, False), "wavy"),
                                                     (5, 8, 17, 3, (False, False), "wavy"), (12, 18, 8, 2, (False, False), "wavy"),
                                                     (7, 10, 13, 2, (False, False), "wavy"), (10, 15, 9, 2, (False, False), "wavy"),
                                                     (11, 16, 9, 2, (False, False), "wavy"), (11, 17, 3, 2, (False, False), "wavy"),
                                                     (3, 5, 27, 2, (False, False), "wavy"), (4, 6, 22, 3, (True, True), "box"),
                                                     (5, 7, 19, 2, (False, False), "wavy"), (7, 11, 12, 2, (False, False), "wavy"),
                                                     (6, 8, 5, 3, (False, False), "wavy"), (4, 9, 5, 3, (False, False), "wavy")])
def test_grad_and_advect(sem, ctx, nr, nrd, Ex, Ey, per, deform):
    od = {"box": so.fixU, "wavy": so.wavy, "annulus": so.annulus}[deform]
    gd = {"box": sem.fixU, "wavy": sem.wavy, "annulus": sem.annulus}[deform]
    oV, oD = so.make_mesh(nr, nr, Ex, Ey, per, od), so.make_mesh(nrd, nrd, Ex, Ey, per, od)
    gV, gD = sem.Mesh(nr, nr, Ex, Ey, per, gd, ctx=ctx), sem.Mesh(nrd, nrd, Ex, Ey, per, gd, ctx=ctx)
    try:
        T = so.splitmix_uniform(gV.shape, seed=4)
        vx, vy = 1.0 + 0.3 * oV.x, np.cos(oV.y)
        # the device meshes computed their own metric terms; the amplification of coordinate rounding (~N^2 E eps)
        # bounds the agreement of everything that touches rx..sy
        gx, gy = sem.grad(T, gV)
        ox, oy = so.grad(T, oV)
        assert relerr(gx, ox) < 1e-11 and relerr(gy, oy) < 1e-11
        assert relerr(sem.advect(T, vx, vy, gV), so.advect(T, vx, vy, oV)) < 1e-11
        assert relerr(sem.advect(T, vx, vy, gV, gD), so.advect(T, vx, vy, oV, oD)) < 1e-11
    finally:
        gV.free()
        gD.free()


def test_cd2d_stepping(sem, ctx):
    """examples/cd2d.jl (advected sine wave, periodic x, BDF3/EXT3, nr=8 / nrd=12) for 12 steps, with a small
    viscosity so the implicit operator is a genuine Helmholtz solve with the diagonal preconditioner."""
    kx = ky = 1.0
    ux, uy = 1.0, 0.0
    ut = lambda x, y, t: np.sin(kx * np.pi * (x - ux * t)) * np.sin(ky * np.pi * (y - uy * t))
    zero = lambda x, y, t: 0 * x
    visc = lambda x, y, t: 1e-3 + 0 * x
    per = [True, False]
    oV, oD = so.make_mesh(8, 8, 5, 5, per), so.make_mesh(12, 12, 5, 5, per)
    oc = so.ConvectionDiffusion(list("NNDD"), oV, oD, 0 * oV.x + ux, 0 * oV.x + uy, Tf=1.0, dt=5e-3)
    oc.u = np.asfortranarray(ut(oV.x, oV.y, 0.0))
    gV, gD = sem.Mesh(8, 8, 5, 5, per, ctx=ctx), sem.Mesh(12, 12, 5, 5, per, ctx=ctx)
    try:
        gc = sem.ConvectionDiffusion("ps", list("NNDD"), gV, gD, 0 * gV.x + ux, 0 * gV.x + uy, Tf=1.0, dt=5e-3,
                                     set0=ut, setBC=zero, setF=zero, setNu=visc)
        gc.u = ut(gV.x, gV.y, 0.0)
        for _ in range(12):
            so.convdiff_step(oc, setBC=zero, setForcing=zero, setVisc=visc)
            sem.step_b(gc)
        assert gc.pcg_iters == oc.pcg_iters
        assert abs(gc.time[0] - oc.time[0]) < 1e-15 and np.allclose(gc.bdfA, oc.bdfA) and np.allclose(gc.bdfB, oc.bdfB)
        assert relerr(gc.u, oc.u) < 1e-9
        assert np.max(np.abs(gc.u - ut(gV.x, gV.y, gc.time[0]))) < 5e-3
        gc.free()
    finally:
        gV.free()
        gD.free()


def test_cd2d_stepping_with_fdm_preconditioner(sem, ctx):
    """Opt-in: the FDM preconditioner (lapl.jl:105-119) as the opM of the step's solve instead of the reference's
    u./B./b0: the same steps to the solver tolerance, fewer PCG iterations per step; set_precond("reference") restores the
    reference's counts."""
    ut = lambda x, y, t: np.sin(np.pi * (x - t)) * np.sin(np.pi * y)
    zero = lambda x, y, t: 0 * x
    visc = lambda x, y, t: 1e-3 + 0 * x
    per = [True, False]
    gV, gD = sem.Mesh(8, 8, 10, 10, per, ctx=ctx), sem.Mesh(12, 12, 10, 10, per, ctx=ctx)
    try:
        runs = {}
        for kind in ("reference", "fdm"):
            gc = sem.ConvectionDiffusion("ps", list("NNDD"), gV, gD, 0 * gV.x + 1.0, 0 * gV.x, Tf=1.0, dt=5e-3,
                                         set0=ut, setBC=zero, setF=zero, setNu=visc)
            gc.u = ut(gV.x, gV.y, 0.0)
            gc.set_precond(kind)
            for _ in range(8):
                sem.step_b(gc, tol=1e-10)
            runs[kind] = (gc.u, list(gc.pcg_iters))
            gc.free()
        (ur, ir), (uf, jf) = runs["reference"], runs["fdm"]
        assert relerr(uf, ur) < 1e-7                       # two iterates within the solver tolerance of the same steps
        assert sum(jf) * 2 <= sum(ir), (ir, jf)             # (3-4x fewer in practice; BDF start-up changes b0 step by step)
        assert all(j >= 1 for j in jf)
    finally:
        gV.free()
        gD.free()
