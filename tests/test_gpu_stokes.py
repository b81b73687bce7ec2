"""Next row 8f-4: the Stokes pressure/velocity split (gradT, diver, diverT, approxHlmzInv, the Schur operator
opStokesLHS, the pressure solve and pressureProject) against the oracle's reconstruction (oracle/sem_oracle.py:
grad.jl:44-63, diver.jl:17-104, stokes.jl:110-177 -- the reference code itself is not executable)."""
import numpy as np
import pytest

import sem_oracle as so

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


CASES = [(7, 3, 3, (False, False), "wavy", "DDDD", "DDDD"), (8, 4, 3, (True, False), "box", "NNDD", "NNDD"),
         (11, 2, 3, (False, False), "wavy", "DDDD", "DDNN"), (6, 5, 2, (False, True), "annulus", "DDNN", "DDNN")]


def meshes(sem, ctx, nr, Ex, Ey, per, deform):
    od = {"box": so.fixU, "wavy": so.wavy, "annulus": so.annulus}[deform]
    gd = {"box": sem.fixU, "wavy": sem.wavy, "annulus": sem.annulus}[deform]
    oV, oP = so.make_mesh(nr, nr, Ex, Ey, per, od), so.make_mesh(nr - 2, nr - 2, Ex, Ey, per, od)
    gV, gP = sem.Mesh(nr, nr, Ex, Ey, per, gd, ctx=ctx), sem.Mesh(nr - 2, nr - 2, Ex, Ey, per, gd, ctx=ctx)
    return oV, oP, gV, gP


@pytest.fixture(params=["tiled", "generic"])
def stokes_path(request):
    """the register-tiled element kernels + one-pass gatherScatter (default) and the generic chain of launches they
    replace (SEMB_NO_TILED_STOKES=1: ABu kernels, gradT/diver tile kernels, gs_x + seam_y + pointwise passes)"""
    import os
    if request.param == "generic":
        os.environ["SEMB_NO_TILED_STOKES"] = "1"
    yield request.param
    os.environ.pop("SEMB_NO_TILED_STOKES", None)


# ragged batches of the tiled kernels: Ex not a multiple of 128 // nr
CASES_OPS = CASES + [(9, 17, 2, (False, False), "wavy", "DDDD", "DDDD"), (5, 27, 2, (True, False), "wavy", "NNDD", "NNDD"),
                     (13, 10, 2, (False, False), "wavy", "DDDD", "DDDD"), (4, 33, 2, (False, False), "wavy", "DDDD", "DDDD")]


@pytest.mark.parametrize("nr,Ex,Ey,per,deform,bcx,bcy", CASES_OPS)
def test_stokes_operators(sem, ctx, stokes_path, nr, Ex, Ey, per, deform, bcx, bcy):
    oV, oP, gV, gP = meshes(sem, ctx, nr, Ex, Ey, per, deform)
    osk = so.make_stokes(list(bcx), list(bcy), oV, oP, b0=1.5)
    gsk = sem.Stokes(bcx, bcy, gV, gP, b0=1.5)
    try:
        u, v = so.splitmix_uniform(gV.shape, seed=1), so.splitmix_uniform(gV.shape, seed=2)
        q = so.splitmix_uniform(gP.shape, seed=3)
        # metric terms are recomputed on the device: agreement is bounded by the amplified coordinate rounding
        tol = 1e-11
        gx, gy = sem.gradT(u, gV)
        ox, oy = so.gradT(u, oV)
        assert relerr(gx, ox) < tol and relerr(gy, oy) < tol
        assert relerr(sem.approxHlmzInv(u, 1.5, gV, bcx), so.approxHlmzInv(u, 1.5, oV, osk.Mvx)) < tol
        assert relerr(sem.diver(u, v, gsk), so.diver(u, v, oV, osk.JrPV, osk.JsPV)) < tol
        dx, dy = sem.diverT(q, gsk)
        ex, ey = so.diverT(q, oV, osk.JrPV, osk.JsPV)
        assert relerr(dx, ex) < tol and relerr(dy, ey) < tol
        assert relerr(sem.opStokesLHS(q, gsk), so.opStokesLHS(q, osk)) < tol
        assert relerr(sem.makeStokesRHS(u, v, gsk), so.makeStokesRHS(u, v, osk)) < tol
        # adjointness on the device itself: <diver(u,v), q> = <u, qx> + <v, qy>
        lhs = np.sum(sem.diver(u, v, gsk) * q)
        rhs = np.sum(u * dx) + np.sum(v * dy)
        assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0)
    finally:
        gsk.free()
        gV.free()
        gP.free()


@pytest.mark.parametrize("nr,Ex,Ey,per,deform,bcx,bcy", CASES[:3])
def test_pressure_projection(sem, ctx, nr, Ex, Ey, per, deform, bcx, bcy):
    """pressureProject!: same PCG trajectory as the oracle (iteration count, iterate) and a discretely
    divergence-free velocity afterwards."""
    oV, oP, gV, gP = meshes(sem, ctx, nr, Ex, Ey, per, deform)
    osk = so.make_stokes(list(bcx), list(bcy), oV, oP, b0=1.0)
    gsk = sem.Stokes(bcx, bcy, gV, gP, b0=1.0)
    try:
        vx = so.mask(so.gatherScatter(so.splitmix_uniform(gV.shape, seed=5) * oV.mult, oV), osk.Mvx)
        vy = so.mask(so.gatherScatter(so.splitmix_uniform(gV.shape, seed=6) * oV.mult, oV), osk.Mvy)
        pr = np.zeros(gP.shape, order="F")
        tolp = 1e-9
        ox, oy, op = so.pressureProject(vx, vy, pr, osk, tol=tolp)
        gx, gy, gp = sem.pressureProject(vx, vy, pr, gsk, tol=tolp)
        # ~200 iterations on a singular (constant-pressure null space) system: rounding-level differences in the dot
        # products shift the stopping iteration by a few counts (the reference's own summation order is unpinned)
        assert abs(gsk.pcg_iters[-1] - osk.pcg_iters[-1]) <= max(2, 0.03 * osk.pcg_iters[-1]), (gsk.pcg_iters, osk.pcg_iters)
        assert gsk.resinf <= tolp
        scale = max(np.max(np.abs(vx)), np.max(np.abs(vy)))
        assert np.max(np.abs(gx - ox)) < 1e-6 * scale and np.max(np.abs(gy - oy)) < 1e-6 * scale
        div0 = np.max(np.abs(sem.makeStokesRHS(vx, vy, gsk)))
        div1 = np.max(np.abs(sem.makeStokesRHS(gx, gy, gsk)))
        assert div1 < 50 * tolp and div1 < 1e-6 * div0
    finally:
        gsk.free()
        gV.free()
        gP.free()


def test_stokes_error_behaviour(sem, ctx):
    """DimensionMismatch-style failures surface as SembError (no silent fallback): meshes that do not pair, fields of the
    wrong mesh, aliasing outputs, a zero b0, a bad bc string."""
    gV = sem.Mesh(7, 7, 3, 2, (False, False), sem.wavy, ctx=ctx)
    gP = sem.Mesh(5, 5, 3, 2, (False, False), sem.wavy, ctx=ctx)
    gQ = sem.Mesh(5, 5, 2, 2, (False, False), sem.wavy, ctx=ctx)   # different element count
    try:
        with pytest.raises(sem.SembError):
            sem.Stokes("DDDD", "DDDD", gV, gQ)
        with pytest.raises(sem.SembError):
            sem.Stokes("DDDD", "DDDD", gV, gP, b0=0.0)
        with pytest.raises(sem.SembError):
            sem.Stokes("DDXD", "DDDD", gV, gP)
        sks = sem.Stokes("DDDD", "DDDD", gV, gP)
        fv, fp, fq = gV.field(), gP.field(), gQ.field()
        lib = ctx.lib
        assert lib.semb_diver(sks.h, fv.h, fp.h, fp.h) < 0          # uy is a pressure-mesh field
        assert lib.semb_diverT(sks.h, fv.h, fv.h, fv.h) < 0         # pr must live on mshP
        assert lib.semb_diverT(sks.h, fp.h, fv.h, fv.h) < 0         # outputs alias
        assert lib.semb_stokes_op(sks.h, fp.h, fp.h) < 0            # out aliases q
        assert lib.semb_stokes_op(sks.h, fq.h, fp.h) < 0            # field of another mesh
        assert lib.semb_gradT(gV.h, fv.h, fv.h, fv.h) < 0
        assert lib.semb_approx_hlmz_inv(gV.h, fv.h, 0.0, b"DDDD", gV.field().h) < 0
        assert b"b0" in lib.semb_last_error()
        # a well-formed call still works after the failures
        q = so.splitmix_uniform(gP.shape, seed=3)
        assert np.all(np.isfinite(sem.opStokesLHS(q, sks)))
        sks.free()
    finally:
        gV.free()
        gP.free()
        gQ.free()
