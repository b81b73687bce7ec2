"""Pins of the CPU oracle (oracle/sem_oracle.py).  The reference ships no golden vectors for this path
and cannot run here (no Julia), so the restatement is pinned by: closed-form GLL rules, polynomial
exactness of derivMat/interpMat, the explicit Kronecker-assembled operators of
/root/reference/examples/p2d_explicit.jl:142-195, analytic solutions, and algebraic identities."""
import math

import numpy as np
import pytest

import sem_oracle as so


def vec(a):
    return np.asarray(a).reshape(-1, order="F")


def test_gausslobatto_closed_forms():
    z, w = so.gausslobatto(2)
    assert np.allclose(z, [-1, 1]) and np.allclose(w, [1, 1])
    z, w = so.gausslobatto(3)
    assert np.allclose(z, [-1, 0, 1], atol=1e-16) and np.allclose(w, [1 / 3, 4 / 3, 1 / 3], rtol=1e-15)
    z, w = so.gausslobatto(4)
    assert np.allclose(z, [-1, -1 / math.sqrt(5), 1 / math.sqrt(5), 1], rtol=1e-15)
    assert np.allclose(w, [1 / 6, 5 / 6, 5 / 6, 1 / 6], rtol=1e-15)
    z, w = so.gausslobatto(5)
    assert np.allclose(z, [-1, -math.sqrt(3 / 7), 0, math.sqrt(3 / 7), 1], rtol=1e-15, atol=1e-16)
    assert np.allclose(w, [1 / 10, 49 / 90, 32 / 45, 49 / 90, 1 / 10], rtol=1e-15)
    for n in range(2, 21):
        z, w = so.gausslobatto(n)
        assert abs(w.sum() - 2) < 1e-14 and np.all(np.diff(z) > 0) and np.allclose(z, -z[::-1], atol=1e-16)
        # GLL quadrature is exact to degree 2n-3
        for d in range(0, 2 * n - 2, 2):
            assert abs(np.sum(w * z ** d) - 2 / (d + 1)) < 1e-13


@pytest.mark.parametrize("n", [2, 3, 5, 8, 9, 13, 17])
def test_derivmat_and_interpmat_exactness(n):
    z, _ = so.gausslobatto(n)
    D = so.derivMat(z)
    for k in range(n):  # D differentiates polynomials of degree <= n-1 exactly (derivMat.jl:9-35)
        dk = k * z ** (k - 1) if k else 0 * z
        assert np.max(np.abs(D @ z ** k - dk)) < 5e-12 * max(1, n ** 2)
    assert np.max(np.abs(D.sum(axis=1))) < 1e-12  # constants in the null space
    zo = np.linspace(-1, 1, 2 * n + 1)
    J = so.interpMat(zo, z)
    for k in range(n):
        assert np.max(np.abs(J @ z ** k - zo ** k)) < 1e-12
    assert np.allclose(so.interpMat(z, z), np.eye(n), atol=1e-13)


def test_semq_semmesh_ndgrid():
    Q = so.semq(3, 4, False)
    assert Q.shape == (12, 10) and np.all(Q.sum(axis=1) == 1) and Q[3, 3] == 1 and Q[4, 3] == 1
    Qp = so.semq(3, 4, True)
    assert Qp.shape == (12, 9) and Qp[11, 0] == 1 and Qp[0, 0] == 1
    z, w = so.semmesh(4, 5)
    assert z.size == 20 and z[0] == -1 and z[-1] == 1 and abs(w.sum() - 2) < 1e-14
    assert z[4] == z[5]  # duplicated interface node (mesh.jl:94 layout)
    x, y = so.ndgrid(np.arange(3.0), np.arange(2.0))
    assert x.shape == (3, 2) and x[2, 0] == 2 and y[0, 1] == 1


@pytest.mark.parametrize("nr,Ex,Ey,per,deform", [(5, 3, 3, (False, False), so.wavy), (4, 2, 3, (False, True), so.annulus),
                                                 (6, 2, 2, (True, False), so.fixU)])
def test_kronecker_assembled_cross_check(nr, Ex, Ey, per, deform):
    """examples/p2d_explicit.jl:142-180: A = Drs' G Drs, QQt = Q Q', M = R'R built with Kronecker products
    must agree with the matrix-free laplace / gatherScatter / mask."""
    m = so.make_mesh(nr, nr, Ex, Ey, per, deform)
    A, Bm, Q = so.kron_operators(m)
    u = so.splitmix_uniform(m.x.shape)
    shp = m.x.shape
    ref = (A @ vec(u)).reshape(shp, order="F")
    out = so.laplace(u, m.Dr, m.Ds, m.G11, m.G12, m.G22)
    assert np.max(np.abs(out - ref)) < 1e-13 * np.max(np.abs(ref))
    assert np.array_equal(so.mass(u, m), (Bm @ vec(u)).reshape(shp, order="F"))
    QQt = Q @ Q.T
    assert np.array_equal(so.gatherScatter(u, m), (QQt @ vec(u)).reshape(shp, order="F"))
    # index form == dense form, bit for bit
    assert np.array_equal(so.gatherScatter_index(u, nr, nr, Ex, Ey, per), so.gatherScatter(u, m))
    assert np.array_equal(m.mult, 1.0 / (QQt @ np.ones(QQt.shape[0])).reshape(shp, order="F"))
    if not any(per):
        M = so.generateMask(list("DDDD"), m)
        nx, ny = shp
        Rx = np.eye(nx)[1:-1]
        Ry = np.eye(ny)[1:-1]
        Rl = np.kron(Ry, Rx)
        assert np.array_equal(vec(M).astype(float), np.diag(Rl.T @ Rl))


def test_three_way_poisson_solve_p2d_explicit():
    """examples/p2d_explicit.jl:164-195: rank-deficient local system, full-rank global system and the
    matrix-free operator all solve the same Poisson problem."""
    nr, E = 6, 3
    m = so.make_mesh(nr, nr, E, E, (False, False), so.wavy)
    A, Bm, Q = so.kron_operators(m)
    nx = ny = nr * E
    ng = E * (nr - 1) + 1
    Rg = np.kron(np.eye(ng)[1:-1], np.eye(ng)[1:-1])
    f = 1.0 + m.x * m.y
    AA = Rg @ Q.T @ A @ Q @ Rg.T
    bb = Rg @ Q.T @ Bm @ vec(f)
    uu = (Q @ Rg.T @ np.linalg.solve(AA, bb)).reshape(nx, ny, order="F")
    M = so.generateMask(list("DDDD"), m).astype(float)
    b = so.gatherScatter(so.mask(so.mass(f, m), M), m)
    info = {}
    u = so.pcg(b, lambda v: so.opLHS(v, 1.0, 0.0, M, m), mult=m.mult, tol=1e-13, info=info)
    assert np.max(np.abs(u - uu)) < 1e-10 * np.max(np.abs(uu))
    # rank-deficient local system with explicit matrices through the same pcg (p2d_explicit.jl:168-172)
    Ml = np.diag(vec(M))
    QQt = Q @ Q.T
    Aloc = QQt @ Ml @ A
    uloc = so.pcg((QQt @ Ml @ Bm @ vec(f)).reshape(nx, ny, order="F"), Aloc, mult=m.mult, tol=1e-13)
    assert np.max(np.abs(uloc - uu)) < 1e-9 * np.max(np.abs(uu))


def test_operator_identities():
    m = so.make_mesh(7, 7, 4, 4, (False, False), so.wavy)
    u = so.splitmix_uniform(m.x.shape, seed=1)
    v = so.splitmix_uniform(m.x.shape, seed=2)
    L = lambda w: so.laplace(w, m.Dr, m.Ds, m.G11, m.G12, m.G22)
    assert abs(np.sum(v * L(u)) - np.sum(u * L(v))) < 1e-11 * abs(np.sum(v * L(u)))  # symmetric
    assert np.max(np.abs(L(np.ones_like(u)))) < 1e-10  # constants in the null space
    assert abs(np.sum(m.B * m.mult) - np.sum(m.B)) > 0  # mult really de-duplicates
    box = so.make_mesh(6, 6, 3, 5)
    assert abs(np.sum(box.B) - 4.0) < 1e-13  # sum of local mass = area of [-1,1]^2
    ann = so.make_mesh(10, 10, 4, 6, (False, True), so.annulus)
    assert abs(np.sum(ann.B) - math.pi * (1 - 0.25)) < 1e-8
    gs = so.gatherScatter(u, m)
    assert np.array_equal(so.gatherScatter(m.mult * gs, m), gs)  # gs(mult*gs(u)) == gs(u), exact: 1/2, 1/4 weights
    M = so.generateMask(list("DDDD"), m).astype(float)
    op = lambda w: so.opLHS(w, 1.0, 0.3, M, m)
    uc, vc = so.mask(so.gatherScatter(u, m), M), so.mask(so.gatherScatter(v, m), M)
    a, b = np.sum(vc * op(uc) * m.mult), np.sum(uc * op(vc) * m.mult)
    assert abs(a - b) < 1e-11 * abs(a)  # self-adjoint in the mult inner product on continuous masked fields


def test_generate_mask_conventions():
    m = so.make_mesh(4, 4, 2, 3, (False, True), so.fixU)
    M = so.generateMask(["D", "N", "D", "D"], m)  # periodic y overrides 'D' (mesh.jl:164-165)
    assert M.dtype == bool and not M[0, :].any() and M[-1, :].all() and M[1:, 0].all() and M[1:, -1].all()
    assert np.array_equal(so.mask(np.ones(M.shape), np.zeros(0)), np.ones(M.shape))  # length(M)==0 copies


def test_analytic_annulus_and_manufactured():
    """examples/p2d.jl as shipped: -lap u = 1 on 0.5<r<1, u=0 on both circles; closed form."""
    m = so.make_mesh(8, 8, 5, 5, [False, True], so.annulus)
    d = so.Diffusion(["D", "D", "N", "N"], m)
    one = lambda x, y, t: 1 + 0 * x
    zero = lambda x, y, t: 0 * x
    so.diffusion_simulate(d, setIC=zero, setBC=zero, setForcing=one, setVisc=one)
    r = np.hypot(m.x, m.y)
    exact = (1 - r ** 2) / 4 - (3.0 / 16.0) * np.log(r) / np.log(0.5)
    assert np.max(np.abs(d.u - exact)) < 1e-6
    assert abs(d.u.max() - 0.0316514) < 1e-6
    assert 150 <= d.pcg_iters[0] <= 200
    # manufactured solution on the box: ~36 iterations, spectral accuracy
    b = so.make_mesh(9, 9, 8, 8)
    M = so.generateMask(list("DDDD"), b).astype(float)
    us = np.sin(np.pi * b.x) * np.sin(np.pi * b.y)
    rhs = so.gatherScatter(so.mask(so.mass(2 * np.pi ** 2 * us, b), M), b)
    info = {}
    x = so.pcg(rhs, lambda v: so.opLHS(v, 1.0, 0.0, M, b), mult=b.mult, info=info)
    assert info["iters"] < 60 and np.max(np.abs(x - us)) < 1e-8


def test_bdf_coefficients():
    a, b = so.bdfExtK(np.zeros(4))  # steady (time.jl:48-50)
    assert np.array_equal(a, [1, 0, 0]) and np.array_equal(b, [0, 0, 0, 0])
    a, b = so.bdfExtK(np.array([0.03, 0.02, 0.01, 0.0]))  # BDF3/EXT3, dt = 0.01
    assert np.allclose(a, [3, -3, 1]) and np.allclose(b * 0.01, [11 / 6, -3, 1.5, -1 / 3])
    a, b = so.bdfExtK(np.array([0.01, 0.0, 0.0, 0.0]))  # first step: BDF1
    assert np.allclose(a, [1, 0, 0]) and np.allclose(b * 0.01, [1, -1, 0, 0])


def test_pcg_reference_semantics():
    m = so.make_mesh(5, 5, 3, 3)
    M = so.generateMask(list("DDDD"), m).astype(float)
    b = so.gatherScatter(so.mask(so.mass(np.ones(m.x.shape), m), M), m)
    op = lambda v: so.opLHS(v, 1.0, 0.0, M, m)
    info = {}
    x = so.pcg(b, op, mult=m.mult, maxiter=3, info=info)
    assert info["iters"] == 3 and info["warned"]  # pcg.jl:39
    info = {}
    so.pcg(0 * b, op, mult=m.mult, info=info)
    assert info["iters"] == 0  # loop not entered, pcg.jl:36
    assert so.splitmix_uniform((3, 2))[0, 0] == so.splitmix_uniform((6, 1))[0, 0]  # column-major stream


# ---- the Stokes reconstruction has no executable reference; these identities are what pins it ----------
def _stokes_setup(nr=7, E=3, deform=None):
    deform = deform or so.wavy
    V = so.make_mesh(nr, nr, E, E, (False, False), deform)
    P = so.make_mesh(nr - 2, nr - 2, E, E, (False, False), deform)
    return V, P, so.make_stokes(list("DDDD"), list("DDDD"), V, P, 1.0)


def test_gradT_and_diverT_are_exact_transposes():
    V, P, sks = _stokes_setup()
    rng = np.random.default_rng(0)
    u, v, w = (rng.standard_normal(V.x.shape) for _ in range(3))
    q = rng.standard_normal(P.x.shape)
    gx, gy = so.grad(u, V)
    tx, _ = so.gradT(v, V)
    _, ty = so.gradT(w, V)
    assert abs(np.sum(gx * v) - np.sum(u * tx)) < 1e-11 * np.sum(np.abs(gx * v))
    assert abs(np.sum(gy * w) - np.sum(u * ty)) < 1e-11 * np.sum(np.abs(gy * w))
    d = so.diver(u, v, V, sks.JrPV, sks.JsPV)
    qx, qy = so.diverT(q, V, sks.JrPV, sks.JsPV)
    assert abs(np.sum(d * q) - np.sum(qx * u) - np.sum(qy * v)) < 1e-11 * np.sum(np.abs(d * q))


def test_diver_of_a_linear_field_is_its_divergence():
    """diver integrates div(u) against the pressure basis: for u = (x, 2y) on the undeformed box, div = 3 and
    sum(diver) = 3 * area (the pressure interpolants sum to one)."""
    V, P, sks = _stokes_setup(deform=so.fixU)
    d = so.diver(V.x.copy(), 2.0 * V.y, V, sks.JrPV, sks.JsPV)
    assert abs(np.sum(d) - 3.0 * 4.0) < 1e-11


def test_schur_operator_is_symmetric_negative_semidefinite():
    V, P, sks = _stokes_setup()
    rng = np.random.default_rng(1)
    q = so.gatherScatter(rng.standard_normal(P.x.shape) * P.mult, P)
    r = so.gatherScatter(rng.standard_normal(P.x.shape) * P.mult, P)
    a = np.sum(r * so.opStokesLHS(q, sks) * P.mult)
    b = np.sum(q * so.opStokesLHS(r, sks) * P.mult)
    assert abs(a - b) < 1e-10 * abs(a)
    assert np.sum(q * so.opStokesLHS(q, sks) * P.mult) < 0
    one = np.ones(P.x.shape)  # constant pressure: in the null space for an enclosed (all-Dirichlet) flow
    assert np.max(np.abs(so.opStokesLHS(one, sks))) < 1e-10


def test_pressure_projection_removes_the_divergence():
    V, P, sks = _stokes_setup()
    rng = np.random.default_rng(2)
    vx = so.mask(so.gatherScatter(rng.standard_normal(V.x.shape) * V.mult, V), sks.Mvx)
    vy = so.mask(so.gatherScatter(rng.standard_normal(V.x.shape) * V.mult, V), sks.Mvy)
    r0 = so.makeStokesRHS(vx, vy, sks)
    nx, ny, p = so.pressureProject(vx, vy, np.zeros(P.x.shape), sks, tol=1e-10)
    r1 = so.makeStokesRHS(nx, ny, sks)
    assert np.max(np.abs(r1)) < 1e-8 and np.max(np.abs(r1)) < 1e-7 * np.max(np.abs(r0))
    # the corrected velocity stays continuous and keeps its Dirichlet values
    assert np.max(np.abs(so.gatherScatter(nx * V.mult, V) - nx)) < 1e-12
    assert np.max(np.abs(nx * (1 - sks.Mvx))) == 0.0


def test_gordonHall_reproduces_bilinear_quadrilaterals():
    """geom.jl:8-31 (with the corner matrix indexed [r, s]): exact for straight-sided quadrilaterals, identity included"""
    zr, _ = so.gausslobatto(6)
    zs, _ = so.gausslobatto(5)
    R, S = np.meshgrid(zr, zs, indexing="ij")
    for fx, fy in ((lambda r, s: r, lambda r, s: s),
                   (lambda r, s: 1 + 2 * r + 0.3 * s + 0.2 * r * s, lambda r, s: -0.5 + 0.1 * r + 1.5 * s - 0.25 * r * s)):
        x, y = so.gordonHall(fx(-1, zs), fx(1, zs), fx(zr, -1), fx(zr, 1), fy(-1, zs), fy(1, zs), fy(zr, -1), fy(zr, 1), zr, zs)
        assert np.max(np.abs(x - fx(R, S))) < 1e-13 and np.max(np.abs(y - fy(R, S))) < 1e-13
    # the literal reference form does not reproduce the identity map (the flagged deviation)
    x, _ = so.gordonHall(-1 + 0 * zs, 1 + 0 * zs, zr, zr, zs, zs, -1 + 0 * zr, 1 + 0 * zr, zr, zs, as_written=True)
    assert np.max(np.abs(x - R)) > 0.1


def test_explicit_argument_forms_match_the_mesh_forms():
    """lapl(u,M,Jr,Js,QQtx,QQty,...) lapl.jl:54-68 and mass(u,M,B,Jr,Js,QQtx,QQty,mult) mass.jl:32-50 with `[]` for
    Jr, Js (examples/p2d_explicit.jl:183-188) are mask(gs(lapl(u,msh))) / mask(gs(mass(u,msh)))."""
    m = so.make_mesh(6, 6, 3, 2, (False, True), so.wavy)
    M = so.generateMask(list("DDNN"), m).astype(np.float64)
    u = so.splitmix_uniform(m.x.shape, seed=9)
    a = so.lapl_explicit(u, M, [], [], m.QQtx, m.QQty, m.Dr, m.Ds, m.G11, m.G12, m.G22, m.mult)
    b = so.mask(so.gatherScatter(so.lapl(u, m), m), M)
    assert np.array_equal(a, b)
    a = so.mass_explicit(u, M, m.B, [], [], m.QQtx, m.QQty, m.mult)
    assert np.array_equal(a, so.mask(so.gatherScatter(so.mass(u, m), m), M))
    # dealiased: on the affine box mesh the integrand has degree 8 per direction, so every finer GLL rule with
    # 2n-3 >= 8 integrates it exactly: over-integration on 8 and on 11 points agree (and differ from the 5-point rule)
    mb, md, me = so.make_mesh(5, 5, 2, 2), so.make_mesh(8, 8, 2, 2), so.make_mesh(11, 11, 2, 2)
    ub = so.splitmix_uniform(mb.x.shape, seed=10)
    dea = [so.laplace_dealias(ub, so.interpMat(q.zr, mb.zr), so.interpMat(q.zs, mb.zs), mb.Dr, mb.Ds, q.G11, q.G12, q.G22)
           for q in (md, me)]
    assert np.max(np.abs(dea[0] - dea[1])) < 1e-11 * np.max(np.abs(dea[0]))
    assert np.max(np.abs(dea[0] - so.laplace(ub, mb.Dr, mb.Ds, mb.G11, mb.G12, mb.G22))) > 1e-3


def test_time_steppers_track_the_examples_analytic_solutions():
    """The two time-dependent examples carry closed-form solutions (their callbacks print the error against them):
    examples/d2d.jl:13-16 (forced heat equation, u = sin 2pi x sin 2pi y cos 2pi t) and examples/cd2d.jl:11-16
    (travelling wave u = sin pi(x - t) sin pi y, zero viscosity, periodic in x, dealiased advection 8 -> 12).  The
    oracle's BDF3/EXT3 drivers are run with the examples' own meshes and steps; the error stays at the level of the
    third-order time discretisation and shrinks by about 2^3 per halving of dt."""
    kx = ky = kt = 2.0

    def ut(x, y, t):
        return np.sin(kx * np.pi * x) * np.sin(ky * np.pi * y) * np.cos(kt * np.pi * t)

    def forcing(x, y, t):  # d2d.jl:29-31
        return ut(x, y, t) * ((kx ** 2 + ky ** 2) * np.pi ** 2) \
            - np.sin(kx * np.pi * x) * np.sin(ky * np.pi * y) * np.sin(kt * np.pi * t) * (kt * np.pi)

    msh = so.make_mesh(8, 8, 5, 5)  # d2d.jl:62-65
    errs = []
    for dt, steps in ((0.01, 20), (0.005, 40)):
        d = so.Diffusion(list("DDDD"), msh, Ti=0.0, Tf=1.0, dt=dt)
        so.diffusion_simulate(d, setIC=ut, setBC=lambda x, y, t: 0.0 * x, setForcing=forcing,
                              setVisc=lambda x, y, t: 1.0 + 0.0 * x, max_steps=steps)
        assert d.istep == steps and abs(d.time[0] - 0.2) < 1e-12
        errs.append(float(np.max(np.abs(d.u - ut(msh.x, msh.y, d.time[0])))))
    assert errs[0] < 5e-3 and errs[1] < errs[0] / 3.0, errs

    def wave(x, y, t):  # cd2d.jl:11-16 with ux = 1, uy = 0
        return np.sin(np.pi * (x - t)) * np.sin(np.pi * y)

    mV = so.make_mesh(8, 8, 5, 5, (True, False))   # cd2d.jl:54-60
    mD = so.make_mesh(12, 12, 5, 5, (True, False))
    errs = []
    for dt, steps in ((5e-3, 20), (2.5e-3, 40)):
        c = so.ConvectionDiffusion(list("NNDD"), mV, mD, 1.0 + 0.0 * mV.x, 0.0 * mV.x, Ti=0.0, Tf=1.0, dt=dt)
        c.u = np.asfortranarray(wave(mV.x, mV.y, 0.0))
        for _ in range(steps):
            so.convdiff_step(c)   # set0!/set∂!/setF!/setν! of the example leave ub = f = nu = 0
        assert c.istep == steps
        errs.append(float(np.max(np.abs(c.u - wave(mV.x, mV.y, c.time[0])))))
    assert errs[0] < 1e-3 and errs[1] < errs[0] / 3.0, errs
