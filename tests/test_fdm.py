"""FDM preconditioner: the reference's commented-out lapl_fdm (lapl.jl:105-119) and its set-up sketch
(examples/p2d_explicit.jl:109-141), restated in the oracle, and the overlapping form built on the device.
CPU part: the restatement against a Kronecker-assembled direct inverse, the host eigen-decomposition of libsemb against
the oracle's (scipy), symmetry / positive definiteness, and the iteration counts that justify the overlap.
GPU part: h = opM(r) against the oracle to 1e-12, and preconditioned pcg counts."""
import numpy as np
import pytest

import sem_oracle as so


def test_lapl_fdm_literal_inverts_the_element_neumann_laplacian():
    """lapl_fdm as written (no overlap): on a box mesh it is the exact inverse of the element-wise Neumann Laplacian
    A_e = By (x) Ax + Ay (x) Bx on the complement of its null space (constants per element)."""
    msh = so.make_mesh(5, 5, 2, 3)
    Bi, Sx, Sy, Sxi, Syi, Di = so.fdm_setup_reference(msh)
    u = so.splitmix_uniform(msh.x.shape, seed=4)
    # remove the element means w.r.t. B (the null mode that Di cuts off)
    B = 1.0 / Bi
    for ex in range(2):
        for ey in range(3):
            sl = np.s_[ex * 5:(ex + 1) * 5, ey * 5:(ey + 1) * 5]
            u[sl] -= np.sum(B[sl] * u[sl]) / np.sum(B[sl])
    Au = so.laplace(u, msh.Dr, msh.Ds, msh.G11, msh.G12, msh.G22)   # element-local D'GD (no gather-scatter): Neumann per element
    back = so.lapl_fdm(Au, Bi, Sx, Sy, Sxi, Syi, Di)
    assert np.max(np.abs(back - u)) < 1e-11 * np.max(np.abs(u))


@pytest.mark.parametrize("n,left,right", [(9, "N", "N"), (9, "D", "N"), (9, "N", "F"), (5, "D", "D"), (13, "F", "N"), (3, "N", "N")])
def test_host_tables_match_the_oracle(sem, n, left, right):
    """libsemb's Jacobi eigen-solver (semb_fdm_tables) against scipy.linalg.eigh in the oracle: same eigenvalues, and the
    same operator S f(lam) S' (eigenvectors are only defined up to sign / rotation inside an eigenspace)."""
    z, w = so.gausslobatto(n)
    D = so.derivMat(z)
    So, lo = so._fdm_1d_extended(D, w, 1.0, 1.0, 1.0, left, right)
    Sg, lg = sem.fdm_tables(D, w, left, right)
    fin = np.isfinite(lo)
    assert np.array_equal(fin, np.isfinite(lg))
    assert np.max(np.abs(lg[fin] - lo[fin])) < 1e-10 * np.max(np.abs(lo[fin]))
    f = lambda lam: np.where(np.isfinite(lam), 1.0 / (1.0 + np.where(np.isfinite(lam), lam, 0.0)), 0.0)
    Oo, Og = So @ np.diag(f(lo)) @ So.T, Sg @ np.diag(f(lg)) @ Sg.T
    assert np.max(np.abs(Og - Oo)) < 1e-11 * np.max(np.abs(Oo))


@pytest.mark.parametrize("nr,E,per,deform,bc,k", [(9, 4, (False, False), so.wavy, "DDDD", 0.0),
                                                 (6, 5, (True, False), so.wavy, "NNDN", 0.5),
                                                 (8, 4, (False, True), so.annulus, "DDNN", 0.0)])
def test_oracle_fdm_is_spd_and_cuts_the_iteration_count(nr, E, per, deform, bc, k):
    msh = so.make_mesh(nr, nr, E, E, per, deform)
    M = so.generateMask(list(bc), msh).astype(np.float64)
    P = so.fdm_schwarz(msh, bc, 1.0, k)
    cont = lambda seed: so.mask(so.gatherScatter(so.splitmix_uniform(msh.x.shape, seed=seed) * msh.mult, msh), M)
    u, v = cont(1), cont(2)
    dot = lambda a, b: float(np.sum(a * b * msh.mult))
    assert abs(dot(P(u), v) - dot(u, P(v))) < 1e-13 * abs(dot(P(u), v))     # symmetric in pcg's inner product
    assert dot(P(u), u) > 0 and dot(P(v), v) > 0                            # positive definite
    b = so.gatherScatter(so.mask(so.mass(np.ones(msh.x.shape), msh), M), msh)
    opA = lambda w_: so.opLHS(w_, 1.0, k, M, msh)
    i0, i1 = {}, {}
    x0 = so.pcg(b, opA, mult=msh.mult, tol=1e-9, info=i0)
    x1 = so.pcg(b, opA, opM=P, mult=msh.mult, tol=1e-9, info=i1)
    assert i1["iters"] * 3 < i0["iters"], (i0["iters"], i1["iters"])        # (4-8x in practice)
    assert np.max(np.abs(x1 - x0)) < 1e-7 * np.max(np.abs(x0))
    # true neighbour sizes instead of the element's own: same count within a couple of iterations
    i2 = {}
    so.pcg(b, opA, opM=so.fdm_schwarz(msh, bc, 1.0, k, uniform_neighbours=False), mult=msh.mult, tol=1e-9, info=i2)
    assert abs(i2["iters"] - i1["iters"]) <= max(3, i1["iters"] // 10)


GPU_CASES = [(9, 8, 8, (False, False), "wavy", "DDDD", 0.0), (8, 5, 5, (False, True), "annulus", "DDNN", 0.0),
             (5, 3, 4, (True, False), "wavy", "NNDN", 0.7), (13, 3, 2, (False, False), "wavy", "DDDD", 0.0),
             (3, 6, 5, (True, True), "wavy", "NNNN", 1.0), (9, 40, 3, (False, False), "wavy", "DNND", 0.0),
             (17, 2, 2, (False, False), "wavy", "DDDD", 0.0), (4, 1, 1, (False, False), "wavy", "DDDD", 0.0)]
DEFORMS = {"wavy": so.wavy, "annulus": so.annulus, "box": so.fixU}


@pytest.mark.gpu
@pytest.mark.parametrize("nr,Ex,Ey,per,deform,bc,k", GPU_CASES)
def test_gpu_fdm_apply_matches_oracle(sem, ctx, nr, Ex, Ey, per, deform, bc, k):
    om = so.make_mesh(nr, nr, Ex, Ey, per, DEFORMS[deform])
    gm = sem.Mesh.from_arrays(nr, nr, Ex, Ey, per, om.Dr, om.Ds, om.G11, om.G12, om.G22, om.B, ctx=ctx)
    try:
        Po, Pg = so.fdm_schwarz(om, bc, 1.3, k), sem.FdmPrecond(gm, bc, 1.3, k)
        M = so.generateMask(list(bc), om).astype(np.float64)
        r = so.mask(so.gatherScatter(so.splitmix_uniform(om.x.shape, seed=8) * om.mult, om), M)   # continuous, as pcg's residual
        ho, hg = Po(r), Pg(r)
        assert np.max(np.abs(hg - ho)) < 1e-12 * np.max(np.abs(ho))
        assert np.all(hg[M == 0.0] == 0.0)
        assert np.max(np.abs(so.gatherScatter(hg * om.mult, om) - hg)) < 1e-14 * np.max(np.abs(hg))   # continuous
        assert np.array_equal(Pg(r), hg)                                                              # deterministic
    finally:
        gm.free()


@pytest.mark.gpu
@pytest.mark.parametrize("nr,E,deform,bc,k", [(9, 8, "wavy", "DDDD", 0.0), (13, 4, "wavy", "DDDD", 0.0), (8, 6, "wavy", "DNDN", 0.5)])
def test_gpu_pcg_with_fdm(sem, ctx, nr, E, deform, bc, k):
    """pcg(b, opA; opM = FDM): same iteration count as the oracle's (short solves: below the rounding horizon), same
    solution, and several times fewer iterations than without."""
    om = so.make_mesh(nr, nr, E, E, (False, False), DEFORMS[deform])
    gm = sem.Mesh.from_arrays(nr, nr, E, E, (False, False), om.Dr, om.Ds, om.G11, om.G12, om.G22, om.B, ctx=ctx)
    try:
        M = so.generateMask(list(bc), om).astype(np.float64)
        b = so.gatherScatter(so.mask(so.mass(np.ones(om.x.shape), om), M), om)
        opo = lambda v: so.opLHS(v, 1.0, k, M, om)
        io, ig, i0 = {}, {}, {}
        xo = so.pcg(b, opo, opM=so.fdm_schwarz(om, bc, 1.0, k), mult=om.mult, tol=1e-8, info=io)
        xg = sem.pcg(b, sem.OpLHS(gm, 1.0, k, bc=bc), opM=sem.FdmPrecond(gm, bc, 1.0, k), mult=gm.mult, tol=1e-8, info=ig)
        sem.pcg(b, sem.OpLHS(gm, 1.0, k, bc=bc), mult=gm.mult, tol=1e-8, info=i0)
        assert ig["converged"] and abs(ig["iters"] - io["iters"]) <= (0 if io["iters"] <= 60 else 2), (ig, io["iters"])
        assert np.max(np.abs(xg - xo)) < 1e-6 * np.max(np.abs(xo))
        assert ig["iters"] * 3 < i0["iters"]
        # tight solves agree to 1e-10 (north_star)
        xo12 = so.pcg(b, opo, opM=so.fdm_schwarz(om, bc, 1.0, k), mult=om.mult, tol=1e-12)
        xg12 = sem.pcg(b, sem.OpLHS(gm, 1.0, k, bc=bc), opM=sem.FdmPrecond(gm, bc, 1.0, k), mult=gm.mult, tol=1e-12)
        assert np.max(np.abs(xg12 - xo12)) < 1e-10 * np.max(np.abs(xo12))
    finally:
        gm.free()


@pytest.mark.parametrize("n", list(range(3, 18)))
def test_interior_class_eigenvectors_are_even_or_odd(sem, n):
    """What the device kernel's even-odd tables rely on (semb_fdm.cu, FdmTab): the extended 1-D operator of an element with
    a neighbour on both sides is symmetric under i <-> n+1-i, so every eigenvector of libsemb's decomposition is even or
    odd, with ceil((n+2)/2) even and floor((n+2)/2) odd ones; a boundary class is not (it takes the full products)."""
    z, w = so.gausslobatto(n)
    D = so.derivMat(z)
    S, lam = sem.fdm_tables(D, w, "N", "N")
    assert np.all(np.isfinite(lam))
    ne = no = 0
    for m in range(n + 2):
        v = S[:, m]
        even, odd = np.max(np.abs(v - v[::-1])), np.max(np.abs(v + v[::-1]))
        assert min(even, odd) < 1e-10 * np.max(np.abs(v)), (n, m)
        ne, no = ne + (even < odd), no + (odd < even)
    assert ne == (n + 3) // 2 and no == (n + 2) // 2
    Sb, _ = sem.fdm_tables(D, w, "D", "N")
    assert any(min(np.max(np.abs(Sb[:, m] - Sb[::-1, m])), np.max(np.abs(Sb[:, m] + Sb[::-1, m]))) > 1e-6 for m in range(n))


@pytest.mark.parametrize("nr,Ex,Ey,per,deform,bc,nu,k", [(5, 3, 2, (False, False), so.wavy, "DNDD", 1.3, 0.5),
                                                        (4, 2, 3, (True, False), so.fixU, "NNDN", 1.0, 0.2),
                                                        (6, 2, 2, (False, True), so.wavy, "DDNN", 0.7, 0.0)])
def test_fdm_schwarz_equals_a_dense_assembly_with_direct_subdomain_solves(nr, Ex, Ey, per, deform, bc, nu, k):
    """Independent pin of the oracle's fdm_schwarz: the same preconditioner assembled WITHOUT eigen-decompositions -- every
    extended subdomain operator nu*(By (x) Ax + Ay (x) Bx) + k*(By (x) Bx) built as a dense matrix from the 1-D stiffness /
    mass pairs and solved directly -- must give the same h = opM(r) (the tensor solve of lapl_fdm, lapl.jl:112-119, is
    exactly that inverse written in the generalized eigenbasis)."""
    msh = so.make_mesh(nr, nr, Ex, Ey, per, deform)
    n = nr
    M = so.generateMask(list(bc), msh).astype(np.float64)
    z1, w = so.gausslobatto(n)
    D = so.derivMat(z1)
    A0 = D.T @ np.diag(w) @ D
    elavg = lambda a: a.reshape(n, Ex, n, Ey, order="F").mean(axis=(0, 2))
    Jw = msh.B / np.asfortranarray(np.outer(np.kron(np.ones(Ex), w), np.kron(np.ones(Ey), w)))
    hx, hy = elavg(Jw * np.sqrt(msh.G22 / msh.B)), elavg(Jw * np.sqrt(msh.G11 / msh.B))

    def one_d(e, E, periodic, bclo, bchi, h):
        """(A, B diagonal, active tile indices, global node of every tile index) of element e extended by one node"""
        A, Bd = np.zeros((n + 2, n + 2)), np.zeros(n + 2)
        A[1:n + 1, 1:n + 1] += A0 / h
        Bd[1:n + 1] += h * w
        glob = [None] + [e * n + i for i in range(n)] + [None]
        act = np.ones(n + 2, dtype=bool)
        if e > 0 or periodic:
            A[0:2, 0:2] += A0[n - 2:, n - 2:] / h
            Bd[0:2] += h * w[n - 2:]
            glob[0] = ((e - 1) % E) * n + n - 2
        else:
            act[0] = False
            act[1] = bclo != "D"
        if e < E - 1 or periodic:
            A[n:, n:] += A0[:2, :2] / h
            Bd[n:] += h * w[:2]
            glob[n + 1] = ((e + 1) % E) * n + 1
        else:
            act[n + 1] = False
            act[n] = bchi != "D"
        idx = np.nonzero(act)[0]
        return A[np.ix_(idx, idx)], Bd[idx], idx, [glob[i] for i in idx]

    r = so.mask(so.gatherScatter(so.splitmix_uniform(msh.x.shape, seed=8) * msh.mult, msh), M)
    cnt = np.zeros(msh.x.shape)
    parts = []
    for ex in range(Ex):
        for ey in range(Ey):
            Ax, Bx, ix, gx = one_d(ex, Ex, per[0], bc[0], bc[1], hx[ex, ey])
            Ay, By, iy, gy = one_d(ey, Ey, per[1], bc[2], bc[3], hy[ex, ey])
            # every tile node with a global image counts (Dirichlet nodes too: the oracle's counting ignores the mask)
            gxa = [g for g in ([((ex - 1) % Ex) * n + n - 2] if (ex > 0 or per[0]) else []) + [ex * n + i for i in range(n)]
                   + ([((ex + 1) % Ex) * n + 1] if (ex < Ex - 1 or per[0]) else [])]
            gya = [g for g in ([((ey - 1) % Ey) * n + n - 2] if (ey > 0 or per[1]) else []) + [ey * n + j for j in range(n)]
                   + ([((ey + 1) % Ey) * n + 1] if (ey < Ey - 1 or per[1]) else [])]
            cnt[np.ix_(gxa, gya)] += 1.0
            parts.append((Ax, Bx, gx, Ay, By, gy))
    W = 1.0 / np.sqrt(np.maximum(so.gatherScatter(cnt, msh), 1.0))
    rw = W * r
    z = np.zeros(msh.x.shape)
    for Ax, Bx, gx, Ay, By, gy in parts:
        T = rw[np.ix_(gx, gy)]
        # column-major vec: vec(Ax T By) = (By (x) Ax) vec(T)
        Ae = nu * (np.kron(np.diag(By), Ax) + np.kron(Ay, np.diag(Bx))) + k * np.kron(np.diag(By), np.diag(Bx))
        if k == 0.0 and np.linalg.matrix_rank(Ae) < Ae.shape[0]:
            U = np.linalg.pinv(Ae) @ T.reshape(-1, order="F")       # all-free subdomain: the null mode is cut off
        else:
            U = np.linalg.solve(Ae, T.reshape(-1, order="F"))
        z[np.ix_(gx, gy)] += U.reshape(T.shape, order="F")
    h_dense = so.mask(W * so.gatherScatter(z, msh), M)
    h_eig = so.fdm_schwarz(msh, bc, nu, k)(r)
    assert np.max(np.abs(h_eig - h_dense)) < 1e-10 * np.max(np.abs(h_dense))
