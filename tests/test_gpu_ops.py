"""GPU parity tests proper: the CUDA path (through the C ABI of libsemb.so) against the CPU oracle on
the same seeded inputs.  Tolerance: 1e-12 relative, normwise (||y - y_ref||inf / ||y_ref||inf), the
figure BASELINE.json's north_star states for FP64 operator applies; gather-scatter, mask and mult
are required to be BIT-exact (their sums are 2-term)."""
import numpy as np
import pytest

import sem_oracle as so

pytestmark = pytest.mark.gpu

TOL = 1e-12


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


DEFORMS = {"box": so.fixU, "wavy": so.wavy, "annulus": so.annulus}

# (nr, Ex, Ey, periodic, deform)
CASES = [
    (9, 8, 8, (False, False), "wavy"),     # BASELINE cfg1 (8x8, order 8)
    (8, 5, 5, (False, True), "annulus"),   # examples/p2d.jl as shipped
    (8, 8, 8, (False, False), "box"),
    (5, 3, 4, (True, False), "wavy"),
    (4, 1, 1, (False, False), "wavy"),     # single element
    (2, 3, 3, (False, False), "box"),      # smallest order
    (13, 2, 3, (False, False), "wavy"),    # order 12 (cfg3's order)
    (17, 2, 2, (True, True), "wavy"),      # largest templated size
    (9, 40, 6, (False, False), "wavy"),    # two strips (strip seam + partial strip)
    (6, 70, 3, (True, False), "wavy"),     # three strips, periodic x
    (7, 33, 5, (False, True), "annulus"),  # strip of one element
]


def make_pair(sem, ctx, nr, Ex, Ey, per, deform):
    om = so.make_mesh(nr, nr, Ex, Ey, per, DEFORMS[deform])
    gm = sem.Mesh.from_arrays(nr, nr, Ex, Ey, per, om.Dr, om.Ds, om.G11, om.G12, om.G22, om.B, ctx=ctx)
    return om, gm


@pytest.mark.parametrize("nr,Ex,Ey,per,deform", CASES)
def test_local_operators(sem, ctx, nr, Ex, Ey, per, deform):
    om, gm = make_pair(sem, ctx, nr, Ex, Ey, per, deform)
    try:
        u = so.splitmix_uniform(gm.shape)
        assert relerr(sem.lapl(u, gm), so.lapl(u, om)) < TOL
        nu = 1.0 + 0.5 * so.splitmix_uniform(gm.shape, seed=7) ** 2
        k = 2.0 + so.splitmix_uniform(gm.shape, seed=9)
        assert relerr(sem.hlmz(u, 0.7, 1.3, gm), so.hlmz(u, 0.7, 1.3, om)) < TOL
        assert relerr(sem.hlmz(u, nu, k, gm), so.hlmz(u, nu, k, om)) < TOL
        assert relerr(sem.lapl(u, nu, gm), so.lapl(u, om, nu)) < TOL
        assert np.array_equal(sem.mass(u, gm), so.mass(u, om))
    finally:
        gm.free()


@pytest.mark.parametrize("nr,Ex,Ey,per,deform", CASES)
def test_gather_scatter_mask_bitexact(sem, ctx, nr, Ex, Ey, per, deform):
    om, gm = make_pair(sem, ctx, nr, Ex, Ey, per, deform)
    try:
        u = so.splitmix_uniform(gm.shape, seed=3)
        assert np.array_equal(sem.gatherScatter(u, gm), so.gatherScatter(u, om))
        assert np.array_equal(gm.mult, om.mult)
        for bc in ("DDDD", "DDNN", "NDND", "NNNN"):
            Mo = so.generateMask(list(bc), om)
            Mg = sem.generateMask(list(bc), gm)
            assert np.array_equal(Mg, Mo)
            assert np.array_equal(sem.mask(u, Mg, gm), so.mask(u, Mo))
        assert np.array_equal(sem.mask(u, None, gm), u)
    finally:
        gm.free()


@pytest.mark.parametrize("nr,Ex,Ey,per,deform", CASES)
def test_fused_oplhs(sem, ctx, nr, Ex, Ey, per, deform):
    """opLHS = mask(gs(hlmz(u))) (diffusion.jl:36-45): fused strip + seam kernels vs the oracle, for
    every seam configuration (1 chunk ... one chunk per element row)."""
    om, gm = make_pair(sem, ctx, nr, Ex, Ey, per, deform)
    try:
        u = so.splitmix_uniform(gm.shape, seed=11)
        nu = 1.0 + 0.5 * so.splitmix_uniform(gm.shape, seed=7) ** 2
        for bc in ("DDDD", "DDNN"):
            M = so.generateMask(list(bc), om).astype(np.float64)
            ref0 = so.opLHS(u, 1.0, 0.0, M, om)
            ref1 = so.opLHS(u, nu, 1.5, M, om)
            for nch in sorted({1, 2, max(1, Ey // 2), Ey}):
                if nch > Ey:
                    continue
                gm.set_chunks(nch)
                out0 = sem.OpLHS(gm, 1.0, 0.0, bc=bc)(u)
                assert relerr(out0, ref0) < TOL, (bc, nch)
                out1 = sem.OpLHS(gm, nu, 1.5, M=M)(u)
                assert relerr(out1, ref1) < TOL, (bc, nch)
                # the result must be continuous across duplicated nodes and honour the mask exactly
                assert np.all(out0[M == 0.0] == 0.0)
                gsd = so.gatherScatter(out0 * om.mult, om)
                assert relerr(gsd, out0) < 1e-14
    finally:
        gm.free()


@pytest.mark.parametrize("nr,Ex,Ey,per,deform", [CASES[0], CASES[3], CASES[7], CASES[8], CASES[9], CASES[10],
                                                 (2, 300, 4, (True, True), "box")])
def test_fused_tail_equals_seam_kernels(sem, ctx, monkeypatch, nr, Ex, Ey, per, deform):
    """One apply is ONE launch (several ranks by default, PCG-mode applies always; here forced): the strip kernel's own
    CTAs finish the strip / chunk interfaces (semb_tail.cuh).  The separate seam kernels remain (SEMB_NO_TAIL=1, the NCCL fallback and the pipelined host
    twin use them): both must give the same BITS for every chunking (2-term interface sums, gatherScatter.jl:13)."""
    monkeypatch.setenv("SEMB_FORCE_TAIL", "1")   # (one rank: plain applies default to the seam kernels, PCG to the tail)
    om, gt = make_pair(sem, ctx, nr, Ex, Ey, per, deform)
    monkeypatch.delenv("SEMB_FORCE_TAIL")
    monkeypatch.setenv("SEMB_NO_TAIL", "1")
    _, gs = make_pair(sem, ctx, nr, Ex, Ey, per, deform)
    monkeypatch.delenv("SEMB_NO_TAIL")
    try:
        assert gt.fused_tail() and not gs.fused_tail()
        u = so.splitmix_uniform(gt.shape, seed=17)
        M = so.generateMask(list("DNDN"), om).astype(np.float64)
        for nch in sorted({1, 2, 3, Ey}):
            if nch > Ey:
                continue
            gt.set_chunks(nch)
            gs.set_chunks(nch)
            for kw in (dict(bc="DDDD"), dict(bc="NNNN"), dict(M=M)):
                a, b = sem.OpLHS(gt, 1.0, 0.3, **kw)(u), sem.OpLHS(gs, 1.0, 0.3, **kw)(u)
                assert np.array_equal(a, b), (nch, kw.keys())
            fu, fo = gt.field(u), gt.field()
            l0 = ctx.launch_count()
            gt.oplhs_device(fu, fo, nu=1.0, k=0.0, bc="DDDD")
            assert ctx.launch_count() - l0 == 1
            fu.free(); fo.free()
    finally:
        gt.free()
        gs.free()


def test_fused_is_deterministic(sem, ctx):
    om, gm = make_pair(sem, ctx, 9, 40, 8, (False, False), "wavy")
    try:
        u = so.splitmix_uniform(gm.shape, seed=5)
        a = sem.OpLHS(gm, 1.0, 0.0, bc="DDDD")(u)
        for _ in range(3):
            assert np.array_equal(sem.OpLHS(gm, 1.0, 0.0, bc="DDDD")(u), a)
        # chunking must not change a single bit (interface sums are commutative 2-term sums)
        gm.set_chunks(4)
        assert np.array_equal(sem.OpLHS(gm, 1.0, 0.0, bc="DDDD")(u), a)
    finally:
        gm.free()


@pytest.mark.parametrize("nr,ns,Ex,Ey,per", [(5, 7, 3, 2, (False, False)), (8, 6, 4, 3, (True, True)),
                                            (20, 20, 2, 2, (False, False))])
def test_generic_path(sem, ctx, nr, ns, Ex, Ey, per):
    """nr != ns (or nr > 17) takes the generic kernels: same contract."""
    om = so.make_mesh(nr, ns, Ex, Ey, per, so.wavy)
    gm = sem.Mesh.from_arrays(nr, ns, Ex, Ey, per, om.Dr, om.Ds, om.G11, om.G12, om.G22, om.B, ctx=ctx)
    try:
        assert gm.plan()["fast"] == 0
        u = so.splitmix_uniform(gm.shape, seed=13)
        assert relerr(sem.lapl(u, gm), so.lapl(u, om)) < TOL
        assert relerr(sem.hlmz(u, 0.5, 2.0, gm), so.hlmz(u, 0.5, 2.0, om)) < TOL
        assert np.array_equal(sem.gatherScatter(u, gm), so.gatherScatter(u, om))
        M = so.generateMask(list("DDDD"), om).astype(np.float64)
        assert relerr(sem.OpLHS(gm, 1.0, 0.3, bc="DDDD")(u), so.opLHS(u, 1.0, 0.3, M, om)) < TOL
    finally:
        gm.free()


@pytest.mark.parametrize("nr,Ex,Ey,per,deform", [(9, 8, 8, (False, False), "wavy"),
                                                 (8, 5, 5, (False, True), "annulus"),
                                                 (5, 3, 4, (True, False), "box")])
def test_mesh_geometry(sem, ctx, nr, Ex, Ey, per, deform):
    """Mesh(...) with a host deform closure: jac + B/G factors computed on device (jac.jl, mesh.jl:114-123).
    Differentiating coordinates amplifies rounding by ~N^2*Ex, hence the looser (still tiny) bound."""
    om = so.make_mesh(nr, nr, Ex, Ey, per, DEFORMS[deform])
    fn = {"box": sem.fixU, "wavy": sem.wavy, "annulus": sem.annulus}[deform]
    gm = sem.Mesh(nr, nr, Ex, Ey, per, fn, ctx=ctx)
    try:
        for name in ("x", "y"):
            assert relerr(getattr(gm, name), getattr(om, name)) < 1e-15
        for name in ("Jac", "Jaci", "rx", "ry", "sx", "sy", "B", "Bi", "G11", "G22"):
            assert relerr(getattr(gm, name), getattr(om, name)) < 2e-12, name
        # G12 vanishes analytically on orthogonal maps (annulus, box): measure it on the scale of G
        gscale = max(np.max(np.abs(om.G11)), np.max(np.abs(om.G22)))
        assert np.max(np.abs(gm.G12 - om.G12)) < 2e-12 * gscale
        assert np.array_equal(gm.mult, om.mult)
        J, Ji, rx, ry, sx, sy = sem.jac(om.x, om.y, om.Dr, om.Ds, msh=gm)
        assert relerr(J, om.Jac) < 2e-12 and relerr(sy, om.sy) < 2e-12
        # built-in device deformation agrees with the host closure path
        gd = sem.Mesh(nr, nr, Ex, Ey, per, {"box": "identity"}.get(deform, deform), ctx=ctx)
        try:
            for name in ("x", "y", "G11", "G22", "B"):
                assert relerr(getattr(gd, name), getattr(om, name)) < 5e-12, name
            assert np.max(np.abs(gd.G12 - om.G12)) < 5e-12 * gscale
        finally:
            gd.free()
    finally:
        gm.free()


def test_abu_generic(sem, ctx):
    """ABu(As,Br,u), ABu.jl:9-37, rectangular blocks and identities."""
    rng = np.random.default_rng(0)
    u = np.asfortranarray(rng.standard_normal((24, 18)))
    Br = rng.standard_normal((5, 8))   # 8-row chunks -> 5
    As = rng.standard_normal((4, 6))   # 6-col chunks -> 4
    for A, B in ((As, Br), (so.EMPTY, Br), (As, so.EMPTY), (so.EMPTY, so.EMPTY)):
        ref = so.ABu(A, B, u)
        out = sem.ABu(A, B, u, ctx=ctx)
        assert out.shape == ref.shape
        assert relerr(out, ref) < TOL if ref.size else True
    with pytest.raises(ValueError):
        sem.ABu(so.EMPTY, rng.standard_normal((5, 7)), u, ctx=ctx)
    # dense-QQt gather-scatter through ABu, gatherScatter.jl:13
    om = so.make_mesh(5, 5, 3, 3, (False, True), so.wavy)
    v = so.splitmix_uniform((15, 15))
    assert np.array_equal(sem.gatherScatter(v, om.QQtx, om.QQty), so.gatherScatter(v, om.QQtx, om.QQty))


def test_reductions(sem, ctx):
    om, gm = make_pair(sem, ctx, 9, 8, 8, (False, False), "wavy")
    try:
        a = so.splitmix_uniform(gm.shape, seed=1)
        b = so.splitmix_uniform(gm.shape, seed=2)
        fa, fb = gm.field(a), gm.field(b)
        ref = float(np.sum(a * b * om.mult))
        assert abs(gm.dot_mult(fa, fb) - ref) <= 1e-13 * np.sum(np.abs(a * b * om.mult))
        assert gm.norm_inf(fa) == float(np.max(np.abs(a)))
        # deterministic: same bits every time
        d0 = gm.dot_mult(fa, fb)
        assert all(gm.dot_mult(fa, fb) == d0 for _ in range(3))
        # device random fill reproduces the portable splitmix stream bit for bit
        fr = gm.field().fill_random(0x5EED)
        assert np.array_equal(fr.download(), so.splitmix_uniform(gm.shape, seed=0x5EED))
    finally:
        gm.free()


def test_errors(sem, ctx):
    om, gm = make_pair(sem, ctx, 4, 2, 2, (False, False), "box")
    try:
        with pytest.raises(ValueError):
            sem.lapl(np.zeros((3, 3)), gm)  # DimensionMismatch
        with pytest.raises(sem.SembError):
            f = gm.field()
            gm.lapl_device(f, f)  # aliasing
        with pytest.raises(TypeError):
            sem.pcg(np.zeros(gm.shape), lambda v: v)  # host closure: no CPU fallback
    finally:
        gm.free()


def test_explicit_argument_forms(sem, ctx):
    """lapl(u,M,Jr,Js,QQtx,QQty,Dr,Ds,G11,G12,G22,mult) lapl.jl:54-68 (plain and dealiased, examples/p2d_explicit.jl:183,
    semPS.jl:168), laplace(...) lapl.jl:70-103, mass(u,M,B,Jr,Js,QQtx,QQty,mult) mass.jl:32-50, mask on plain arrays."""
    m = so.make_mesh(6, 6, 3, 2, (False, True), so.wavy)
    md = so.make_mesh(9, 9, 3, 2, (False, True), so.wavy)
    M = so.generateMask(list("DDNN"), m).astype(np.float64)
    u = so.splitmix_uniform(m.x.shape, seed=9)
    Jr, Js = so.interpMat(md.zr, m.zr), so.interpMat(md.zs, m.zs)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    assert rel(sem.laplace(u, m.Dr, m.Ds, m.G11, m.G12, m.G22), so.laplace(u, m.Dr, m.Ds, m.G11, m.G12, m.G22)) < 1e-12
    assert rel(sem.laplace(u, Jr, Js, m.Dr, m.Ds, md.G11, md.G12, md.G22),
               so.laplace_dealias(u, Jr, Js, m.Dr, m.Ds, md.G11, md.G12, md.G22)) < 1e-12
    args = (M, [], [], m.QQtx, m.QQty, m.Dr, m.Ds, m.G11, m.G12, m.G22, m.mult)
    assert rel(sem.lapl(u, *args), so.lapl_explicit(u, *args)) < 1e-12
    args = (M, Jr, Js, m.QQtx, m.QQty, m.Dr, m.Ds, md.G11, md.G12, md.G22, m.mult)
    assert rel(sem.lapl(u, *args), so.lapl_explicit(u, *args)) < 1e-12
    for args in ((M, m.B, [], [], m.QQtx, m.QQty, m.mult), ([], md.B, Jr, Js, [], [], m.mult), (M, [], Jr, Js, m.QQtx, m.QQty, m.mult)):
        assert rel(sem.mass(u, *args), so.mass_explicit(u, *args)) < 1e-12
    assert np.array_equal(sem.mask(u, M), so.mask(u, M)) and np.array_equal(sem.mask(u, []), u)
    with pytest.raises(ValueError):
        sem.laplace(u[:-1], m.Dr, m.Ds, m.G11, m.G12, m.G22)


def test_gordonHall(sem, ctx):
    """geom.jl:8-31 on the device ABu kernels against the oracle, curved edges (a quarter annulus)"""
    zr, _ = so.gausslobatto(7)
    zs, _ = so.gausslobatto(6)
    R = lambda r: 0.75 + 0.25 * r
    th = lambda s: np.pi / 4 * (s + 1)
    edges = (R(-1) * np.cos(th(zs)), R(1) * np.cos(th(zs)), R(zr) * np.cos(th(-1)), R(zr) * np.cos(th(1)),
             R(-1) * np.sin(th(zs)), R(1) * np.sin(th(zs)), R(zr) * np.sin(th(-1)), R(zr) * np.sin(th(1)))
    for lit in (False, True):
        gx, gy = sem.gordonHall(*edges, zr, zs, as_written=lit)
        ox, oy = so.gordonHall(*edges, zr, zs, as_written=lit)
        assert np.max(np.abs(gx - ox)) < 1e-14 and np.max(np.abs(gy - oy)) < 1e-14
