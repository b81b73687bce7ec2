"""Two independent restatements of the reference path must agree: oracle/sem_oracle.py (NumPy, the reference's
GEMM-per-ABu structure) against oracle/sem_oracle.c (plain C loops per element).  Neither is pinned against the
true Julia reference (not installable here); agreement of two differently structured restatements, each citing
the reference lines it follows, is the strongest cross-check available on this side of the boundary."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import sem_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libsem_oracle_c.so")
dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def co():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(ROOT, "oracle", "sem_oracle.c")):
        if shutil.which("make") is None or shutil.which(os.environ.get("CC", "gcc")) is None:
            pytest.skip("no C toolchain to build the C oracle")
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    lib = C.CDLL(LIB)
    lib.so_mesh_create.restype = C.c_void_p
    lib.so_mesh_create.argtypes = [C.c_int] * 7
    lib.so_mesh_array.restype = dp
    lib.so_mesh_array.argtypes = [C.c_void_p, C.c_int]
    lib.so_mesh_free.argtypes = [C.c_void_p]
    for name, extra in (("so_laplace", []), ("so_mass", []), ("so_gather_scatter", [])):
        getattr(lib, name).argtypes = [C.c_void_p, dp, dp]
        getattr(lib, name).restype = None
    lib.so_hlmz.argtypes = [C.c_void_p, dp, dp, C.c_double, dp, C.c_double, dp]
    lib.so_hlmz.restype = None
    lib.so_oplhs.argtypes = [C.c_void_p, dp, dp, C.c_double, dp, C.c_double, dp, dp]
    lib.so_oplhs.restype = None
    lib.so_generate_mask.argtypes = [C.c_void_p, C.c_char_p, dp]
    lib.so_generate_mask.restype = None
    lib.so_pcg.argtypes = [C.c_void_p, dp, dp, C.c_double, dp, C.c_double, dp, C.c_double, C.c_double, C.c_longlong,
                           dp, dp, C.POINTER(C.c_int)]
    lib.so_pcg.restype = C.c_longlong
    lib.so_abu.argtypes = [dp, C.c_int, C.c_int, dp, C.c_int, C.c_int, dp, C.c_int, C.c_int, dp]
    lib.so_abu.restype = C.c_int
    lib.so_gausslobatto.argtypes = [C.c_int, dp, dp]
    lib.so_interpmat.argtypes = [C.c_int, dp, C.c_int, dp, dp]
    lib.so_interpmat.restype = None
    return lib


DEFORM = {"fixU": (0, so.fixU), "annulus": (1, so.annulus), "wavy": (2, so.wavy)}
CASES = [(9, 9, 4, 4, (False, False), "wavy"), (8, 8, 5, 5, (False, True), "annulus"), (5, 7, 3, 2, (True, False), "fixU"),
         (13, 13, 3, 3, (False, False), "wavy"), (4, 4, 6, 5, (True, True), "fixU")]
ARRAYS = ["x", "y", "Jac", "Jaci", "rx", "ry", "sx", "sy", "B", "Bi", "G11", "G12", "G22", "mult"]


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def field(lib, h, which, shape):
    n = shape[0] * shape[1]
    return np.ctypeslib.as_array(lib.so_mesh_array(h, which), shape=(n,)).reshape(shape, order="F").copy()


@pytest.mark.parametrize("nr,ns,Ex,Ey,per,deform", CASES)
def test_mesh_and_operators_agree(co, nr, ns, Ex, Ey, per, deform):
    kind, fn = DEFORM[deform]
    om = so.make_mesh(nr, ns, Ex, Ey, per, fn)
    h = co.so_mesh_create(nr, ns, Ex, Ey, int(per[0]), int(per[1]), kind)
    try:
        shape = om.x.shape
        for i, name in enumerate(ARRAYS):
            got, want = field(co, h, i, shape), getattr(om, name)
            if name == "mult":
                assert np.array_equal(got, want)  # {1, 1/2, 1/4} exactly (mesh.jl:94-96)
            elif name in ("G12", "ry", "sx"):
                # cross terms vanish analytically on orthogonal maps (annulus, box): what is left is rounding noise,
                # so they are measured on the scale of their diagonal partners
                scale = {"G12": om.G11, "ry": om.rx, "sx": om.sy}[name]
                assert float(np.max(np.abs(got - want)) / np.max(np.abs(scale))) < 1e-12, name
            else:
                assert relerr(got, want) < 1e-12, name
        Dr = np.ctypeslib.as_array(co.so_mesh_array(h, 20), shape=(nr * nr,)).reshape((nr, nr), order="F")
        assert relerr(Dr, om.Dr) < 1e-13
        # the C mesh's own arrays differ from NumPy's in the last bits (different GLL iteration / summation order), so the
        # operators are compared on a common input through each side's own mesh: the contract is the normwise 1e-12
        u = np.asfortranarray(so.splitmix_uniform(shape))
        out = np.zeros(shape, order="F")
        co.so_laplace(h, P(u), P(out))
        assert relerr(out, so.lapl(u, om)) < 1e-12
        co.so_mass(h, P(u), P(out))
        assert relerr(out, so.mass(u, om)) < 1e-13
        nu = np.asfortranarray(1.0 + 0.5 * om.x ** 2)
        kk = np.asfortranarray(2.0 + np.cos(om.y))
        co.so_hlmz(h, P(u), P(nu), 0.0, P(kk), 0.0, P(out))
        assert relerr(out, so.hlmz(u, nu, kk, om)) < 1e-12
        co.so_gather_scatter(h, P(u), P(out))
        assert np.array_equal(out, so.gatherScatter(u, om))  # two-term sums: bit-exact, also against the dense QQ^T
        for bc in ("DDDD", "DNND", "NNNN", "NDDN"):
            M = np.zeros(shape, order="F")
            co.so_generate_mask(h, bc.encode(), P(M))
            Mo = so.generateMask(list(bc), om).astype(np.float64)
            assert np.array_equal(M, Mo), bc
            co.so_oplhs(h, P(u), None, 0.7, None, 1.3, P(M), P(out))
            assert relerr(out, so.opLHS(u, 0.7, 1.3, Mo, om)) < 1e-12, bc
    finally:
        co.so_mesh_free(h)


@pytest.mark.parametrize("precond", [False, True])
def test_pcg_agrees(co, precond):
    nr, E = 8, 5
    om = so.make_mesh(nr, nr, E, E, (False, False), so.wavy)
    h = co.so_mesh_create(nr, nr, E, E, 0, 0, 2)
    try:
        shape = om.x.shape
        Mo = so.generateMask(list("DDDD"), om).astype(np.float64)
        M = np.asfortranarray(Mo)
        b = np.asfortranarray(so.gatherScatter(so.mask(so.mass(np.ones(shape), om), Mo), om))
        k = 3.0 if precond else 0.0
        b0 = 3.0 if precond else 0.0
        info = {}
        opM = (lambda r: r / om.B / b0) if precond else (lambda r: r)
        xo = so.pcg(b, lambda v: so.opLHS(v, 1.0, k, Mo, om), opM=opM, mult=om.mult, tol=1e-10, info=info)
        x = np.zeros(shape, order="F")
        res, warned = C.c_double(), C.c_int()
        it = co.so_pcg(h, P(b), None, 1.0, None, k, P(M), b0, 1e-10, -1, P(x), C.byref(res), C.byref(warned))
        assert not warned.value and not info["warned"]
        assert abs(it - info["iters"]) <= max(1, info["iters"] // 50), (it, info["iters"])  # within the oracle's own spread
        assert relerr(x, xo) < 1e-8
        # maxiter semantics (pcg.jl:39): the iterate after exactly maxiter iterations, flagged
        it = co.so_pcg(h, P(b), None, 1.0, None, k, P(M), b0, 1e-10, 7, P(x), C.byref(res), C.byref(warned))
        xo7 = so.pcg(b, lambda v: so.opLHS(v, 1.0, k, Mo, om), opM=opM, mult=om.mult, tol=1e-10, maxiter=7, info=info)
        assert it == 7 and warned.value == 1 and info["iters"] == 7 and info["warned"]
        assert relerr(x, xo7) < 1e-11
    finally:
        co.so_mesh_free(h)


def test_abu_rectangular_and_setup_helpers(co):
    rng = np.random.default_rng(7)
    u = np.asfortranarray(rng.standard_normal((12, 10)))
    As = np.asfortranarray(rng.standard_normal((7, 5)))   # acts on 5-column chunks: 10 -> 14 columns
    Br = np.asfortranarray(rng.standard_normal((6, 4)))   # acts on 4-row chunks:    12 -> 18 rows
    out = np.zeros((18, 14), order="F")
    assert co.so_abu(P(As), 7, 5, P(Br), 6, 4, P(u), 12, 10, P(out)) == 0
    assert relerr(out, so.ABu(As, Br, u)) < 1e-13
    out2 = np.zeros((18, 10), order="F")
    assert co.so_abu(None, 0, 0, P(Br), 6, 4, P(u), 12, 10, P(out2)) == 0   # `[]` = identity (ABu.jl:23)
    assert relerr(out2, so.ABu(so.EMPTY, Br, u)) < 1e-13
    assert co.so_abu(P(As), 7, 5, None, 0, 0, P(u), 12, 9, P(out)) == -1     # InexactError (ABu.jl:26)
    for n in (2, 3, 4, 5, 9, 13, 17):
        z, w = np.zeros(n), np.zeros(n)
        assert co.so_gausslobatto(n, P(z), P(w)) == 0
        zo, wo = so.gausslobatto(n)
        assert np.max(np.abs(z - zo)) < 1e-15 and np.max(np.abs(w - wo)) < 1e-14
    zo, _ = so.gausslobatto(14)
    zi, _ = so.gausslobatto(9)
    J = np.zeros((14, 9), order="F")
    co.so_interpmat(14, P(np.ascontiguousarray(zo)), 9, P(np.ascontiguousarray(zi)), P(J))
    assert relerr(J, so.interpMat(zo, zi)) < 1e-13


@pytest.mark.parametrize("nr,nrd,Ex,Ey,per,deform", [(8, 12, 3, 4, (True, False), "fixU"), (9, 14, 3, 3, (False, False), "wavy"),
                                                     (5, 7, 4, 2, (False, True), "annulus")])
def test_grad_and_dealiased_advect_agree(co, nr, nrd, Ex, Ey, per, deform):
    """grad.jl:15-34 and advect.jl:27-64 (the explicit convection term of cd2d): element-tile loops in C against the
    ABu-structured NumPy form, with and without the dealiasing mesh."""
    co.so_grad.argtypes = [C.c_void_p, dp, dp, dp]
    co.so_grad.restype = None
    co.so_advect.argtypes = [C.c_void_p, C.c_void_p, dp, dp, dp, dp]
    co.so_advect.restype = C.c_int
    kind, fn = DEFORM[deform]
    oV, oD = so.make_mesh(nr, nr, Ex, Ey, per, fn), so.make_mesh(nrd, nrd, Ex, Ey, per, fn)
    hV = co.so_mesh_create(nr, nr, Ex, Ey, int(per[0]), int(per[1]), kind)
    hD = co.so_mesh_create(nrd, nrd, Ex, Ey, int(per[0]), int(per[1]), kind)
    try:
        shape = oV.x.shape
        T = np.asfortranarray(np.sin(1.3 * oV.x) * np.cos(0.7 * oV.y) + 0.2 * oV.x * oV.y)
        vx = np.asfortranarray(1.0 + 0.3 * oV.y)
        vy = np.asfortranarray(-0.5 + 0.2 * oV.x ** 2)
        gx, gy, out = (np.zeros(shape, order="F") for _ in range(3))
        co.so_grad(hV, P(T), P(gx), P(gy))
        ox, oy = so.grad(T, oV)
        scale = max(np.max(np.abs(ox)), np.max(np.abs(oy)))
        assert np.max(np.abs(gx - ox)) / scale < 1e-12 and np.max(np.abs(gy - oy)) / scale < 1e-12
        assert co.so_advect(hV, None, P(T), P(vx), P(vy), P(out)) == 0
        assert relerr(out, so.advect(T, vx, vy, oV)) < 1e-12
        assert co.so_advect(hV, hD, P(T), P(vx), P(vy), P(out)) == 0
        assert relerr(out, so.advect(T, vx, vy, oV, oD)) < 1e-12
    finally:
        co.so_mesh_free(hV)
        co.so_mesh_free(hD)
