"""SURVEY 8f-3, second half: would the reference's commented-out element-wise fast-diagonalisation (FDM)
Laplacian solve (`lapl.jl:105-119`, construction sketch `examples/p2d_explicit.jl:109-141`) pay off as the
`opM` of `pcg` (`pcg.jl:16-60`)?  Checker-side experiment on the CPU oracle only (nothing here ships).

The sketch solves each element's Neumann problem (`eigen(Ax,Bx)` per direction, the null mode cut off at
`1/lambda > 1e8`, no overlap, no restriction, no coarse grid).  Wired the only way that keeps `pcg`'s vectors
continuous -- `opM(r) = mask(gs(w .* lapl_fdm(w .* r)))`, `w = sqrt(mult)` (symmetric), element lengths from the
averaged metric on deformed meshes -- it is measured against no preconditioner and against the reference's
diagonal `1/(B b0)` (`convectionDiffusion.jl:87-91`).  Result (order 8, f = 1, tol 1e-8, iterations):

    mesh            k      none   fdm   diag
    8x8   box       0      156    281   301
    8x8   wavy      0      284   1238   680
    16x16 box       0      292   1466   623
    16x16 wavy      0      575   6358  1186
    16x16 wavy      1      558   1426  1182
    16x16 wavy      100    218    171   645

i.e. for Poisson / weak Helmholtz (BASELINE configs 1-3) it is worse than no preconditioner -- the element-constant
error components are never corrected -- and only a strongly mass-dominated system gains.  That is why the FDM
preconditioner is documented as not built: making cfg3 practical needs an overlapping
Schwarz + coarse-grid method, which the reference does not contain in any form.

    python tests/tools/fdm_prototype.py [E ...]
"""
import os
import sys

import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sem_oracle as so  # noqa: E402


def make_prec(msh, M, k=0.0, nu=1.0):
    nr, ns, Ex, Ey = msh.nr, msh.ns, msh.Ex, msh.Ey
    Br, Bs = np.diag(msh.wr), np.diag(msh.ws)
    lr, Sr = sl.eigh(msh.Dr.T @ Br @ msh.Dr, Br)  # p2d_explicit.jl:126-127 on the reference element
    ls, Ss = sl.eigh(msh.Ds.T @ Bs @ msh.Ds, Bs)
    lr[0] = ls[0] = 0.0

    def elavg(a):
        return a.reshape(nr, Ex, ns, Ey, order="F").mean(axis=(0, 2))

    xr, ys = 1.0 / elavg(msh.rx), 1.0 / elavg(msh.sy)  # half-lengths (p2d_explicit.jl:114-116 uses node [1] only)
    Di = np.zeros((nr * Ex, ns * Ey))
    for ex in range(Ex):
        for ey in range(Ey):
            lam = nu * ((ys[ex, ey] / xr[ex, ey]) * lr[:, None] + (xr[ex, ey] / ys[ex, ey]) * ls[None, :]) \
                + k * xr[ex, ey] * ys[ex, ey]
            with np.errstate(divide="ignore"):
                d = 1.0 / lam
            d[np.abs(d) > 1e8] = 0.0  # p2d_explicit.jl:132-134
            Di[ex * nr:(ex + 1) * nr, ey * ns:(ey + 1) * ns] = d
    w = np.sqrt(msh.mult)
    SrT, SsT = Sr.T.copy(), Ss.T.copy()

    def lapl_fdm(b):  # lapl.jl:110-119 with Bi folded into S^-1 = S' B
        return so.ABu(Ss, Sr, so.ABu(SsT, SrT, b) * Di)

    return lambda r: so.mask(so.gatherScatter(w * lapl_fdm(w * r), msh), M)


def run(nr, E, deform, k):
    msh = so.make_mesh(nr, nr, E, E, (False, False), deform)
    M = so.generateMask(list("DDDD"), msh).astype(float)
    b = so.gatherScatter(so.mask(so.mass(np.ones(msh.x.shape), msh), M), msh)
    opA = lambda v: so.opLHS(v, 1.0, k, M, msh)  # noqa: E731
    its = []
    for opM in (lambda x: x, make_prec(msh, M, k=k), lambda r: so.mask(r / msh.B / max(k, 1.0), M)):
        info = {}
        so.pcg(b, opA, opM=opM, mult=msh.mult, tol=1e-8, info=info)
        its.append(info["iters"])
    print("%2dx%-2d %-5s k=%-5g none %5d  fdm %5d  diag %5d" % (E, E, deform.__name__, k, *its))


if __name__ == "__main__":
    for E in [int(a) for a in sys.argv[1:]] or [4, 8]:
        for deform in (so.fixU, so.wavy):
            for k in (0.0, 1.0, 100.0):
                run(9, E, deform, k)
